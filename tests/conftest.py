import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without CUDA skips the gpu tests instead of failing in them.  On a GPU box a missing
    libklhr_sm100.so is NOT a reason to skip: the product has no CPU fallback and the tests must fail loudly."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_tape(name):
    """Golden tape of the unmodified reference (made by oracle/make_golden.py)."""
    t = dict(np.load(GOLDEN / f"{name}.npz"))
    meta = json.loads(str(t.pop("meta_json")))
    data = json.loads(str(t.pop("data_json")))
    return t, meta, data


@pytest.fixture(scope="session")
def tapes():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_tape(name)
        return cache[name]
    return get
