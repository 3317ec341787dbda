"""Exploration (not a test): error statistics of the CUDA replay step vs the oracle."""
import glob, os, sys, time
import numpy as np, torch
HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from conftest import load_tape
from gpu_util import replay_both, rel_errors

for path in sorted(glob.glob("tests/golden/*.npz")):
    name = os.path.basename(path)[:-4]
    t, meta, data = load_tape(name)
    if "rho" not in t:          # stats_* / freerun_* tapes hold no per-draw inputs
        continue
    for dtype in (torch.float64, torch.float32):
        t0 = time.time()
        gpu, ref = replay_both(meta["model"], data, meta["family"], t["theta0"], t["rho"], t["z_init"],
                               t["z_prop"], t["u"], init4=t.get("init4"), dtype=dtype,
                               xw=(t["x_nodes"], t["w_nodes"]))
        em, es, ez, er = rel_errors(gpu, ref, meta["family"])
        q = lambda a: "%.1e/%.1e/%.1e" % (np.median(a), np.quantile(a, 0.999), a.max())
        print(f"{name:28s} {str(dtype)[6:]:8s} m {q(em)} ls {q(es)} zp {q(ez)} r {q(er)} "
              f"acc_mism {(gpu['accept'] != ref['accept']).sum()} evals_mism {(gpu['evals'] != ref['evals']).sum()}"
              f" theta {np.abs(gpu['theta'] - ref['theta']).max():.1e}", flush=True)
