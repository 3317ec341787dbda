"""CPU-side checks: the C-ABI library loads and exports every symbol include/klhr_sm100.h
declares (no compute calls), host adaptation logic, world_size-2 gloo reduction."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    from klhr_b200 import _lib
    if not _lib.LIB_PATH.exists():
        g.build()
    header = (ROOT / "include" / "klhr_sm100.h").read_text()
    declared = set(re.findall(r"^(?:int|size_t|int64_t)\s+(klhr_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_lib.EXPORTS), (declared, set(_lib.EXPORTS))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.klhr_abi_version() == _lib.ABI_VERSION == 4
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.klhr_model_eval(None, 0, None, None, None, 0, None) < 0
    assert "model is NULL" in _lib.last_error()


def test_struct_layouts_match_header():
    """ctypes mirrors must have the C sizes (checked by compiling a probe with gcc)."""
    import ctypes as C
    from klhr_b200 import _lib
    src = '#include <stdio.h>\n#include "klhr_sm100.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(klhr_model_t),sizeof(klhr_fit_t),sizeof(klhr_direction_t),sizeof(klhr_trace_t),' \
          'sizeof(klhr_accum_t),sizeof(klhr_slice_t));return 0;}'
    exe = ROOT / "build" / "abi_probe"
    exe.parent.mkdir(exist_ok=True)
    (exe.parent / "abi_probe.c").write_text(src)
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(exe.parent / "abi_probe.c"), "-o", str(exe)], check=True)
    sizes = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    mine = [C.sizeof(t) for t in (_lib.ModelDesc, _lib.FitDesc, _lib.DirectionDesc, _lib.TraceDesc, _lib.AccumDesc,
                                     _lib.SliceDesc)]
    assert sizes == mine


def test_windowed_adaptation_matches_reference_schedule():
    from klhr_b200.adaptation import WindowedAdaptation
    from oracle.adapt import window_closures
    assert WindowedAdaptation(1000, 50, 2).closures == [50, 150, 350, 1000]
    for warm in (0, 10, 49, 51, 100, 101, 777, 1000, 15000):
        for ws in (5, 25, 50):
            if warm == ws:
                continue
            assert WindowedAdaptation(warm, ws, 2).closures == window_closures(warm, ws, 2)
    wa = WindowedAdaptation(100, 50, 2)
    assert [m for m in range(1, 300) if wa.window_closed(m)] == [50, 100]
    assert wa.next_closure(0) == 50 and wa.next_closure(50) == 100 and wa.next_closure(100) is None


def test_pooled_moments_and_pca_against_oracle():
    from klhr_b200.adaptation import OnlineMoments, OnlinePCA
    from oracle import adapt
    rng = np.random.default_rng(0)
    X = rng.normal(size=(600, 5)) * [1, 2, 3, 4, 5] + 3
    om = OnlineMoments(5, shift=torch.as_tensor(X[0]))
    for blk in np.array_split(X, 7):
        om.update(blk)
    rm = adapt.RunningMoments(5)
    for x in X:
        rm.update(x)
    assert np.allclose(om.mean().numpy(), rm.mean(), rtol=1e-12)
    assert np.allclose(om.var().numpy(), rm.var(), rtol=1e-10)
    assert torch.all(OnlineMoments(5).var() == 1)                     # N <= 2 -> ones (onlinemoments.py:20-23)
    pca = OnlinePCA(5, K=2)
    U = X - X.mean(0)
    pca.update(U)
    w, V = np.linalg.eigh(U.T @ U / len(U))
    assert np.allclose(pca.values(), w[::-1][:2] + 1e-10)
    assert np.allclose(np.abs(pca.vectors().T @ V[:, ::-1][:, :2]), np.eye(2), atol=1e-9)


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from klhr_b200.adaptation import OnlineMoments, OnlinePCA, allreduce_adaptation
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    X = rng.normal(size=(400, 6)) @ rng.normal(size=(6, 6))
    mine = np.array_split(X, world)[rank]                             # chains sharded over ranks
    mom, gmom, pca = OnlineMoments(6), OnlineMoments(6), OnlinePCA(6, K=2)
    mom.update(mine)
    gmom.update(-mine)
    pca.update(mine)
    allreduce_adaptation([mom, gmom], pca)
    res = dict(N=mom.N, n=pca.n, mean=mom.mean().numpy(), var=mom.var().numpy(), gvar=gmom.var().numpy(),
               vals=pca.values(), vecs=pca.vectors())
    torch.save(res, f"{out}.{rank}")
    dist.destroy_process_group()


def test_window_closure_allreduce_world_size_2_gloo(tmp_path):
    """The only collective of the sampler: raw pooled sums all-reduced at a window closure give
    every rank the single-process answer."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "res")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0", weights_only=False), torch.load(out + ".1", weights_only=False)
    rng = np.random.default_rng(123)
    X = rng.normal(size=(400, 6)) @ rng.normal(size=(6, 6))
    assert r0["N"] == r1["N"] == 400 and r0["n"] == 400
    for k in ("mean", "var", "gvar", "vals", "vecs"):
        assert np.array_equal(r0[k], r1[k])                            # bit-identical on every rank
    assert np.allclose(r0["mean"], X.mean(0)) and np.allclose(r0["var"], X.var(0, ddof=1))
    w, V = np.linalg.eigh(X.T @ X / 400)
    assert np.allclose(r0["vals"], w[::-1][:2] + 1e-10)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "klhr_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p


def test_fit_config_validation():
    import klhr_b200 as kb
    with pytest.raises(ValueError):
        kb.FitConfig(family="laplace")
    with pytest.raises(ValueError):
        kb.FitConfig(N=64)
    f = kb.FitConfig()
    assert np.isclose(f.w.sum(), 1) and f.descriptor().n_nodes == 8
    assert kb.FitConfig().for_dtype(torch.float32).gtol2 >= 1e-5


def test_draws_parquet_round_trip(tmp_path):
    """klhr_b200.output: the parquet layout the reference scripts intended (experiment_ar1.py:93-94)."""
    import numpy as np
    import pyarrow.parquet as pq
    from klhr_b200.output import write_draws
    rng = np.random.default_rng(0)
    draws = rng.normal(size=(7, 5, 3))
    n = write_draws(tmp_path / "d.parquet", draws, ["a", "b.1", "b.2"], chains=2, thin=3, first_iteration=1)
    t = pq.read_table(tmp_path / "d.parquet").to_pandas()
    assert n == 14 and list(t.columns) == ["chain", "iteration", "a", "b.1", "b.2"]
    assert t["iteration"].tolist()[:4] == [1, 1, 4, 4] and t["chain"].tolist()[:4] == [0, 1, 0, 1]
    assert np.array_equal(t[["a", "b.1", "b.2"]].to_numpy().reshape(7, 2, 3), draws[:, :2])
    one = rng.normal(size=(4, 3))                      # a single chain, (M, D) like the reference
    assert write_draws(tmp_path / "one.parquet", one, ["x", "y", "z"]) == 4
    with __import__("pytest").raises(ValueError):
        write_draws(tmp_path / "bad.parquet", draws, ["a"])


def test_argument_validation_of_every_entry_point_without_a_gpu():
    """Every export validates its arguments before any CUDA call: NULL descriptors, bad dtypes, negative sizes
    and inconsistent slice / direction settings come back as negative codes with a message; empty batches are 0."""
    import ctypes as C
    from klhr_b200 import _lib
    lib = _lib.load()
    model = _lib.ModelDesc(id=0, dim=4)
    fit = _lib.FitDesc(family=0, n_nodes=8, n1=12, n2=24, nb=8, initscale=0.1, tol=1e-12, scale_clip=600.0,
                       gtol1=1e-4, gtol2=1e-10, step_cap=2.0, c1=1e-4, basin=1e-3)
    sl = _lib.SliceDesc(w=1.0, lower=-float("inf"), upper=float("inf"), tol=1e-12, cap=8)
    acc = _lib.AccumDesc(thin=1)
    m, f, s = C.byref(model), C.byref(fit), C.byref(sl)
    # empty batches succeed without touching the device
    assert lib.klhr_run(m, f, None, 0, None, 0, 0, 0, 5, 1, None, None, None) == 0
    assert lib.klhr_slice_run(m, s, None, 0, None, 0, 0, 0, 5, 1, None, None, None) == 0
    assert lib.klhr_slice_replay(m, s, 0, None, None, None, None, None, None, 0, None) == 0
    assert lib.klhr_kl_eval(m, f, 0, None, None, None, None, None, None, 0, None) == 0
    assert lib.klhr_mh_run(m, 0, None, 0.5, 0, 0, 0, 3, 1, None, None, None) == 0
    assert lib.klhr_math_eval(0, None, None, 0, None) == 0
    # NULL buffers with a non-empty batch
    assert lib.klhr_run(m, f, None, 0, None, 8, 0, 0, 5, 1, None, None, None) < 0 and "theta" in _lib.last_error()
    assert lib.klhr_kl_eval(m, f, 0, None, None, None, None, None, None, 8, None) < 0
    assert lib.klhr_slice_replay(m, s, 0, None, None, None, None, None, None, 8, None) < 0
    assert lib.klhr_math_eval(2, None, None, 4, None) < 0 and "op" in _lib.last_error()
    # bad descriptors
    assert lib.klhr_run(None, f, None, 0, None, 8, 0, 0, 5, 1, None, None, None) < 0
    assert lib.klhr_run(m, None, None, 0, None, 8, 0, 0, 5, 1, None, None, None) < 0
    assert lib.klhr_run(m, f, None, 7, None, 8, 0, 0, 5, 1, None, None, None) < 0 and "dtype" in _lib.last_error()
    assert lib.klhr_run(m, f, None, 0, None, -1, 0, 0, 5, 1, None, None, None) < 0
    bad = _lib.SliceDesc(w=0.0, lower=-1.0, upper=1.0, tol=0.0, cap=8)
    assert lib.klhr_slice_run(m, C.byref(bad), None, 0, None, 8, 0, 0, 5, 1, None, None, None) < 0 and "w" in _lib.last_error()
    bad = _lib.SliceDesc(w=1.0, lower=0.5, upper=1.0, tol=0.0, cap=8)
    assert lib.klhr_slice_run(m, C.byref(bad), None, 0, None, 8, 0, 0, 5, 1, None, None, None) < 0
    bad_model = _lib.ModelDesc(id=99, dim=4)
    assert lib.klhr_model_eval(C.byref(bad_model), 0, None, None, None, 0, None) < 0
    pooled_only_s1 = _lib.AccumDesc(thin=1, pooled_s1=1)
    assert lib.klhr_run(m, f, None, 0, None, 8, 0, 0, 5, 1, C.byref(pooled_only_s1), None, None) < 0
    assert lib.klhr_outer_scratch_doubles(0, 10) == 0 and lib.klhr_outer_scratch_doubles(1000, 10) > 0
    del acc


def test_corr_pack_cholesky_layout_and_errors():
    """klhr_corr_pack_cholesky (host helper of the C ABI, no device needed): the packed stream holds, for column tile
    nt, k-pair p >= nt and lane (r8, k4), the B-fragment values L[8p + k4][8nt + r8] and L[8p + 4 + k4][8nt + r8]."""
    import ctypes as C
    from klhr_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for D in (128, 256):
        T = D // 8
        L = np.tril(rng.normal(size=(D, D)))
        n = lib.klhr_corr_pack_cholesky(None, D, None)
        assert n == 64 * T * (T + 1) // 2
        out = np.full(n, np.nan)
        assert lib.klhr_corr_pack_cholesky(L.ctypes.data, D, out.ctypes.data) == n
        off = lambda nt: 64 * (nt * T - nt * (nt - 1) // 2)
        for nt, p, lane in ((0, 0, 0), (0, T - 1, 31), (3, 3, 5), (T - 1, T - 1, 17), (7, T - 2, 30)):
            r8, k4 = lane >> 2, lane & 3
            base = off(nt) + ((p - nt) * 32 + lane) * 2
            assert out[base] == L[8 * p + k4, 8 * nt + r8] and out[base + 1] == L[8 * p + 4 + k4, 8 * nt + r8]
        assert not np.isnan(out).any()
        # every entry of the lower triangle's tiles appears exactly once
        assert np.isclose(np.sort(out[out != 0]), np.sort(L[L != 0])).all()
    assert lib.klhr_corr_pack_cholesky(None, 100, None) < 0 and "128 or 256" in _lib.last_error()
    assert lib.klhr_corr_pack_cholesky(None, 128, np.empty(1).ctypes.data) < 0
