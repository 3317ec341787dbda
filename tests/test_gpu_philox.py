"""The in-kernel Philox4x32-10 streams: known-answer vectors (Random123) through the C ABI, and every variate the
step kernels emit re-derived on the CPU from (seed, chain, draw) by oracle/philox.py."""
import ctypes as C

import numpy as np
import pytest
import torch

import klhr_b200 as kb
from gpu_util import device, fit_pair, up
from oracle import philox as P

pytestmark = pytest.mark.gpu


def _device_blocks(rows):
    lib = kb._lib.load()
    inp = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=device()).to(torch.int32).contiguous()
    out = torch.zeros(len(rows), 4, dtype=torch.int32, device=device())
    rc = lib.klhr_philox_eval(inp.data_ptr(), out.data_ptr(), len(rows), None)
    assert rc == 0
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint32)


def test_known_answer_vectors_on_device():
    rows = [list(c) + list(k) for c, k, _ in P.KAT]
    rows = [[(v - (1 << 32)) if v >= (1 << 31) else v for v in r] for r in rows]       # as int32 bit patterns
    got = _device_blocks(rows)
    for g, (_, _, out) in zip(got, P.KAT):
        assert tuple(int(x) for x in g) == out
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 1 << 32, size=(5000, 6), dtype=np.uint64)
    dev = _device_blocks(rnd.astype(np.int64).astype(np.uint32).view(np.int32).astype(np.int64))
    ref = np.stack(P.philox4x32_10(*(rnd[:, k] for k in range(6))), axis=1)
    assert np.array_equal(dev.astype(np.uint64), ref)


@pytest.mark.parametrize("kernel", ["lane", "tile", "octet"])
def test_emitted_variates_are_the_documented_function_of_seed_chain_draw(kernel):
    """u (53-bit) bitwise, z_prop to 1e-13, z_init and the direction to the accuracy of the fp32 approximate
    units -- including the |z| > 4 tail of the direction normals -- for chains with 64-bit ids and a draw index
    beyond 2^32."""
    D, B, seed = 100, 3000, 0x1234_5678_9ABC_DEF0
    chain_offset, draw_offset = (1 << 33) + 5, (1 << 32) + 7
    model = kb.BSModel(stan_file="stan/normal.stan", data={"D": D}, device=device())
    kfit, _ = fit_pair("gauss")
    kfit.force_octet, kfit.force_tile = kernel == "octet", kernel == "tile"
    sd = np.linspace(0.5, 2.0, D)
    cols = np.zeros((2, D))
    cols[0, 1], cols[1, 50] = 3.0, -2.0
    cdf = np.array([0.3, 0.8, 1.0])
    direction = kb.Direction(mean_cols=up(cols), sd=up(sd), cdf=up(cdf), n_zero_cols=1)
    S = 2
    th = up(np.zeros((B, D)))
    tr = kb.Trace(S, B, D, 2, torch.float64, device(), variates=True, rho=True)
    kb.run(model, kfit, th, S, seed, direction, chain_offset=chain_offset, draw_offset=draw_offset, trace=tr)
    torch.cuda.synchronize()
    chains = chain_offset + np.arange(B)
    n_tail = 0
    for s in range(S):
        u_col, z_init, z_prop, u = P.chain_scalars(seed, chains, draw_offset + s)
        assert np.array_equal(tr.u[s].cpu().numpy(), u)
        assert np.allclose(tr.z_prop[s].cpu().numpy(), z_prop, rtol=1e-13, atol=1e-13)
        assert np.allclose(tr.z_init[s].cpu().numpy(), z_init, rtol=0, atol=5e-6)
        z = P.direction_normals(seed, chains, draw_offset + s, D)
        j = np.searchsorted(cdf.astype(np.float32), u_col.astype(np.float32), side="right")
        mean = np.where((j < 2)[:, None], cols[np.minimum(j, 1)], 0.0)
        x = mean.astype(np.float32) + sd.astype(np.float32) * z
        rho = x / np.linalg.norm(x + 1e-12, axis=1, keepdims=True)
        assert np.allclose(tr.rho[s].cpu().numpy(), rho, rtol=0, atol=2e-6)
        n_tail += int((np.abs(z) > 4).sum())
    assert n_tail >= 10          # 6e5 normals: ~38 expected beyond 4 sigma, all reproduced above
