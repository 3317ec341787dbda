"""The dense kernel (csrc/klhr_densek.cuh): corr-normal D = 128 / 256 through the Cholesky factor of the precision
(V = L' rho, w = L' theta) on the FP64 tensor cores.  Parity against the batched oracle (which evaluates
rho' P rho and theta' P rho directly, stan/corr-normal.stan:5-13) and against the octet kernel."""
import numpy as np
import pytest
import torch

import klhr_b200 as kb
from oracle import batched, stan_models
from gpu_util import device, fit_pair, rel_errors, replay_both, up

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,rho", [(256, 0.9), (128, 0.5), (256, 0.99)])
def test_dense_kernel_is_dispatched_and_replays_the_oracle(N, rho):
    """Replay: same theta, rho, variates -> fitted (m, log s), proposal, ratio within 1e-10, flags identical."""
    data = {"N": N, "rho": rho}
    model = kb.BSModel(stan_file="stan/corr-normal.stan", data=data, device=device())
    kfit, _ = fit_pair("gauss")
    info = kb.launch_info(model, kfit, dtype=torch.float64, free_running=True, accumulate=False, device=device())
    # free-running launches: the warp-specialised kernel (8 tensor + 8 producer warps; w = L'theta fp64, two fp32 direction
    # buffers, per-warp cp.async rings); replay below runs on dense_kernel, sample() rows too
    want_smem = 32 * (N + 8) * 8 + 8 * 4 * (N // 64) * 64 * 8 + (8 * 32 * 2 + 2 * 32 + 2 * 5 * 32) * 8 + 2 * 32 * (N + 4) * 4 + N * 4
    assert info["threads"] == 512 and info["ctas_per_sm"] == 1 and info["smem"] == want_smem
    rng = np.random.default_rng(N)
    for B in (1, 77, 1000):                                   # ragged last CTA (32 chains per CTA)
        theta = rng.normal(size=(B, N)) * 0.7
        r = rng.normal(size=(B, N))
        r /= np.linalg.norm(r, axis=1, keepdims=True)
        z_init, z_prop, u = rng.normal(size=B), rng.normal(size=B), rng.random(B)
        gpu, ref = replay_both("corr-normal", data, "gauss", theta, r, z_init, z_prop, u)
        em, es, ez, er = rel_errors(gpu, ref, "gauss")
        assert max(em.max(), es.max(), ez.max(), er.max()) <= 1e-10
        assert np.array_equal(gpu["accept"], ref["accept"])
        assert np.allclose(gpu["theta"], ref["theta"], rtol=1e-10, atol=1e-10)
        oct_, _ = replay_both("corr-normal", data, "gauss", theta, r, z_init, z_prop, u, force_octet=True)
        assert np.allclose(gpu["eta"], oct_["eta"], rtol=1e-11, atol=1e-11)
        assert np.array_equal(gpu["accept"], oct_["accept"])


def test_dense_kernel_long_launch_tracks_the_octet_kernel():
    """w = L' theta is carried in registers across the draws of a launch (w += zp V): after 300 draws the state
    must still agree with the octet kernel, which recomputes theta' P rho from theta every draw, and a launch
    split in two (w rebuilt from theta) must agree with the single launch to rounding."""
    N, B, S = 256, 96, 300
    model = kb.BSModel(stan_file="stan/corr-normal.stan", data={"N": N, "rho": 0.9}, device=device())
    kfit, _ = fit_pair("gauss")
    rng = np.random.default_rng(5)
    theta0 = rng.normal(size=(B, N))
    cols = up(rng.normal(size=(2, N)) * 0.3)
    direction = kb.Direction(mean_cols=cols, sd=up(0.5 + rng.random(N)), cdf=up(np.array([0.3, 0.8, 1.0])), n_zero_cols=1)
    a, b, c = up(theta0), up(theta0), up(theta0)
    acc_a = torch.zeros(B, dtype=torch.int64, device=device())
    acc_b = torch.zeros_like(acc_a)
    kb.run(model, kfit, a, S, 9, direction, accept_count=acc_a)
    kfit_o, _ = fit_pair("gauss")
    kfit_o.force_octet = True
    kb.run(model, kfit_o, b, S, 9, direction, accept_count=acc_b)
    kb.run(model, kfit, c, 120, 9, direction)
    kb.run(model, kfit, c, S - 120, 9, direction, draw_offset=120)
    torch.cuda.synchronize()
    assert torch.equal(acc_a, acc_b)
    scale = float(b.abs().max())
    assert float((a - b).abs().max()) <= 1e-9 * scale
    assert float((a - c).abs().max()) <= 1e-11 * scale
    assert not torch.equal(a, up(theta0))


def test_dense_kernel_thinned_draws_match_octet_kernel():
    """sample(M, thin) rows written by the dense kernel (adaptation on: pooled moments + PCA directions) against
    the same sampler forced onto the octet kernel; the last kept row is the final state."""
    N, B = 128, 700
    model = kb.BSModel(stan_file="stan/corr-normal.stan", data={"N": N, "rho": 0.5}, device=device())
    out = []
    for force in (False, True):
        s = kb.KLHR(model, seed=3, chains=B, warmup=120, windowsize=25, device=device())
        s._fit.force_octet = force
        s.run(130)
        rows = s.sample(7, thin=5)
        torch.cuda.synchronize()
        assert torch.equal(rows[-1], s.theta)
        assert abs(s.acceptance_probability - 1.0) < 1e-12     # Gaussian target, Gaussian family: always accepted
        out.append(rows)
    scale = float(out[1].abs().max())
    assert float((out[0] - out[1]).abs().max()) <= 1e-7 * scale
    assert float((out[0][1] - out[0][0]).abs().max()) > 0
