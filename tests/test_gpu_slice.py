"""Slice sampling along KLHR's adapted directions (reference slice.py): the CUDA path
through the C ABI against tapes of the unmodified reference and against the batched oracle."""
import numpy as np
import pytest
import torch

import klhr_b200 as kb
from conftest import load_tape
from gpu_util import device, up
from klhr_b200.diagnostics import chain_summary
from oracle import batched, stan_models

pytestmark = pytest.mark.gpu

TAPES = ["slice_normal_d2", "slice_funnel_d2", "slice_funnel_d11", "slice_illnormal_d100",
         "slice_rosenbrock_d4_w3", "slice_ark_t200_method2"]


@pytest.mark.parametrize("name", TAPES)
def test_slice_replay_matches_reference_tape(name):
    """Every draw of the reference tape (theta0, rho, e, u0, shrinkage uniforms) through klhr_slice_replay:
    same number of shrinkage proposals and value calls, accepted coordinate to 1e-10 of the slice width."""
    t, meta, data = load_tape(name)
    model = kb.BSModel(stan_file=f"stan/{meta['model']}.stan", data=data, device=device())
    cap = t["shrink_u"].shape[1]
    cfg = kb.SliceConfig(w=meta["w"], tol=meta["tol"], cap=cap)
    th = up(t["theta0"])
    tr = kb.slice_replay(model, cfg, th, up(t["rho"]), up(t["e"]), up(t["u0"]), up(t["shrink_u"]))
    torch.cuda.synchronize()
    x1 = tr.zp[0].cpu().numpy()
    n_shrink = tr.slice_n[0].cpu().numpy()
    evals = tr.evals[0].cpu().numpy()
    # discrete decisions (inside / outside the slice) can only differ where a test value sits within
    # rounding of the level: allow at most 0.1 % of the draws, exact agreement everywhere else
    same = (n_shrink == t["n_shrink"]) & (evals == t["evals"])
    assert same.mean() >= 0.999, same.mean()
    width = meta["w"] + np.abs(t["x1"])
    assert (np.abs(x1 - t["x1"])[same] <= 1e-10 * width[same]).all()
    th1 = np.vstack([t["theta0"][1:], t["theta_last"][None]])
    err = np.abs(th.cpu().numpy() - th1)[same]
    assert err.max() <= 1e-10 * (1 + np.abs(th1).max())
    assert (tr.accept[0].cpu().numpy() == 1).all()


def test_slice_replay_fp32_and_bounds_and_exhausted_tape():
    rng = np.random.default_rng(3)
    B, D = 4096, 10
    data = {"D": D - 1}
    model = kb.BSModel(stan_file="stan/funnel.stan", data=data, device=device())
    om = stan_models.make_model("funnel", data)
    theta = rng.normal(size=(B, D))
    rho = rng.normal(size=(B, D))
    rho /= np.linalg.norm(rho, axis=1, keepdims=True)
    e, u0, su = rng.exponential(size=B), rng.random(B), rng.random((B, 32))
    # fp32: the oracle sees the same rounded inputs; decisions agree except at rounding distance of the level
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    cfg = kb.SliceConfig(w=2.0, cap=32)
    th = up(theta, torch.float32)
    tr = kb.slice_replay(model, cfg, th, up(rho, torch.float32), up(e, torch.float32), up(u0, torch.float32),
                         up(su, torch.float32))
    ref = batched.slice_step(om, f32(theta), f32(rho), f32(e), f32(u0), f32(su), w=2.0)
    same = tr.slice_n[0].cpu().numpy() == ref["n_shrink"]
    assert same.mean() > 0.99
    assert np.abs(tr.zp[0].double().cpu().numpy() - ref["x1"])[same].max() < 2e-4 * 4
    # bounds on the line coordinate (slice.py:20-21,127-128), fp64
    cfgb = kb.SliceConfig(w=0.5, lower=-0.25, upper=0.75, cap=32)
    th = up(theta)
    trb = kb.slice_replay(model, cfgb, th, up(rho), up(e), up(u0), up(su))
    refb = batched.slice_step(om, theta, rho, e, u0, su, w=0.5, lower=-0.25, upper=0.75)
    assert np.array_equal(trb.slice_n[0].cpu().numpy(), refb["n_shrink"])
    assert np.array_equal(trb.evals[0].cpu().numpy(), refb["evals"])
    x1 = trb.zp[0].cpu().numpy()
    assert np.allclose(x1, refb["x1"], rtol=0, atol=1e-12) and x1.min() >= -0.25 and x1.max() <= 0.75
    # no shrinkage uniforms left: the chain stays put and reports cap + 1
    cfg2 = kb.SliceConfig(cap=2)
    th = up(theta)
    tr2 = kb.slice_replay(model, cfg2, th, up(rho), up(e), up(u0), up(np.full((B, 2), np.nan)))
    assert (tr2.slice_n[0].cpu().numpy() == 3).all() and np.array_equal(th.cpu().numpy(), theta)
    # empty batch, wrong shapes, bad configuration
    kb.slice_replay(model, cfg2, up(np.zeros((0, D))), up(np.zeros((0, D))), up(np.zeros(0)), up(np.zeros(0)),
                    up(np.zeros((0, 2))))
    with pytest.raises(ValueError):
        kb.slice_replay(model, cfg2, up(theta), up(rho), up(e), up(u0), up(su))
    with pytest.raises(ValueError):
        kb.SliceConfig(w=0.0)
    with pytest.raises(ValueError):
        kb.SliceConfig(lower=0.5)


@pytest.mark.parametrize("model_name,data,w", [("funnel", {"D": 1}, 1.0), ("ill-normal", {"D": 100}, 1.0),
                                              ("corr-normal", {"N": 40, "rho": 0.9}, 0.7),
                                              ("rosenbrock", {"D": 2}, 3.0), ("ar1", {"N": 300}, 1.0)])
def test_slice_free_run_equals_oracle_on_emitted_variates(model_name, data, w):
    """Free-running kernel (in-kernel Philox directions and variates) with traces on: the emitted
    (rho, e, u0, shrinkage uniforms) replayed through the batched oracle reproduce every draw."""
    model = kb.BSModel(stan_file=f"stan/{model_name}.stan", data=data, device=device())
    om = stan_models.make_model(model_name, data)
    B, S, D, cap = 1500, 4, model.dim(), 40
    rng = np.random.default_rng(1)
    theta0 = rng.normal(size=(B, D)) * 0.5
    th = up(theta0)
    cfg = kb.SliceConfig(w=w, cap=cap)
    tr = kb.Trace(S, B, D, 0, torch.float64, device(), variates=True, rho=True, slice_cap=cap)
    evals_total = torch.zeros(1, dtype=torch.int64, device=device())
    acc = torch.zeros(B, dtype=torch.int64, device=device())
    kb.slice_run(model, cfg, th, S, seed=11, trace=tr, evals_total=evals_total, accept_count=acc)
    torch.cuda.synchronize()
    cur = theta0
    tot = 0
    for k in range(S):
        rho = tr.rho[k].cpu().numpy()
        assert np.allclose(np.linalg.norm(rho + 1e-12, axis=1), 1.0, atol=1e-12)
        ref = batched.slice_step(om, cur, rho, tr.z_init[k].cpu().numpy(), tr.u[k].cpu().numpy(),
                                 tr.slice_u[k].cpu().numpy(), w=w)
        n_dev = tr.slice_n[k].cpu().numpy()
        assert (n_dev <= cap).all()
        assert np.array_equal(n_dev, ref["n_shrink"])
        assert np.array_equal(tr.evals[k].cpu().numpy(), ref["evals"])
        scale = w + np.abs(ref["x1"])
        assert (np.abs(tr.zp[k].cpu().numpy() - ref["x1"]) <= 1e-10 * scale).all()
        cur = ref["theta"]
        tot += int(ref["evals"].sum())
    assert np.allclose(th.cpu().numpy(), cur, rtol=1e-11, atol=1e-11)
    assert int(evals_total.item()) == tot and (acc == S).all()
    e = tr.z_init.cpu().numpy().ravel()
    u = tr.u.cpu().numpy().ravel()
    n = e.size
    assert abs(e.mean() - 1) < 5 / np.sqrt(n) and abs(e.var() - 1) < 5 * np.sqrt(8 / n)        # Exp(1)
    assert abs(u.mean() - 0.5) < 5 / np.sqrt(12 * n) and 0 < u.min() and u.max() < 1


def test_slice_class_posterior_and_reference_tape():
    """Posterior parity: funnel without adaptation against the long tape of the unmodified reference Slice
    (means and the x variance within 4 standard errors, same mean number of value calls per draw), and
    ill-normal with adaptation against the analytic truth."""
    t, meta, data = load_tape("stats_slice_funnel_d2_noadapt")
    model = kb.BSModel(stan_file="stan/funnel.stan", data=data, device=device())
    s = kb.Slice(model, seed=5, chains=8192, warmup=0)
    s.run(1500)
    e0 = s.grad_evals
    S = 1000
    s1, s2 = s.run(S, chain_stats=True)
    assert s.acceptance_probability == 1.0
    evals_per_draw = (s.grad_evals - e0) / (S * s.chains)
    assert abs(evals_per_draw - float(t["evals_mean"])) < 0.15, (evals_per_draw, float(t["evals_mean"]))
    summ = chain_summary(s1, s2, S)
    th = t["theta_thin10"][300:]
    nb = 30
    bse = lambda x: x[:(len(x) // nb) * nb].reshape(nb, -1, *x.shape[1:]).mean(1).std(0, ddof=1) / np.sqrt(nb)
    m_ref, v_ref = th.mean(0), th.var(0, ddof=1)
    dm = np.abs(summ["mean"].cpu().numpy() - m_ref) / np.sqrt(bse(th) ** 2 + summ["mcse_mean"].cpu().numpy() ** 2)
    dv = np.abs(summ["var"].cpu().numpy() - v_ref) / np.sqrt(bse((th - m_ref) ** 2) ** 2 + summ["mcse_var"].cpu().numpy() ** 2)
    assert dm.max() <= 4 and dv[0] <= 4, (dm, dv)
    assert abs(float(summ["var"][0]) - 9.0) < 0.5                       # x ~ N(0, 3^2), experiment_funnel.py:68

    D = 16
    im = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": D}, device=device())
    truth = np.arange(1, D + 1) ** 2 / D
    a = kb.Slice(im, seed=2, chains=4096, warmup=600)
    a.run(600)
    assert np.allclose(a._cov, truth, rtol=0.25)
    s1, s2 = a.run(1500, chain_stats=True)
    summ = chain_summary(s1, s2, 1500)
    z = np.abs(summ["mean"].cpu().numpy()) / summ["mcse_mean"].cpu().numpy()
    zv = np.abs(summ["var"].cpu().numpy() - truth) / summ["mcse_var"].cpu().numpy()
    assert z.max() <= 4.5 and zv.max() <= 4.5, (z, zv)


def test_slice_class_api_checkpoint_and_single_chain(tmp_path):
    model = kb.BSModel(stan_file="stan/normal.stan", data={"D": 2}, device=device())
    one = kb.Slice(model, seed=3, warmup=100)                      # chains = 1: NumPy in / out like the reference
    draws = one.sample(300)
    assert isinstance(draws, np.ndarray) and draws.shape == (300, 2) and one.D == 2
    assert one.J == 2 and one._eigvecs.shape == (2, 3)             # J is not clipped (slice.py:41,58-59)
    assert one.acceptance_probability == 1.0
    assert np.abs(np.diff(draws, axis=0)).sum(1).min() > 0        # every draw moves
    with pytest.raises(NotImplementedError):
        kb.Slice(model, m=5)
    with pytest.raises(NotImplementedError):
        one.fit(np.ones(2))
    im = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": 30}, device=device())
    mk = lambda: kb.Slice(im, seed=12, chains=1000, warmup=200, windowsize=50)
    ref = mk()
    ref.run(330)
    a = mk()
    a.run(120)
    torch.save(a.state_dict(), tmp_path / "ckpt.pt")
    b = mk()
    b.load_state_dict(torch.load(tmp_path / "ckpt.pt", weights_only=False))
    b.run(210)
    assert torch.equal(b.theta, ref.theta) and np.array_equal(b._cov, ref._cov) and b.grad_evals == ref.grad_evals
    # thinned sample() rows are the chain states
    c = mk()
    out = c.sample(11, thin=3)
    d = mk()
    d.run(30)
    assert torch.equal(out[-1], d.theta)
