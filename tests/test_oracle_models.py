"""Oracle model layer: finite-difference and closed-form identities (SURVEY.md 8c:
BridgeStan is absent, so the .stan restatements are self-validated)."""
import numpy as np
import pytest

from oracle import stan_models as sm


def _models():
    rng = np.random.default_rng(5)
    y = sm.simulate_ark_series(T=120, seed=3)
    h = 66.0 + 4.0 * rng.normal(size=80)
    e = 3.0 + 0.5 * (h - 66.0) + 2.0 * rng.normal(size=80)
    return [sm.Normal(3), sm.IllNormal(7), sm.Funnel(1), sm.Funnel(4), sm.CorrNormal(6, 0.9),
            sm.AR1(8), sm.ARK(5, 120, y), sm.Rosenbrock(2), sm.Earnings(80, e, h)], rng


@pytest.mark.parametrize("idx", range(9))
def test_gradient_and_dir2_match_finite_differences(idx):
    models, rng = _models()
    m = models[idx]
    D = m.dim()
    for _ in range(5):
        th = rng.normal(size=D) * 0.7
        rho = rng.normal(size=D)
        rho /= np.linalg.norm(rho)
        lp, g = m.lp_grad(th)
        h = 1e-6
        fd = np.array([(m.lp(th + h * e) - m.lp(th - h * e)) / (2 * h) for e in np.eye(D)])
        assert np.allclose(g, fd, rtol=1e-6, atol=1e-6 * (1 + abs(lp)))
        d1p = m.lp_grad(th + h * rho)[1] @ rho
        d1m = m.lp_grad(th - h * rho)[1] @ rho
        assert np.isclose(m.dir2(th, rho), (d1p - d1m) / (2 * h), rtol=1e-5, atol=1e-5)


def test_batched_shapes():
    models, rng = _models()
    for m in models:
        th = rng.normal(size=(3, 4, m.dim()))
        lp, g = m.lp_grad(th)
        assert lp.shape == (3, 4) and g.shape == th.shape
        assert m.dir2(th, th).shape == (3, 4)
        assert np.allclose(lp[1, 2], m.lp(th[1, 2]))


def test_corr_normal_precision_is_tridiagonal():
    # SURVEY.md 8a M4 validation identity
    m = sm.CorrNormal(9, 0.9)
    r = 0.9
    T = np.zeros((9, 9))
    np.fill_diagonal(T, 1 + r * r)
    T[0, 0] = T[-1, -1] = 1
    for i in range(8):
        T[i, i + 1] = T[i + 1, i] = -r
    assert np.allclose(m.P, T / (1 - r * r), atol=1e-10)


def test_ar1_equals_corr_normal():
    # a stationary AR(1) with alpha = 0.9 has covariance 0.9^|i-j|
    a, c = sm.AR1(12), sm.CorrNormal(12, 0.9)
    th = np.random.default_rng(0).normal(size=(5, 12))
    assert np.allclose(a.lp(th), c.lp(th), atol=1e-10)


def test_ill_normal_scales():
    m = sm.IllNormal(100)
    assert np.isclose(1 / m.inv_s2[0], 1 / 100) and np.isclose(1 / m.inv_s2[-1], 100.0)


def test_ark_against_direct_loop():
    y = sm.simulate_ark_series(T=60, seed=1)
    m = sm.ARK(5, 60, y)
    th = np.array([0.1, 0.05, -0.1, 0.15, -0.2, 0.6, np.log(0.5)])
    a, b, sig = th[0], th[1:6], np.exp(th[6])
    lp = -0.5 * a * a - 0.5 * b @ b - 0.5 * sig * sig + th[6]
    for t in range(5, 60):      # stan/arK.stan:15-17 (1-based t = K+1..T)
        mu = a + b @ y[t - 5:t]
        lp += -np.log(sig) - 0.5 * ((y[t] - mu) / sig) ** 2
    assert np.isclose(m.lp(th), lp, rtol=1e-12)


def test_shim_failure_semantics():
    from oracle.bsmodel import BSModel
    b = BSModel(stan_file="stan/funnel.stan", data={"D": 1})
    lp, g = b.log_density_gradient(np.array([-2000.0, 1.0]))   # exp overflow
    assert lp == -np.inf and np.all(g == 0)
    assert b.log_density(np.array([-2000.0, 1.0])) == -np.inf
    assert b.dim() == 2
