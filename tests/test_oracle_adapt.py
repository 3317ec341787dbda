import numpy as np

from oracle import adapt


def test_window_schedule_known_values():
    # SURVEY.md 8a H16 [probe] values of reference windowedadaptation.py
    assert adapt.window_closures(1000, 50, 2) == [50, 150, 350, 1000]
    assert adapt.window_closures(100, 50, 2) == [50, 100]
    assert adapt.window_closures(15000, 50, 2) == [50, 150, 350, 750, 1550, 3150, 6350, 15000]
    assert adapt.window_closures(40, 50, 2) == []
    s = adapt.WindowSchedule(100, 50, 2)
    assert [m for m in range(1, 300) if s.window_closed(m)] == [50, 100]


def test_pooled_moments_equal_welford():
    rng = np.random.default_rng(1)
    X = rng.normal(size=(500, 4)) * [1, 2, 3, 4] + 7
    rm = adapt.RunningMoments(4)
    for x in X:
        rm.update(x)
    shift = X[0]
    mu, var = adapt.pooled_moments(len(X), (X - shift).sum(0), ((X - shift) ** 2).sum(0), shift)
    assert np.allclose(mu, rm.mean(), rtol=1e-12) and np.allclose(var, rm.var(), rtol=1e-10)
    assert np.allclose(var, X.var(0, ddof=1))
    assert np.all(adapt.pooled_moments(2, X[:2].sum(0), (X[:2] ** 2).sum(0))[1] == 1)


def test_pooled_pca_matches_covariance_eig_and_ccipca_direction():
    rng = np.random.default_rng(2)
    D, n = 20, 4000
    A = rng.normal(size=(D, D))
    X = rng.normal(size=(n, D)) @ A.T
    V, lam = adapt.pooled_pca(n, X.sum(0), X.T @ X, J=2)
    w, U = np.linalg.eigh(np.cov(X.T))
    assert np.allclose(lam, w[::-1][:2] + 1e-10, rtol=1e-9)
    assert np.allclose(np.abs(V.T @ U[:, ::-1][:, :2]), np.eye(2), atol=1e-8)
    pca = adapt.StreamingPCA(D, K=2, l=0)
    for x in X - X.mean(0):
        pca.update(x)
    # CCIPCA is order dependent; it converges towards the same leading direction
    assert abs(pca.vectors()[:, 0] @ V[:, 0]) > 0.95
