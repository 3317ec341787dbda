"""NumPy fp64 restatement of the Stan targets named by the north star.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each class follows one Stan program of
the reference (cited per class) under BridgeStan's defaults ``propto=True,
jacobian=True`` (reference ``bsmodel.py:18,27`` forwards no kwargs), i.e. additive
constants that do not depend on parameters are dropped and constrained parameters are
evaluated on the unconstrained scale with the log-Jacobian added.

All methods are batched: ``theta`` has shape ``(..., D)``.

    lp(theta)            -> (...)
    lp_grad(theta)       -> (...), (..., D)
    dir2(theta, rho)     -> (...)    rho^T Hessian(theta) rho   (second directional
                                     derivative; used by the Newton line fit)

BridgeStan is a third-party dependency absent from /root/reference (un-vendored,
un-pinned): parity of this layer is anchored on the .stan sources plus
finite-difference and closed-form identities (tests/test_oracle_models.py).
"""
from __future__ import annotations

import numpy as np


class StanModelBase:
    name = "?"

    def dim(self) -> int:
        raise NotImplementedError

    def lp(self, theta):
        return self.lp_grad(theta)[0]

    def lp_grad(self, theta):
        raise NotImplementedError

    def dir2(self, theta, rho):
        raise NotImplementedError


class Normal(StanModelBase):
    """reference stan/normal.stan:1-9  --  y ~ normal(0, 1)."""
    name = "normal"

    def __init__(self, D):
        self.D = int(D)

    def dim(self):
        return self.D

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        return -0.5 * np.sum(theta * theta, axis=-1), -theta

    def dir2(self, theta, rho):
        return -np.sum(rho * rho, axis=-1)


class IllNormal(StanModelBase):
    """reference stan/ill-normal.stan:1-12  --  s = linspaced(1..D)/sqrt(D); y ~ normal(0, s)."""
    name = "ill-normal"

    def __init__(self, D):
        self.D = int(D)
        s = np.arange(1, self.D + 1, dtype=np.float64) / np.sqrt(float(self.D))
        self.inv_s2 = 1.0 / (s * s)

    def dim(self):
        return self.D

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        g = -theta * self.inv_s2
        return 0.5 * np.sum(theta * g, axis=-1), g

    def dir2(self, theta, rho):
        return -np.sum(rho * rho * self.inv_s2, axis=-1)


class Funnel(StanModelBase):
    """reference stan/funnel.stan:1-11  --  x ~ normal(0,3); alpha ~ normal(0, exp(x/2)).

    Parameters [x, alpha_1..alpha_D]; dims = D + 1.
    lp = -x^2/18 - D x / 2 - exp(-x) sum(alpha^2) / 2.
    """
    name = "funnel"

    def __init__(self, D):
        self.Da = int(D)

    def dim(self):
        return self.Da + 1

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        x = theta[..., 0]
        a = theta[..., 1:]
        with np.errstate(over="ignore", invalid="ignore"):
            ex = np.exp(-x)
            ss = np.sum(a * a, axis=-1)
            lp = -x * x / 18.0 - 0.5 * self.Da * x - 0.5 * ex * ss
            g = np.empty_like(theta)
            g[..., 0] = -x / 9.0 - 0.5 * self.Da + 0.5 * ex * ss
            g[..., 1:] = -a * ex[..., None]
        return lp, g

    def dir2(self, theta, rho):
        x = theta[..., 0]
        a = theta[..., 1:]
        r0 = rho[..., 0]
        ra = rho[..., 1:]
        with np.errstate(over="ignore", invalid="ignore"):
            ex = np.exp(-x)
            ss = np.sum(a * a, axis=-1)
            sa = np.sum(a * ra, axis=-1)
            sr = np.sum(ra * ra, axis=-1)
            # d2/dy2 of -(x0+y r0)^2/18 - D(x0+y r0)/2 - 1/2 e^{-(x0+y r0)} |a + y ra|^2 at y=0
            return -r0 * r0 / 9.0 - 0.5 * ex * (r0 * r0 * ss - 4.0 * r0 * sa + 2.0 * sr)


class CorrNormal(StanModelBase):
    """reference stan/corr-normal.stan:1-20  --  y ~ multi_normal(0, Sigma), Sigma_ij = rho^|i-j|.

    lp = -1/2 y^T P y with the DENSE precision P = Sigma^{-1} (log-det dropped by propto).
    """
    name = "corr-normal"

    def __init__(self, N, rho):
        self.D = int(N)
        self.rho = float(rho)
        idx = np.arange(self.D)
        self.Sigma = self.rho ** np.abs(idx[:, None] - idx[None, :])
        P = np.linalg.inv(self.Sigma)
        self.P = 0.5 * (P + P.T)

    def dim(self):
        return self.D

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        g = -theta @ self.P
        return 0.5 * np.sum(theta * g, axis=-1), g

    def dir2(self, theta, rho):
        return -np.sum((rho @ self.P) * rho, axis=-1)


class AR1(StanModelBase):
    """reference stan/ar1.stan:1-14  --  y1 ~ N(0,1); y_t ~ N(0.9 y_{t-1}, sqrt(1-0.81))."""
    name = "ar1"

    def __init__(self, N, alpha=0.9):
        self.D = int(N)
        self.alpha = float(alpha)
        self.inv_b2 = 1.0 / (1.0 - self.alpha * self.alpha)

    def dim(self):
        return self.D

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        e = theta[..., 1:] - self.alpha * theta[..., :-1]
        lp = -0.5 * theta[..., 0] ** 2 - 0.5 * self.inv_b2 * np.sum(e * e, axis=-1)
        g = np.zeros_like(theta)
        g[..., 0] = -theta[..., 0]
        g[..., 1:] -= self.inv_b2 * e
        g[..., :-1] += self.inv_b2 * self.alpha * e
        return lp, g

    def dir2(self, theta, rho):
        e = rho[..., 1:] - self.alpha * rho[..., :-1]
        return -rho[..., 0] ** 2 - self.inv_b2 * np.sum(e * e, axis=-1)


class ARK(StanModelBase):
    """reference stan/arK.stan:1-20.

    Unconstrained parameters [alpha, beta_1..beta_K, u = log sigma]; dims = K + 2.
    lp = -a^2/2 - |b|^2/2 - e^{2u}/2 + u - (T-K) u - e^{-2u}/2 * sum_{t=K+1..T} r_t^2,
    r_t = y_t - a - sum_k b_k y_{t-K+k-1}    (beta_1 multiplies the oldest lag).
    """
    name = "arK"

    def __init__(self, K, T, y):
        self.K = int(K)
        self.T = int(T)
        y = np.asarray(y, dtype=np.float64)
        assert y.shape == (self.T,)
        self.y = y
        n = self.T - self.K
        X = np.empty((n, self.K + 1))
        X[:, 0] = 1.0
        for k in range(self.K):
            X[:, 1 + k] = y[k:k + n]
        self.X = X
        self.yt = y[self.K:]
        self.n = n

    def dim(self):
        return self.K + 2

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        phi = theta[..., :-1]
        u = theta[..., -1]
        r = self.yt - phi @ self.X.T                      # (..., n)
        ssr = np.sum(r * r, axis=-1)
        with np.errstate(over="ignore", invalid="ignore"):
            e2 = np.exp(2.0 * u)
            em2 = np.exp(-2.0 * u)
            lp = (-0.5 * np.sum(phi * phi, axis=-1) - 0.5 * e2 + u
                  - self.n * u - 0.5 * em2 * ssr)
            g = np.empty_like(theta)
            g[..., :-1] = -phi + em2[..., None] * (r @ self.X)
            g[..., -1] = -e2 + 1.0 - self.n + em2 * ssr
        return lp, g

    def dir2(self, theta, rho):
        phi = theta[..., :-1]
        u = theta[..., -1]
        rp = rho[..., :-1]
        ru = rho[..., -1]
        r = self.yt - phi @ self.X.T
        dr = -(rp @ self.X.T)
        q0 = np.sum(r * r, axis=-1)
        q1 = np.sum(r * dr, axis=-1)
        q2 = np.sum(dr * dr, axis=-1)
        with np.errstate(over="ignore", invalid="ignore"):
            e2 = np.exp(2.0 * u)
            em2 = np.exp(-2.0 * u)
            # S(y) = q0 + 2 q1 y + q2 y^2 ; term = -1/2 e^{-2(u + y ru)} S(y)
            return (-np.sum(rp * rp, axis=-1) - 2.0 * ru * ru * e2
                    - 0.5 * em2 * (4.0 * ru * ru * q0 - 8.0 * ru * q1 + 2.0 * q2))


class Rosenbrock(StanModelBase):
    """reference stan/rosenbrock.stan:1-12  --  v ~ N(1,1); theta ~ N(v^2, 0.1).

    Parameters [v_1..v_D, t_1..t_D]; dims = 2 D.  lp = -1/2 sum (v-1)^2 - 50 sum (t - v^2)^2.
    """
    name = "rosenbrock"

    def __init__(self, D):
        self.Dh = int(D)

    def dim(self):
        return 2 * self.Dh

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        v = theta[..., :self.Dh]
        t = theta[..., self.Dh:]
        c = t - v * v
        lp = -0.5 * np.sum((v - 1.0) ** 2, axis=-1) - 50.0 * np.sum(c * c, axis=-1)
        g = np.empty_like(theta)
        g[..., :self.Dh] = -(v - 1.0) + 200.0 * v * c
        g[..., self.Dh:] = -100.0 * c
        return lp, g

    def dir2(self, theta, rho):
        v = theta[..., :self.Dh]
        t = theta[..., self.Dh:]
        rv = rho[..., :self.Dh]
        rt = rho[..., self.Dh:]
        c0 = t - v * v
        c1 = rt - 2.0 * v * rv
        c2 = -rv * rv
        # c(y) = c0 + c1 y + c2 y^2 ; d2/dy2 c^2 at 0 = 2 c1^2 + 4 c0 c2
        return -np.sum(rv * rv, axis=-1) - 50.0 * np.sum(2.0 * c1 * c1 + 4.0 * c0 * c2, axis=-1)


class Earnings(StanModelBase):
    """reference stan/earnings.stan:1-17 -- the model of the reference's own self-tests
    (klhr.py:238) and relaxation experiment.  Unconstrained [b1, b2, us = log sigma, ut = log s]:
    lp = -0.01 e^ut + ut  - sum_k [ut + 3 log(1 + b_k^2 e^{-2ut} / 5)]  - 0.1 e^us + us
         - N us - e^{-2us}/2 * sum_i (earn_i - b1 - b2 height_i)^2."""
    name = "earnings"

    def __init__(self, N, earn, height):
        self.N = int(N)
        self.e = np.asarray(earn, dtype=np.float64)
        self.h = np.asarray(height, dtype=np.float64)
        assert self.e.shape == (self.N,) and self.h.shape == (self.N,)

    def dim(self):
        return 4

    def _resid(self, theta):
        return self.e - theta[..., 0:1] - theta[..., 1:2] * self.h          # (..., N)

    def lp_grad(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        b, us, ut = theta[..., :2], theta[..., 2], theta[..., 3]
        r = self._resid(theta)
        ssr = np.sum(r * r, axis=-1)
        with np.errstate(all="ignore"):
            q = b * b * np.exp(-2.0 * ut)[..., None]
            em2 = np.exp(-2.0 * us)
            lp = (-0.01 * np.exp(ut) + ut - 2.0 * ut - 3.0 * np.sum(np.log1p(q / 5.0), axis=-1)
                  - 0.1 * np.exp(us) + us - self.N * us - 0.5 * em2 * ssr)
            g = np.empty_like(theta)
            g[..., 0] = -6.0 * b[..., 0] * np.exp(-2.0 * ut) / (5.0 + q[..., 0]) + em2 * np.sum(r, axis=-1)
            g[..., 1] = -6.0 * b[..., 1] * np.exp(-2.0 * ut) / (5.0 + q[..., 1]) + em2 * np.sum(r * self.h, axis=-1)
            g[..., 2] = -0.1 * np.exp(us) + 1.0 - self.N + em2 * ssr
            g[..., 3] = -0.01 * np.exp(ut) - 1.0 + np.sum(6.0 * q / (5.0 + q), axis=-1)
        return lp, g

    def dir2(self, theta, rho):
        b, us, ut = theta[..., :2], theta[..., 2], theta[..., 3]
        rb, rs, rt = rho[..., :2], rho[..., 2], rho[..., 3]
        r = self._resid(theta)
        dr = -(rb[..., 0:1] + rb[..., 1:2] * self.h)
        S = np.sum(r * r, axis=-1)
        S1 = 2.0 * np.sum(r * dr, axis=-1)
        S2 = 2.0 * np.sum(dr * dr, axis=-1)
        with np.errstate(all="ignore"):
            emt = np.exp(-ut)[..., None]
            w = b * emt
            w1 = rb * emt - rt[..., None] * w
            w2 = -rt[..., None] * rb * emt - rt[..., None] * w1
            f1 = -6.0 * w / (5.0 + w * w)
            f2 = -6.0 * (5.0 - w * w) / (5.0 + w * w) ** 2
            em2 = np.exp(-2.0 * us)
            return (-0.01 * rt * rt * np.exp(ut) + np.sum(f2 * w1 * w1 + f1 * w2, axis=-1)
                    - 0.1 * rs * rs * np.exp(us) - 0.5 * em2 * (4.0 * rs * rs * S - 2.0 * 2.0 * rs * S1 + S2))


MODEL_NAMES = ("normal", "ill-normal", "funnel", "corr-normal", "ar1", "arK", "rosenbrock", "earnings")


def make_model(name: str, data: dict) -> StanModelBase:
    """Build a model from the Stan program's stem and its JSON data dict."""
    if name == "normal":
        return Normal(data["D"])
    if name == "ill-normal":
        return IllNormal(data["D"])
    if name == "funnel":
        return Funnel(data["D"])
    if name == "corr-normal":
        return CorrNormal(data["N"], data["rho"])
    if name == "ar1":
        return AR1(data["N"])
    if name == "arK":
        return ARK(data["K"], data["T"], data["y"])
    if name == "rosenbrock":
        return Rosenbrock(data["D"])
    if name == "earnings":
        return Earnings(data["N"], data["earn"], data["height"])
    raise ValueError(f"unknown Stan model {name!r}; known: {MODEL_NAMES}")


def simulate_ark_series(T=10_000, seed=20261018):
    """Synthetic AR(5) series of BASELINE.md config 5 (coefficients stated there)."""
    beta = np.array([0.05, -0.10, 0.15, -0.20, 0.60])
    rng = np.random.default_rng(seed)
    burn = 500
    eps = rng.standard_normal(T + burn)
    y = np.zeros(T + burn)
    for t in range(5, T + burn):
        y[t] = 0.2 + beta @ y[t - 5:t] + 0.5 * eps[t]
    return y[burn:]
