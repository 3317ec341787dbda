"""Single-chain CPU port of the reference KLHR step -- TEST INFRASTRUCTURE.

This is the reference algorithm with its optimiser left in place: stage 1 and stage 2 of
the line fit call ``scipy.optimize.minimize(method="BFGS")`` exactly as reference
``klhr.py:126-141`` / ``klhr_sinh.py:182-201`` do.  Arithmetic follows the cited lines
operation by operation so that, fed the same variates, it reproduces the tapes of the
unmodified reference in ``tests/golden`` BIT FOR BIT (tests/test_oracle_port.py).  It is
used (a) to pin the oracle and (b) as the CPU baseline ``bench.py`` times on the GPU
box's host cores, one process per chain like reference ``run_experiments:27``
(``cpu_baseline.kind == "port"``: /root/reference does not exist on that box).

Structure differs from the reference on purpose: one driver (``ChainSampler``) and a
family object (``GaussLine`` / ``SinhLine``) instead of two parallel classes.
"""
from __future__ import annotations

import numpy as np
import scipy.special as sp
from numpy.polynomial.hermite import hermgauss
from scipy.optimize import minimize

from .adapt import RunningMoments, StreamingPCA, WindowSchedule


def gauss_hermite_probabilists(N):
    """Nodes/weights with sum w = 1 for E_{N(0,1)}[.]  (reference ``klhr.py:46-49``)."""
    x, w = hermgauss(N)
    x *= np.sqrt(2)
    w /= np.sqrt(np.pi)
    return x, w


class GaussLine:
    """q = N(m, s^2) along the line; eta = (m, log s).  reference ``klhr.py:81-85,106-120,155-158``."""
    n_eta = 2
    stage2_options = None

    def __init__(self, x, w, tol, scale_clip):
        self.x, self.w, self.tol, self.clip = x, w, tol, scale_clip

    def unpack(self, eta):
        return eta[0], np.exp(np.clip(eta[1], -self.clip, self.clip))

    def kl(self, eta, theta, rho, model):
        m, s = self.unpack(eta)
        acc = 0.0
        g = np.zeros(2)
        for xn, wn in zip(self.x, self.w):
            y = s * xn + m
            lp, glp = model.log_density_gradient(y * rho + theta)
            acc += wn * lp
            wd = wn * glp.dot(rho)
            g[0] += wd
            g[1] += wd * xn * s
        acc += eta[1]
        g[1] += 1
        return -acc, -g

    def stage2_start(self, xi_hat, half_log_s2, rng, initscale):
        return np.array([xi_hat, half_log_s2])

    def propose(self, eta, z):
        m, s = self.unpack(eta)
        return m + s * z

    def logq(self, x, eta):
        m, s = self.unpack(eta)
        z = (x - m) / s
        return -np.log(s) - 0.5 * z * z

    def model_grad(self, model, theta):
        return model.log_density_gradient(theta)


class SinhLine:
    """Sinh-arcsinh family; eta = (m, log s, log d, e).
    reference ``klhr_sinh.py:78-84,100-176,233-240``."""
    n_eta = 4
    stage2_options = {"gtol": 1e-3}          # klhr_sinh.py:199

    def __init__(self, x, w, tol, scale_clip):
        self.x, self.w, self.tol, self.clip = x, w, tol, scale_clip

    def unpack(self, eta):
        c = self.clip
        s = np.exp(np.clip(eta[1], -c, c)) + self.tol
        d = np.exp(np.clip(eta[2], -c, c)) + self.tol
        return eta[0], s, d, eta[3]

    def _cl(self, a):
        return np.clip(a, -self.clip, self.clip)

    def transport(self, x, eta):                       # _T, klhr_sinh.py:112-114
        m, s, d, e = self.unpack(eta)
        return m + s * np.sinh(self._cl((np.arcsinh(x) + e) / d))

    def transport_inv(self, x, eta):                   # _T_inv, :126-129
        m, s, d, e = self.unpack(eta)
        z = (x - m) / s
        return np.sinh(self._cl(d * np.arcsinh(z) - e))

    def grad_transport(self, x, eta):                  # _grad_T, :116-124
        m, s, d, e = self.unpack(eta)
        g = np.ones(4)
        invd = 1 / d
        a = (np.arcsinh(x) + e) * invd
        g[1] = s * np.sinh(self._cl(a))
        g[2] = -s * np.cosh(self._cl(a)) * a
        g[3] = s * np.cosh(self._cl(a)) * invd
        return g

    def log_abs_jac(self, x, eta):                     # :139-144
        _, _, d, e = self.unpack(eta)
        out = eta[2] - eta[1]
        a = (np.arcsinh(x) + e) / d
        out -= np.log(np.cosh(self._cl(a)))
        return out

    def grad_log_abs_jac(self, x, eta):                # :146-156
        m, s, d, e = self.unpack(eta)
        invd = 1 / d
        g = np.zeros(4)
        g[1] = -1
        a = (np.arcsinh(x) + e) * invd
        t = np.tanh(self._cl(a))
        g[2] = 1 + t * a
        g[3] = -t * invd
        return g

    def model_grad(self, model, theta):                # _logp_grad, :158-161
        lp, g = model.log_density_gradient(theta)
        return lp, np.clip(g, -self.clip, self.clip)

    def kl(self, eta, theta, rho, model):              # :163-176
        acc = 0.0
        g = np.zeros(4)
        for xn, wn in zip(self.x, self.w):
            t = self.transport(xn, eta)
            lp, glp = self.model_grad(model, t * rho + theta)
            laj = self.log_abs_jac(xn, eta)
            acc += wn * (laj - lp)
            glaj = self.grad_log_abs_jac(xn, eta)
            gT = self.grad_transport(xn, eta)
            g -= wn * glp.dot(rho) * gT
            g += wn * glaj
        return acc, g

    def stage2_start(self, xi_hat, half_log_s2, rng, initscale):   # :191-193
        init = rng.normal(size=4) * initscale
        init[0] = xi_hat
        init[1] = half_log_s2
        return init

    def propose(self, eta, z):                          # :246
        return self.transport(z, eta)

    def logq(self, x, eta):                             # _log_q, :233-240
        m, s, d, e = self.unpack(eta)
        ld = -0.5 * self.transport_inv(x, eta) ** 2
        z = (x - m) / s
        ld += np.log(np.cosh(self._cl(d * np.arcsinh(z) - e)))
        ld += eta[2] - eta[1]
        ld -= 0.5 * np.log1p(z * z)
        return ld


class SubSinhLine(SinhLine):
    """3-parameter sinh-arcsinh family (d = 1); eta = (m, log s, e).
    reference ``sub_klhr_sinh.py:92-169,226-233`` -- same maths as SinhLine with d removed; the
    model gradient is clipped with ``grad_clip`` here (:152-154)."""
    n_eta = 3

    def __init__(self, x, w, tol, scale_clip, grad_clip=1e15):
        super().__init__(x, w, tol, scale_clip)
        self.grad_clip = grad_clip

    def unpack(self, eta):
        return eta[0], np.exp(np.clip(eta[1], -self.clip, self.clip)) + self.tol, eta[2]

    def transport(self, x, eta):
        m, s, e = self.unpack(eta)
        return m + s * np.sinh(self._cl(np.arcsinh(x) + e))

    def transport_inv(self, x, eta):
        m, s, e = self.unpack(eta)
        return np.sinh(self._cl(np.arcsinh((x - m) / s) - e))

    def grad_transport(self, x, eta):
        m, s, e = self.unpack(eta)
        g = np.ones(3)
        a = np.arcsinh(x) + e
        g[1] = s * np.sinh(self._cl(a))
        g[2] = s * np.cosh(self._cl(a))
        return g

    def log_abs_jac(self, x, eta):
        _, _, e = self.unpack(eta)
        out = -eta[1]
        out -= np.log(np.cosh(self._cl(np.arcsinh(x) + e)))
        return out

    def grad_log_abs_jac(self, x, eta):
        _, _, e = self.unpack(eta)
        g = np.zeros(3)
        g[1] = -1
        g[2] = -np.tanh(self._cl(np.arcsinh(x) + e))
        return g

    def model_grad(self, model, theta):
        lp, g = model.log_density_gradient(theta)
        return lp, np.clip(g, -self.grad_clip, self.grad_clip)

    def kl(self, eta, theta, rho, model):
        acc = 0.0
        g = np.zeros(3)
        for xn, wn in zip(self.x, self.w):
            t = self.transport(xn, eta)
            lp, glp = self.model_grad(model, t * rho + theta)
            laj = self.log_abs_jac(xn, eta)
            acc += wn * (laj - lp)
            glaj = self.grad_log_abs_jac(xn, eta)
            gT = self.grad_transport(xn, eta)
            g -= wn * glp.dot(rho) * gT
            g += wn * glaj
        return acc, g

    def stage2_start(self, xi_hat, half_log_s2, rng, initscale):
        init = rng.normal(size=3) * initscale
        init[0] = xi_hat
        init[1] = half_log_s2
        return init

    def logq(self, x, eta):
        m, s, e = self.unpack(eta)
        ld = -0.5 * self.transport_inv(x, eta) ** 2
        z = (x - m) / s
        ld += np.log(np.cosh(self._cl(np.arcsinh(z) - e)))
        ld -= eta[1]
        ld -= 0.5 * np.log1p(z * z)
        return ld


class ChainSampler:
    """One chain of KLHR (family="gauss") or KLHRSINH (family="sinh").

    Keyword names and defaults are the reference's (``klhr.py:16-34``,
    ``klhr_sinh.py:15-32``) except ``overrelaxed`` (False for both families here).  Over-relaxed
    proposals draw r and v from SciPy's global RNG like the reference; the Smoother-adapted K
    (``klhr.py:212-214``) is not ported: K stays at its constructor value.
    """

    clip_J = True

    def __init__(self, model, family="gauss", theta=None, seed=None, rng=None, N=8, K=10, J=2,
                 l=4, initscale=0.1, warmup=1_000, windowsize=50, windowscale=2, tol=None,
                 grad_clip=1e15, scale_clip=None, scale_dir_cov=False, overrelaxed=False,
                 eigen_method_one=None, max_init_tries=100):
        gauss = family == "gauss"
        self.overrelaxed = overrelaxed
        self.K = K
        tol = (1e-12 if gauss else 1e-10) if tol is None else tol
        scale_clip = (600 if gauss else 300) if scale_clip is None else scale_clip
        eigen_method_one = gauss if eigen_method_one is None else eigen_method_one
        self.model = model
        self.D = model.dim()
        self.rng = rng if rng is not None else np.random.default_rng(seed)
        self.family_name = family
        self.N = N
        self.J = (J if J < self.D else self.D - 1) if (gauss and self.clip_J) else J   # klhr.py:39 vs klhr_sinh.py:37, slice.py:41
        self.tol, self.initscale = tol, initscale
        self.x, self.w = gauss_hermite_probabilists(N)
        self.line = {"gauss": GaussLine, "sinh": SinhLine, "subsinh": SubSinhLine}[family](
            self.x, self.w, tol, scale_clip)
        self.schedule = WindowSchedule(warmup, windowsize, windowscale)
        self.mom_theta = RunningMoments(self.D)
        self.mom_grad = RunningMoments(self.D)
        self.pca = StreamingPCA(self.D, K=self.J, l=l)
        self.dir_mean = np.zeros(self.D)
        self.dir_cov = np.ones(self.D)
        self.scale_dir_cov = scale_dir_cov
        self.method_one = eigen_method_one
        ncol = self.J + 1 if eigen_method_one else self.J
        self.eigvecs = np.zeros((self.D, ncol))
        self.eigvals = np.ones(ncol)
        self.n_draw = 0
        self.acceptance_probability = 0
        self.grad_evals = 0
        if theta is not None:
            self.theta = np.array(theta, dtype=np.float64)
        else:
            self.rng.normal(scale=0.1, size=self.D)                      # mcmc.py:14-17 (then overwritten)
            for _ in range(max_init_tries):                              # klhr.py:87-99
                cand = self.rng.normal(size=self.D) * initscale
                lp, g = model.log_density_gradient(cand)
                if np.isfinite(lp) and np.isfinite(np.linalg.norm(g)):
                    self.theta = cand
                    break
            else:
                raise RuntimeError("failed to initialize")

    # ------------------------------------------------------------------ direction (H3)
    def random_direction(self):
        lam = self.eigvals
        p = lam / np.sum(lam)
        if self.method_one:
            j = self.rng.choice(np.size(p), p=p)
            mvec = self.eigvecs[:, j]
        elif self.family_name == "gauss":
            mvec = np.sum(lam * self.eigvecs, axis=1)                    # klhr.py:150
        else:
            mvec = np.sum(p * self.eigvecs, axis=1)                      # klhr_sinh.py:210
        x = self.rng.multivariate_normal(mvec, np.diag(self.dir_cov))
        return x / np.linalg.norm(x + self.tol)

    # ------------------------------------------------------------------------ fit (H5)
    def _neg_line(self, xi, rho):
        lp, g = self.model.log_density_gradient(xi * rho + self.theta)
        return -lp, -g.dot(rho)

    def fit(self, rho):
        o1 = minimize(self._neg_line, self.rng.normal() * self.initscale, args=(rho,),
                      jac=True, method="BFGS")
        self.grad_evals += o1["nfev"]
        s2 = o1["hess_inv"][0, 0]
        start = self.line.stage2_start(o1.x[0], (s2 > 0) * 0.5 * np.log(s2), self.rng,
                                       self.initscale)
        kw = {}
        if self.line.stage2_options:
            kw["options"] = self.line.stage2_options
        o2 = minimize(lambda eta, r: self.line.kl(eta, self.theta, r, self.model), start,
                      args=(rho,), jac=True, method="BFGS", **kw)
        self.grad_evals += o2["nfev"] * self.N
        return o2.x

    # ------------------------------------------------------------------------- MH (H8)
    def overrelaxed_proposal(self, eta):
        """reference ``klhr.py:160-173`` / ``klhr_sinh.py:215-228``; r and v come from SciPy's
        GLOBAL RandomState exactly like the reference (seed it with ``np.random.seed``)."""
        import scipy.stats as st
        K = self.K
        if self.family_name == "gauss":
            m, s = self.line.unpack(eta)
            dist = st.norm(m, s)
            u = dist.cdf(np.array([0]))
            up = u
        else:
            u = sp.ndtr(self.line.transport_inv(np.zeros(1), eta))
            up = 0
        r = st.binom(K, u).rvs()
        if r > K - r:
            v = st.beta(K - r + 1, 2 * r - K).rvs()
            up = u * v
        elif r < K - r:
            v = st.beta(r + 1, K - 2 * r).rvs()
            up = 1 - (1 - u) * v
        elif self.family_name != "gauss":
            up = u
        if self.family_name == "gauss":
            return dist.ppf(up)
        return self.line.transport(sp.ndtri(up), eta)

    def metropolis(self, eta, rho):
        if self.overrelaxed:
            zp = self.overrelaxed_proposal(eta)
        else:
            z = self.rng.normal(size=1)
            zp = self.line.propose(eta, z)
        cand = zp * rho + self.theta
        r = self.model.log_density(cand)
        r -= self.model.log_density(self.theta)
        r += self.line.logq(0, eta)
        r -= self.line.logq(zp, eta)
        a = np.log(self.rng.uniform()) < np.minimum(0, r)
        self.theta = a * cand + (1 - a) * self.theta
        self.acceptance_probability += (a - self.acceptance_probability) / self.n_draw
        self.last = dict(zp=float(np.ravel(zp)[0]), r=float(np.ravel(r)[0]),
                         accept=bool(np.ravel(a)[0]))
        return self.theta

    def transition(self, rho):
        eta = self.fit(rho)
        theta = self.metropolis(eta, rho)
        self.last.update(rho=rho, eta=eta)
        return theta

    # ------------------------------------------------------------------------ draw (H9)
    def draw(self):
        self.n_draw += 1
        rho = self.random_direction()
        theta = self.transition(rho)
        if self.schedule.window_closed(self.n_draw):
            self.dir_mean = self.mom_theta.mean()
            self.dir_cov = self.mom_theta.var()
            if self.scale_dir_cov:
                self.dir_cov /= (self.tol + self.mom_grad.var())
            self.mom_grad.reset()
            self.mom_theta.reset()
            self.eigvecs[:, :self.J] = self.pca.vectors()
            self.eigvals[:self.J] = self.pca.values()
            self.pca.reset()
        else:
            _, g = self.line.model_grad(self.model, theta)
            self.mom_grad.update(g)
            self.mom_theta.update(theta)
            self.pca.update(theta - self.dir_mean)
        return theta

    def sample(self, M):                                                 # mcmc.py:31-37
        out = np.empty((M, self.D))
        out[0] = self.theta
        for m in range(1, M):
            out[m] = self.draw()
        return out


class SliceChain(ChainSampler):
    """Univariate slice sampling (stepping out + shrinkage, Neal 2003) along KLHR's adapted random
    directions: port of reference ``slice.py:12-167`` for its working configuration ``m = inf``
    (the finite-``m`` branch of the reference raises NameError, slice.py:108,124).  Direction law and
    windowed adaptation are the KLHR ones (slice.py:148-176 == klhr.py:143-153,196-223) except that J
    is not clipped to D - 1 (slice.py:41).  Generator call order per draw: [choice], multivariate_normal,
    exponential, uniform(0, w), then one uniform(L, R) per shrinkage proposal."""
    clip_J = False

    def __init__(self, model, w=1, lower=-np.inf, upper=np.inf, tol=1e-12, **kw):
        super().__init__(model, family="gauss", tol=tol, **kw)
        self.w, self.lower, self.upper = w, lower, upper
        self.last = {}

    def transition(self, rho):                                           # _uni_slice, slice.py:84-146
        g = lambda x: self.model.log_density(rho * x + self.theta)
        n0 = getattr(self.model, "n_value_calls", 0)
        logy = g(0.0) - self.rng.exponential()
        u = self.rng.uniform(low=0, high=self.w)
        L = 0.0 - u
        R = 0.0 + (self.w - u)
        while True:                                                      # stepping out, m = inf (:95-107)
            if L <= self.lower:
                break
            if g(L) <= logy:
                break
            L -= self.w
        while True:
            if R >= self.upper:
                break
            if g(R) <= logy:
                break
            R += self.w
        L = max(L, self.lower)                                           # :127-128
        R = min(R, self.upper)
        n_shrink = 0
        while True:                                                      # shrinkage (:131-139)
            x1 = self.rng.uniform(low=L, high=R)
            n_shrink += 1
            if g(x1) >= logy:
                break
            if x1 > 0.0:
                R = x1
            else:
                L = x1
        self.theta = rho * x1 + self.theta
        self.acceptance_probability += (1 - self.acceptance_probability) / self.n_draw
        self.last = dict(rho=rho, x1=x1, L=L, R=R, n_shrink=n_shrink,
                         evals=getattr(self.model, "n_value_calls", 0) - n0)
        return self.theta


class MHChain:
    """Random-walk Metropolis, one chain: port of reference ``mh.py:7-37`` (``stepsize * N(0, I)``
    proposal, symmetric-proposal terms kept like the reference so the arithmetic is identical)."""

    def __init__(self, model, stepsize, seed=None, theta=None):
        self.model, self.D, self.stepsize = model, model.dim(), stepsize
        self.rng = np.random.default_rng(seed)
        self.theta = self.rng.normal(scale=0.1, size=self.D) if theta is None else np.array(theta, dtype=np.float64)
        self.n_draw = 0
        self.acceptance_probability = 0

    def _logq(self, a, b):
        z = (a - b) / self.stepsize
        return -0.5 * z.dot(z)

    def draw(self):
        self.n_draw += 1
        xi = self.rng.normal(size=self.D)
        cand = self.theta + xi * self.stepsize
        r = self.model.log_density(cand)
        r -= self.model.log_density(self.theta)
        r += self._logq(self.theta, cand)
        r -= self._logq(cand, self.theta)
        a = np.log(self.rng.uniform()) < np.minimum(0.0, r)
        self.theta = a * cand + (1 - a) * self.theta
        self.acceptance_probability += (a - self.acceptance_probability) / self.n_draw
        return self.theta


class ReplayRNG:
    """Plays back a golden tape's variates through the Generator calls the sampler makes."""

    def __init__(self, tape, family):
        self.t, self.family, self.i = tape, family, 0
        self._norm_calls = 0

    def choice(self, n, p=None):
        u = float(self.t["ujdir"][self.i])
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, u, side="right"))

    def multivariate_normal(self, mean, cov):
        # replay injects the direction itself: hand back a vector
        # whose normalisation reproduces rho to the last bit is impossible in general, so
        # the driver below bypasses this method.
        raise RuntimeError("ReplayRNG: directions are injected, not drawn")

    def normal(self, loc=0.0, scale=1.0, size=None):
        k = self._norm_calls
        self._norm_calls += 1
        if self.family == "gauss":
            z = self.t["z_init"][self.i] if k % 2 == 0 else np.array([self.t["z_prop"][self.i]])
        else:
            start = self.t["init4"][self.i] if self.family == "sinh" else self.t["init4"][self.i][:3]
            z = (self.t["z_init"][self.i], start, np.array([self.t["z_prop"][self.i]]))[k % 3]
        return loc + scale * z

    def uniform(self):
        u = float(self.t["u"][self.i])
        self.i += 1
        return u


def replay_port(tape, model, family, n=None, **kw):
    """Re-run the port over a golden tape with injected rho/variates; returns per-draw
    (eta, zp, r, accept, theta_after) for comparison with the tape."""
    M = len(tape["u"]) if n is None else n
    rng = ReplayRNG(tape, family)
    s = ChainSampler(model, family=family, theta=tape["theta0"][0], rng=rng, **kw)
    s.random_direction = lambda: tape["rho"][rng.i]
    out = dict(eta=[], zp=[], r=[], accept=[], theta=[])
    for _ in range(M):
        th = s.draw()
        out["eta"].append(np.array(s.last["eta"]))
        out["zp"].append(s.last["zp"])
        out["r"].append(s.last["r"])
        out["accept"].append(s.last["accept"])
        out["theta"].append(np.array(th))
    out = {k: np.array(v) for k, v in out.items()}
    out["sampler"] = s
    return out


class SliceReplayRNG:
    """Plays a slice tape (tests/golden/slice_*.npz) back through the Generator calls ``SliceChain`` makes."""

    def __init__(self, tape):
        self.t, self.i, self.k = tape, 0, 0

    def choice(self, n, p=None):
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return int(np.searchsorted(cdf, float(self.t["ujdir"][self.i]), side="right"))

    def exponential(self):
        return float(self.t["e"][self.i])

    def uniform(self, low=0.0, high=1.0):
        if self.k == 0:
            u = float(self.t["u0"][self.i])
        else:
            u = float(self.t["shrink_u"][self.i][self.k - 1])
        self.k += 1
        return low + (high - low) * u


def replay_slice_port(tape, model, n=None, **kw):
    """Re-run ``SliceChain`` over a slice tape with injected rho / variates."""
    M = len(tape["e"]) if n is None else n
    rng = SliceReplayRNG(tape)
    s = SliceChain(model, theta=tape["theta0"][0], rng=rng, **kw)
    s.random_direction = lambda: tape["rho"][rng.i]
    out = dict(x1=[], n_shrink=[], evals=[], theta=[])
    for i in range(M):
        rng.i, rng.k = i, 0
        th = s.draw()
        out["x1"].append(s.last["x1"])
        out["n_shrink"].append(s.last["n_shrink"])
        out["evals"].append(s.last["evals"])
        out["theta"].append(np.array(th))
    out = {k: np.array(v) for k, v in out.items()}
    out["sampler"] = s
    return out
