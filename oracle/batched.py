"""Batched CPU restatement of one KLHR step with the FIXED-ITERATION optimiser the CUDA
kernels use -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Everything except the two ``scipy.optimize.minimize`` calls follows the reference:

    direction     klhr.py:143-153 / klhr_sinh.py:203-213      -> ``direction_from_normals``
    KL objective  klhr.py:106-120 / klhr_sinh.py:163-176      -> ``_kl_gauss`` / ``_kl_sinh``
    family maths  klhr.py:81-85,155-158 / klhr_sinh.py:78-156,233-240
    MH step       klhr.py:175-194 / klhr_sinh.py:242-260      -> ``step``

``minimize`` (reference ``klhr.py:127-139``) is replaced by a damped Newton iteration with
a fixed iteration budget (``FitConfig``), the same one ``klhr_b200/csrc`` implements
operation for operation:

  stage 1  1-D mode search on l(xi) = lp(theta + xi rho) from xi0 = z_init*initscale:
           Newton step -l'/l'' (at most 64 trust radii) where l'' < 0, otherwise a trust-capped ascent step; the 8
           candidates xi + 2^-k step (k = 0..7) are scored and the best improving one kept.
  stage 2  Newton on the KL objective in scaled coordinates (m/s, log s[, log d, e]) with
           the exact Hessian (second directional derivative l'' of the target, analytic
           second derivatives of the sinh-arcsinh transport), Levenberg shift until
           Cholesky succeeds, step cap, Armijo back-tracking (halving).

The target enters only through the line restriction (l - l(0), l', l'') evaluated at the
quadrature nodes; here it is computed from full-D model calls like the reference does,
the kernels use closed-form line coefficients.  All arrays are batched over
chains: theta (B, D), rho (B, D), variates (B,).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .ref_port import gauss_hermite_probabilists


@dataclass
class FitConfig:
    family: str = "gauss"            # "gauss" | "sinh"
    N: int = 8
    initscale: float = 0.1
    tol: float = 1e-12               # 1e-10 for sinh (klhr_sinh.py:26)
    scale_clip: float = 600.0        # 300 for sinh (klhr_sinh.py:28)
    n1: int = 12                     # stage-1 iteration budget
    n2: int = 24                     # stage-2 iteration budget (Newton steps)
    nb: int = 8                      # back-tracking halvings per Newton step
    kmax: int = 0                    # cap on stage-2 KL evaluations per fit (0: 1 + n2 * nb)
    fix_d: bool = False              # sinh family with d = 1 frozen (reference sub_klhr_sinh.py)
    gtol1: float = 1e-4              # |l'| / sqrt(-l'') at the mode: stage 1 only supplies the START of stage 2
    gtol2: float = 1e-10             # inf-norm of the scaled KL gradient
    step_cap: float = 2.0            # inf-norm cap of a stage-2 step in scaled coordinates
    c1: float = 1e-4                 # Armijo constant (SciPy's c1, _optimize.py:1156)
    basin: float = 1e-3              # below this scaled-gradient norm take the full Newton step
    grad_clip: float = 0.0           # sinh family: elementwise clip of the model gradient inside KL
                                     # (klhr_sinh.py:158-161 clips at scale_clip, sub_klhr_sinh.py:152-154 at grad_clip);
                                     # applied for dim <= CLIP_MAX_D like the kernels; 0 = off
    eps: float = 2.220446049250313e-16

    @staticmethod
    def for_family(family, **kw):
        if family == "sinh":
            base = dict(family="sinh", tol=1e-10, scale_clip=300.0, n2=48, kmax=32)
            # KLHRSINH clips the model gradient at scale_clip (klhr_sinh.py:158-161); SUBKLHRSINH at grad_clip = 1e15
            base["grad_clip"] = 1e15 if kw.get("fix_d") else kw.get("scale_clip", 300.0)
        else:
            base = dict(family="gauss")
        base.update(kw)
        return FitConfig(**base)


CLIP_MAX_D = 16                      # csrc/klhr_fit.cuh:kClipMaxD


# ------------------------------------------------------------------ line restriction
def line_eval(model, theta, rho, y, grad_clip=0.0):
    """(l(y) - l(0), l'(y), l''(y)) along theta + y rho.  y: (B,) or (B, M).
    Non-finite anywhere -> (-inf, 0, 0), the batched form of reference
    ``bsmodel.py:15-30`` (failures become -inf / zero gradient).

    ``grad_clip`` > 0: l' is the projection of the ELEMENTWISE CLIPPED gradient, sum_i clip(g_i, +-c) rho_i, as in
    ``KLHRSINH.KL`` (``klhr_sinh.py:158-161,171-173``), wherever a component is actually clipped; evaluations on
    which nothing is clipped keep the plain l'.  l'' is always that of the unclipped density (the Newton matrix of
    the fixed-iteration optimiser: same fixed points)."""
    y = np.asarray(y, dtype=np.float64)
    flat = y.ndim == 1
    yy = y[:, None] if flat else y
    pts = theta[:, None, :] + yy[..., None] * rho[:, None, :]
    rb = np.broadcast_to(rho[:, None, :], pts.shape)
    with np.errstate(all="ignore"):
        lp, g = model.lp_grad(pts)
        l0 = model.lp(theta)
        l = lp - l0[:, None]
        l1 = np.sum(g * rb, axis=-1)
        l2 = model.dir2(pts, rb)
        bad = ~(np.isfinite(l) & np.isfinite(l1) & np.isfinite(l2))        # judged on the unclipped evaluation
        if grad_clip and grad_clip > 0 and theta.shape[-1] <= CLIP_MAX_D:
            gc = np.clip(g, -grad_clip, grad_clip)
            hit = np.any(gc != g, axis=-1)
            l1 = np.where(hit, np.sum(gc * rb, axis=-1), l1)
    l = np.where(bad, -np.inf, l)
    l1 = np.where(bad, 0.0, l1)
    l2 = np.where(bad, 0.0, l2)
    if flat:
        return l[:, 0], l1[:, 0], l2[:, 0]
    return l, l1, l2


def direction_from_normals(z, mvec, sd, tol):
    """rho = x / ||x + tol||, x = mvec + sd * z   (reference ``klhr.py:151-153``)."""
    x = mvec + sd * z
    nrm = np.sqrt(np.sum((x + tol) ** 2, axis=-1, keepdims=True))
    return x / nrm


def direction_mean(eigvecs, eigvals, method_one, family, uj=None):
    """Mean vector of the direction draw (reference ``klhr.py:144-150``,
    ``klhr_sinh.py:204-210``).  ``uj`` (B,) are the uniforms behind ``rng.choice``."""
    p = eigvals / np.sum(eigvals)
    if method_one:
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        j = np.searchsorted(cdf, uj, side="right")
        j = np.minimum(j, len(p) - 1)
        return eigvecs[:, j].T
    wgt = eigvals if family == "gauss" else p
    return np.sum(wgt * eigvecs, axis=1)[None, :]


# ------------------------------------------------------------------ stage 1
# stage 1: a Newton step may be at most this many trust radii long.  The candidates step * 2^-k are scored by
# l, so a long step is only taken when it really is the best point; 64 lets the exact Newton step of a
# quadratic target through at the posterior scales of the benchmark targets (ill-normal D = 100: |m*| up to
# ~25), which a cap of 8 split into two iterations for almost every warp.
NEWTON_CAP = 64.0


def stage1_mode(model, theta, rho, z_init, cfg: FitConfig):
    """Returns (xi_hat, tau0, n_evals).  tau0 = 1/2 log(-1/l''(xi_hat)) when l'' < 0 else 0
    (reference ``klhr.py:133-134``: ``(s > 0) * 0.5 * log(s)`` with s = hess_inv)."""
    B = theta.shape[0]
    xi = z_init * cfg.initscale
    trust = np.ones(B)
    done = np.zeros(B, dtype=bool)
    l, l1, l2 = line_eval(model, theta, rho, xi)
    nev = np.ones(B, dtype=np.int64)
    ks = 2.0 ** -np.arange(8)
    for _ in range(cfg.n1):
        with np.errstate(all="ignore"):
            concave = l2 < 0
            # |l'| / sqrt(-l'') <= gtol1 without the square root and the division
            conv = concave & (l1 * l1 <= cfg.gtol1 * cfg.gtol1 * (-l2))
            done = done | conv | ~np.isfinite(l)
            if done.all():
                break
            newton = np.where(concave, -l1 / np.where(concave, l2, -1.0), 0.0)
            stepc = np.clip(l1, -trust, trust)
            step = np.where(concave, np.clip(newton, -NEWTON_CAP * trust, NEWTON_CAP * trust), stepc)
        cand = xi[:, None] + step[:, None] * ks[None, :]
        cl, cl1, cl2 = line_eval(model, theta, rho, cand)
        nev += np.where(done, 0, 8)
        best = np.argmax(cl, axis=1)                     # first maximum on ties
        rows = np.arange(B)
        bl = cl[rows, best]
        improve = (bl > l) & ~done
        full = improve & (best == 0)
        xi = np.where(improve, cand[rows, best], xi)
        l1 = np.where(improve, cl1[rows, best], l1)
        l2 = np.where(improve, cl2[rows, best], l2)
        l = np.where(improve, bl, l)
        trust = np.where(full, trust * 2.0, np.where(improve, trust, trust * (1.0 / 256.0)))
        # no candidate improved: shrink and retry; give up once the step is negligible
        done = done | (~improve & (np.abs(step) * (1.0 / 128.0) <= cfg.eps * (1.0 + np.abs(xi))))
    with np.errstate(all="ignore"):
        tau0 = np.where(l2 < 0, -0.5 * np.log(np.where(l2 < 0, -l2, 1.0)), 0.0)
    return xi, tau0, nev


# ------------------------------------------------------------------ KL objective + Hessian
def _kl_gauss(model, theta, rho, eta, x, w, cfg):
    """f, scaled gradient (B,2), scaled Hessian (B,2,2) of reference ``klhr.py:106-120``
    in coordinates (m/s, tau)."""
    m, tau = eta[:, 0], eta[:, 1]
    with np.errstate(all="ignore"):
        s = np.exp(np.clip(tau, -cfg.scale_clip, cfg.scale_clip))
        y = s[:, None] * x[None, :] + m[:, None]
        l, l1, l2 = line_eval(model, theta, rho, y)
        S0 = np.sum(w * l, axis=1)
        S1 = np.sum(w * l1, axis=1)
        S1x = np.sum(w * l1 * x, axis=1)
        S2 = np.sum(w * l2, axis=1)
        S2x = np.sum(w * l2 * x, axis=1)
        S2xx = np.sum(w * l2 * x * x, axis=1)
        f = -(S0 + tau)
        g = np.stack([-S1 * s, -(S1x * s + 1.0)], axis=1)
        H = np.empty((len(m), 2, 2))
        s2 = s * s
        H[:, 0, 0] = -S2 * s2
        H[:, 0, 1] = H[:, 1, 0] = -S2x * s2
        H[:, 1, 1] = -(S2xx * s2 + S1x * s)
    return f, g, H


def _sinh_unpack(eta, cfg):
    c = cfg.scale_clip
    with np.errstate(all="ignore"):
        s = np.exp(np.clip(eta[:, 1], -c, c)) + cfg.tol
        d = np.exp(np.clip(eta[:, 2], -c, c)) + cfg.tol
    if cfg.fix_d:                                               # sub_klhr_sinh.py:92-97
        d = np.ones_like(s)
    return eta[:, 0], s, d, eta[:, 3]


def _kl_sinh(model, theta, rho, eta, x, w, cfg):
    """f, scaled gradient (B,4), scaled Hessian (B,4,4) of reference
    ``klhr_sinh.py:163-176`` in coordinates (m/s, log s, log d, e).  First derivatives are
    the reference's ``_grad_T`` / ``_grad_log_abs_jac`` (:116-124, :146-156); second
    derivatives are those expressions differentiated once more.  The elementwise clip of
    the model gradient (:158-161) enters through ``line_eval(..., grad_clip)``: l' is the projection
    of the clipped gradient, l'' stays that of the unclipped density."""
    m, s, d, e = _sinh_unpack(eta, cfg)
    B = len(m)
    c = cfg.scale_clip
    cx = np.arcsinh(x)[None, :]
    with np.errstate(all="ignore"):
        invd = (1.0 / d)[:, None]
        a = (cx + e[:, None]) * invd
        ac = np.clip(a, -c, c)
        sh, ch, th = np.sinh(ac), np.cosh(ac), np.tanh(ac)
        sech2 = 1.0 - th * th
        sN = s[:, None]
        T = m[:, None] + sN * sh
        l, l1, l2 = line_eval(model, theta, rho, T, grad_clip=cfg.grad_clip)
        logJ = (eta[:, 2] - eta[:, 1])[:, None] - np.log(ch)
        f = np.sum(w * (logJ - l), axis=1)
        # first derivatives (unscaled): order (m, sigma, delta, e)
        one = np.ones_like(a)
        zero = np.zeros_like(a)
        gT = np.stack([one, sN * sh, -sN * ch * a, sN * ch * invd], axis=-1)            # (B,N,4)
        gL = np.stack([zero, -one, 1.0 + th * a, -th * invd], axis=-1)
        g = np.sum(w[None, :, None] * (gL - l1[..., None] * gT), axis=1)               # (B,4)
        # second derivatives
        HT = np.zeros((B, len(x), 4, 4))
        HT[..., 1, 1] = sN * sh
        HT[..., 1, 2] = HT[..., 2, 1] = -sN * a * ch
        HT[..., 1, 3] = HT[..., 3, 1] = sN * ch * invd
        HT[..., 2, 2] = sN * a * ch + sN * a * a * sh
        HT[..., 2, 3] = HT[..., 3, 2] = -(sN * invd) * (ch + a * sh)
        HT[..., 3, 3] = sN * sh * invd * invd
        HL = np.zeros((B, len(x), 4, 4))
        HL[..., 2, 2] = -a * th - a * a * sech2
        HL[..., 2, 3] = HL[..., 3, 2] = (th + a * sech2) * invd
        HL[..., 3, 3] = -sech2 * invd * invd
        H = np.sum(w[None, :, None, None] * (HL - l2[..., None, None] * gT[..., :, None] * gT[..., None, :]
                                             - l1[..., None, None] * HT), axis=1)
        if cfg.fix_d:                  # d frozen: its row and column drop out of the Newton system
            g[:, 2] = 0.0
            H[:, 2, :] = 0.0
            H[:, :, 2] = 0.0
            H[:, 2, 2] = 1.0
        # scale coordinate 0 by s
        g[:, 0] *= s
        H[:, 0, :] *= s[:, None]
        H[:, :, 0] *= s[:, None]
    return f, g, H


def _chol_solve(H, g, n):
    """Solve H p = -g by Cholesky; ok=False where H is not numerically positive definite.
    Written entry by entry (n = 2 or 4) so the CUDA code can mirror the operation order."""
    B = H.shape[0]
    L = np.zeros_like(H)
    iL = np.zeros((B, n))
    ok = np.ones(B, dtype=bool)
    with np.errstate(all="ignore"):
        for j in range(n):
            acc = H[:, j, j].copy()
            for k in range(j):
                acc = acc - L[:, j, k] * L[:, j, k]
            ok &= acc > 0
            ljj = np.sqrt(np.where(acc > 0, acc, 1.0))
            L[:, j, j] = ljj
            iL[:, j] = 1.0 / ljj                    # one division per pivot, products elsewhere
            for i in range(j + 1, n):
                acc = H[:, i, j].copy()
                for k in range(j):
                    acc = acc - L[:, i, k] * L[:, j, k]
                L[:, i, j] = acc * iL[:, j]
        yv = np.zeros((B, n))
        for i in range(n):
            acc = -g[:, i]
            for k in range(i):
                acc = acc - L[:, i, k] * yv[:, k]
            yv[:, i] = acc * iL[:, i]
        p = np.zeros((B, n))
        for i in reversed(range(n)):
            acc = yv[:, i].copy()
            for k in range(i + 1, n):
                acc = acc - L[:, k, i] * p[:, k]
            p[:, i] = acc * iL[:, i]
    ok &= np.all(np.isfinite(p), axis=1)
    return p, ok


LM_SHIFTS = (0.0, 1e-3, 1e-2, 1e-1, 1.0, 10.0, 100.0, 1e3, 1e4, 1e6)


def _newton_direction(g, H, cfg):
    n = g.shape[1]
    B = g.shape[0]
    with np.errstate(all="ignore"):
        mu = np.max(np.abs(np.diagonal(H, axis1=1, axis2=2)), axis=1)
    mu = np.where(np.isfinite(mu) & (mu > 0), mu, 1.0)
    p = np.zeros((B, n))
    have = np.zeros(B, dtype=bool)
    eye = np.eye(n)[None]
    for lam in LM_SHIFTS:
        pk, ok = _chol_solve(H + (lam * mu)[:, None, None] * eye, g, n)
        take = ok & ~have
        p = np.where(take[:, None], pk, p)
        have |= take
        if have.all():
            break
    # no PD shift found (non-finite Hessian): fall back to steepest descent
    with np.errstate(all="ignore"):
        p = np.where(have[:, None], p, -g)
        p = np.where(np.isfinite(p), p, 0.0)
        big = np.max(np.abs(p), axis=1)
        scale = np.where(big > cfg.step_cap, cfg.step_cap / np.where(big > 0, big, 1.0), 1.0)
    return p * scale[:, None]


def stage2_newton(model, theta, rho, eta0, cfg: FitConfig, x, w):
    """Damped Newton on the KL objective; returns (eta, n_kl_evals, converged)."""
    kl = _kl_gauss if cfg.family == "gauss" else _kl_sinh
    B = theta.shape[0]
    eta = eta0.copy()
    f, g, H = kl(model, theta, rho, eta, x, w, cfg)
    nev = np.ones(B, dtype=np.int64)
    kmax = cfg.kmax if 0 < cfg.kmax < 1 + cfg.n2 * cfg.nb else 1 + cfg.n2 * cfg.nb
    done = np.zeros(B, dtype=bool)
    conv = np.zeros(B, dtype=bool)
    g_prev = np.full(B, np.inf)          # scaled-gradient norm when the previous direction was formed
    basin_prev = np.zeros(B, dtype=bool)
    slow = np.zeros(B, dtype=np.int64)
    for _ in range(cfg.n2):
        with np.errstate(all="ignore"):
            gmax = np.max(np.abs(g), axis=1)
        conv = conv | (~done & (gmax <= cfg.gtol2))
        done = done | conv
        # flat valley of the KL surface (the 4-parameter family is nearly non-identifiable there):
        # inside the basin Newton must at least halve the gradient; two steps in a row that do
        # not mean the remaining descent is along a flat ridge -- stop, the fit is already
        # tighter than the reference's own gtol (1e-3, klhr_sinh.py:199)
        with np.errstate(all="ignore"):
            lag = basin_prev & ~(gmax <= 0.5 * g_prev)
        slow = np.where(done, slow, np.where(lag, slow + 1, 0))
        done = done | (slow >= 2)
        g_prev = np.where(done, g_prev, gmax)
        with np.errstate(all="ignore"):
            basin_prev = np.where(done, basin_prev, gmax <= cfg.basin)
        if done.all():
            break
        p = _newton_direction(g, H, cfg)
        with np.errstate(all="ignore"):
            gp = np.sum(g * p, axis=1)
            gp = np.where(np.isfinite(gp), gp, 0.0)
        s_cur = _scale_of(eta, cfg)
        t = np.ones(B)
        accepted = np.zeros(B, dtype=bool)
        for _bt in range(cfg.nb):
            out_of_budget = ~done & ~accepted & (nev >= kmax)      # evaluation cap reached: keep the iterate
            done = done | out_of_budget
            active = ~done & ~accepted
            if not active.any():
                break
            trial = eta.copy()
            trial[:, 0] = eta[:, 0] + t * p[:, 0] * s_cur        # unscale coordinate 0
            trial[:, 1:] = eta[:, 1:] + t[:, None] * p[:, 1:]
            ft, gt, Ht = kl(model, theta, rho, trial, x, w, cfg)
            nev += active
            with np.errstate(all="ignore"):
                slack = 8.0 * cfg.eps * (1.0 + np.abs(f))
                okk = active & np.isfinite(ft) & ((ft <= f + cfg.c1 * t * gp + slack)
                                                  | ~np.isfinite(f) | (gmax <= cfg.basin))
            eta = np.where(okk[:, None], trial, eta)
            f = np.where(okk, ft, f)
            g = np.where(okk[:, None], gt, g)
            H = np.where(okk[:, None, None], Ht, H)
            accepted |= okk
            t = np.where(active & ~okk, t * 0.5, t)
        done = done | (~accepted)        # stalled: keep the current iterate
        done = done | (nev >= kmax)
    with np.errstate(all="ignore"):
        conv = conv | (np.max(np.abs(g), axis=1) <= cfg.gtol2)
    return eta, nev, conv


def _scale_of(eta, cfg):
    with np.errstate(all="ignore"):
        s = np.exp(np.clip(eta[:, 1], -cfg.scale_clip, cfg.scale_clip))
    return s + cfg.tol if cfg.family == "sinh" else s


# ------------------------------------------------------------------ family log densities
def logq_gauss(xv, eta, cfg):
    """reference ``klhr.py:155-158``."""
    with np.errstate(all="ignore"):
        s = np.exp(np.clip(eta[:, 1], -cfg.scale_clip, cfg.scale_clip))
        z = (xv - eta[:, 0]) / s
        return -np.log(s) - 0.5 * z * z


def transport_sinh(zv, eta, cfg):
    """reference ``klhr_sinh.py:112-114``."""
    m, s, d, e = _sinh_unpack(eta, cfg)
    c = cfg.scale_clip
    with np.errstate(all="ignore"):
        return m + s * np.sinh(np.clip((np.arcsinh(zv) + e) / d, -c, c))


def logq_sinh(xv, eta, cfg):
    """reference ``klhr_sinh.py:233-240``."""
    m, s, d, e = _sinh_unpack(eta, cfg)
    c = cfg.scale_clip
    with np.errstate(all="ignore"):
        z = (xv - m) / s
        b = np.clip(d * np.arcsinh(z) - e, -c, c)
        ti = np.sinh(b)
        return -0.5 * ti * ti + np.log(np.cosh(b)) + eta[:, 2] - eta[:, 1] - 0.5 * np.log1p(z * z)


# ------------------------------------------------------------------ the step
def fit(model, theta, rho, z_init, cfg: FitConfig, init4=None, xw=None):
    x, w = xw if xw is not None else gauss_hermite_probabilists(cfg.N)
    xi, tau0, nev1 = stage1_mode(model, theta, rho, z_init, cfg)
    B = theta.shape[0]
    if cfg.family == "gauss":
        eta0 = np.stack([xi, tau0], axis=1)
    else:
        eta0 = np.empty((B, 4))
        eta0[:, 0] = xi
        eta0[:, 1] = tau0
        eta0[:, 2:] = init4[:, 2:] * cfg.initscale           # klhr_sinh.py:191-193
        if cfg.fix_d:                                        # sub_klhr_sinh.py:184-186
            eta0[:, 3] = eta0[:, 2]
            eta0[:, 2] = 0.0
    eta, nev2, conv = stage2_newton(model, theta, rho, eta0, cfg, x, w)
    return eta, nev1 + nev2 * cfg.N, conv


def overrelaxed_quantile(u0, K, r, v):
    """u' of the over-relaxed proposal given u = CDF_q(0), the binomial count r and the beta
    variate v (reference ``klhr.py:163-172`` / ``klhr_sinh.py:217-227``)."""
    up = np.array(u0, dtype=np.float64, copy=True)
    hi = r > K - r
    lo = r < K - r
    up = np.where(hi, u0 * v, up)
    up = np.where(lo, 1.0 - (1.0 - u0) * v, up)
    return up


def step(model, theta, rho, z_init, z_prop, u, cfg: FitConfig, init4=None, xw=None, or_K=0, or_r=None,
         or_v=None):
    """One KLHR draw for every chain.  Returns a dict with eta, zp, r, accept, theta
    (after the step), evals (line evaluations executed), converged.  ``or_K > 0``: over-relaxed
    proposal with injected binomial counts ``or_r`` and beta variates ``or_v``."""
    import scipy.special as sp
    theta = np.asarray(theta, dtype=np.float64)
    rho = np.asarray(rho, dtype=np.float64)
    eta, evals, conv = fit(model, theta, rho, z_init, cfg, init4=init4, xw=xw)
    if cfg.family == "gauss":
        with np.errstate(all="ignore"):
            s = np.exp(np.clip(eta[:, 1], -cfg.scale_clip, cfg.scale_clip))
            if or_K > 0:                                         # klhr.py:160-173
                u0 = sp.ndtr((0.0 - eta[:, 0]) / s)
                z_prop = sp.ndtri(overrelaxed_quantile(u0, or_K, or_r, or_v))
            zp = eta[:, 0] + s * z_prop
        lq0 = logq_gauss(np.zeros_like(zp), eta, cfg)
        lq1 = logq_gauss(zp, eta, cfg)
    else:
        if or_K > 0:                                             # klhr_sinh.py:215-228
            m_, s_, d_, e_ = _sinh_unpack(eta, cfg)
            with np.errstate(all="ignore"):
                ti0 = np.sinh(np.clip(d_ * np.arcsinh((0.0 - m_) / s_) - e_, -cfg.scale_clip, cfg.scale_clip))
                z_prop = sp.ndtri(overrelaxed_quantile(sp.ndtr(ti0), or_K, or_r, or_v))
        zp = transport_sinh(z_prop, eta, cfg)
        lq0 = logq_sinh(np.zeros_like(zp), eta, cfg)
        lq1 = logq_sinh(zp, eta, cfg)
    lz, _, _ = line_eval(model, theta, rho, zp)
    with np.errstate(all="ignore"):
        r = lz + lq0 - lq1
        accept = np.log(u) < np.minimum(0.0, r)              # NaN compares False -> reject
    theta1 = np.where(accept[:, None], theta + zp[:, None] * rho, theta)
    return dict(eta=eta, zp=zp, r=r, accept=accept, theta=theta1, evals=evals + 2, converged=conv)


# ------------------------------------------------------------------ slice sampling along a line (N3)
def slice_step(model, theta, rho, e, u0, shrink_u, w=1.0, lower=-np.inf, upper=np.inf, max_out=1 << 20):
    """One draw of the reference's ``Slice._uni_slice`` (slice.py:84-146, ``m = inf``) for every chain, with
    injected variates: ``e`` (B,) standard exponentials, ``u0`` (B,) the uniform behind
    ``rng.uniform(0, w)``, ``shrink_u`` (B, S) the uniforms behind the shrinkage proposals
    ``rng.uniform(L, R) = L + (R - L) u`` (NaN = not available).  Returns x1 (accepted line coordinate),
    theta after the move, n_shrink (proposals consumed), evals (value calls: 1 + stepping-out tests +
    proposals), L, R (the interval the accepted proposal was drawn from).  A chain that exhausts its
    ``shrink_u`` keeps x1 = 0 and reports n_shrink = S + 1."""
    theta = np.asarray(theta, dtype=np.float64)
    rho = np.asarray(rho, dtype=np.float64)
    B = theta.shape[0]
    e, u0 = np.asarray(e, dtype=np.float64), np.asarray(u0, dtype=np.float64)
    g = lambda y: line_eval(model, theta, rho, y)[0]          # l(y) - l(0); -inf on failure
    logy = -e                                                  # gx0 - e with gx0 = l(0) - l(0) = 0
    off = w * u0
    L = 0.0 - off
    R = 0.0 + (w - off)
    evals = np.ones(B, dtype=np.int64)
    for side in (0, 1):                                        # stepping out (slice.py:95-107)
        act = np.ones(B, dtype=bool)
        for _ in range(max_out):
            act &= (L > lower) if side == 0 else (R < upper)
            if not act.any():
                break
            val = g(L if side == 0 else R)
            evals += act
            act &= ~(val <= logy)
            if side == 0:
                L = np.where(act, L - w, L)
            else:
                R = np.where(act, R + w, R)
    L = np.maximum(L, lower)                                   # slice.py:127-128
    R = np.minimum(R, upper)
    x1 = np.zeros(B)
    done = np.zeros(B, dtype=bool)
    n_shrink = np.zeros(B, dtype=np.int64)
    for k in range(shrink_u.shape[1]):                         # shrinkage (slice.py:131-139)
        uk = shrink_u[:, k]
        act = ~done & ~np.isnan(uk)
        if not act.any():
            break
        cand = L + (R - L) * np.where(act, uk, 0.0)
        val = g(cand)
        evals += act
        n_shrink += act
        ok = act & (val >= logy)
        x1 = np.where(ok, cand, x1)
        done |= ok
        rej = act & ~ok
        R = np.where(rej & (cand > 0.0), cand, R)
        L = np.where(rej & ~(cand > 0.0), cand, L)
    n_shrink = np.where(done, n_shrink, shrink_u.shape[1] + 1)
    return dict(x1=x1, theta=theta + x1[:, None] * rho, n_shrink=n_shrink, evals=evals, L=L, R=R, done=done)
