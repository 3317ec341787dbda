"""Adaptation utilities of the KLHR hot path, restated -- TEST INFRASTRUCTURE.

Follows reference ``onlinemoments.py:3-28`` (Welford running moments),
``onlinepca.py:3-39`` (CCIPCA, Weng et al.) and ``windowedadaptation.py:1-43`` (doubling
window schedule).  The per-sample recurrences are kept as the reference has them so the
single-chain port (ref_port.py) reproduces the reference bit for bit; the ``pooled_*``
functions are the raw-sum forms the multi-chain product reduces across chains and ranks
 and are checked against the recurrences in tests.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- schedule
def window_closures(warmup: int, windowsize: int = 25, windowscale: int = 2) -> list[int]:
    """Iterations at which an adaptation window closes (reference
    ``windowedadaptation.py:12-28``): the window length is multiplied by ``windowscale``
    after every closure and the last window is stretched to end exactly at ``warmup``.
    ``warmup <= windowsize`` yields no closures (the reference raises IndexError on
    equality; callers validate)."""
    out: list[int] = []
    if warmup <= windowsize:
        return out
    size = windowsize
    close = windowsize
    while close <= warmup:
        out.append(close)
        if close == warmup:
            break
        size *= windowscale
        if close + windowscale * size >= warmup:
            close = warmup
        else:
            close = close + size
    return out


class WindowSchedule:
    """``window_closed(m)`` with the reference's cursor semantics
    (``windowedadaptation.py:30-37``)."""

    def __init__(self, warmup, windowsize=25, windowscale=2):
        self.closures = window_closures(warmup, windowsize, windowscale)
        self.warmup = warmup
        self.windowsize = windowsize
        self._i = 0

    def window_closed(self, m: int) -> bool:
        if not self.closures:
            return False
        hit = m == self.closures[self._i]
        if hit and self._i < len(self.closures) - 1:
            self._i += 1
        return hit


# ------------------------------------------------------------------------------ moments
class RunningMoments:
    """Welford recurrence exactly as reference ``onlinemoments.py:10-23``."""

    def __init__(self, D):
        self.D = D
        self.reset()

    def reset(self):
        self.n = 0
        self.m = np.zeros(self.D)
        self.v = np.zeros(self.D)

    def update(self, x):
        self.n += 1
        w = 1 / self.n
        d = x - self.m
        self.m += d * w
        self.v += -self.v * w + d * d * w * (1 - w)

    def mean(self):
        return self.m

    def var(self):
        if self.n > 2:
            return self.v * self.n / (self.n - 1)
        return np.ones(self.D)


def pooled_moments(n, s1, s2, shift=None):
    """Mean / sample variance from raw pooled sums of (x - shift): the merge-safe form
    of ``RunningMoments`` (same ``n <= 2 -> ones`` rule, ``onlinemoments.py:20-23``)."""
    s1 = np.asarray(s1, dtype=np.float64)
    s2 = np.asarray(s2, dtype=np.float64)
    if shift is None:
        shift = np.zeros_like(s1)
    if n == 0:
        return np.zeros_like(s1) + shift * 0.0, np.ones_like(s1)
    mu = s1 / n
    if n <= 2:
        return mu + shift, np.ones_like(s1)
    var = (s2 - n * mu * mu) / (n - 1)
    return mu + shift, var


# ---------------------------------------------------------------------------------- pca
class StreamingPCA:
    """CCIPCA with amnesic parameter ``l`` as reference ``onlinepca.py:13-35``."""

    def __init__(self, D, K=1, l=0, tol=1e-10):
        self.D, self.K, self.l, self.tol = D, K, l, tol
        self.reset()

    def reset(self):
        self.n = 0
        self.v = np.zeros((self.D, self.K))

    def update(self, u):
        self.n += 1
        for i in range(min(self.K, self.n)):
            if i == self.n - 1:
                self.v[:, i] = u
            else:
                w = (self.n - 1 - self.l) / self.n
                v = self.v[:, i]
                nv = np.linalg.norm(v)
                self.v[:, i] = w * v + (1 - w) * u * u.dot(v) / (nv + self.tol)
                v = self.v[:, i]
                nv = np.linalg.norm(v)
                u = u - u.dot(v) * v / (nv * nv + self.tol)

    def values(self):
        nv = np.linalg.norm(self.v, axis=0)
        if np.any(np.isnan(nv) | np.isinf(nv)):
            nv = np.zeros_like(self.v)
        return nv + self.tol

    def vectors(self):
        return self.v / self.values()


def pooled_pca(n, s1, s_outer, J, tol=1e-10):
    """Top-J eigenpairs of the pooled covariance built from raw sums
    (n, sum x, sum x x^T).  Replaces the order-dependent CCIPCA when many chains are
    pooled; returns (vectors (D,J), values (J,)) with the same
    ``+ tol`` convention as ``StreamingPCA.values``.  Sign convention: the largest-|.|
    component of each vector is positive."""
    s1 = np.asarray(s1, dtype=np.float64)
    D = s1.shape[0]
    if n < 2:
        return np.zeros((D, J)), np.full(J, tol)
    mu = s1 / n
    C = (np.asarray(s_outer, dtype=np.float64) - n * np.outer(mu, mu)) / (n - 1)
    C = 0.5 * (C + C.T)
    w, V = np.linalg.eigh(C)
    order = np.argsort(w)[::-1][:J]
    w = np.clip(w[order], 0.0, None)
    V = V[:, order]
    for j in range(V.shape[1]):
        k = np.argmax(np.abs(V[:, j]))
        if V[k, j] < 0:
            V[:, j] = -V[:, j]
    return V, w + tol
