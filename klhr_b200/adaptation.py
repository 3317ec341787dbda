"""Windowed adaptation for many pooled chains.

The reference adapts ONE chain over time with three helpers used only by the samplers:
``WindowedAdaptation`` (windowedadaptation.py:1-43), ``OnlineMoments``
(onlinemoments.py:3-28) and ``OnlinePCA`` (onlinepca.py:3-39).  The same three names and
methods exist here, but every ``update`` consumes a whole batch of chains and stores RAW
SUMS so that partial results merge exactly across CTAs, launches and ranks: one collective
on a flat fp64 buffer per window closure (an all-gather whose rows are then added in a
canonical tree order) is the only communication of the whole sampler.

CCIPCA is order dependent and cannot be pooled; ``OnlinePCA`` here keeps the second-moment
matrix sum (x x^T) of the same centred inputs the reference feeds to CCIPCA
(``theta - _mean``, klhr.py:218) and returns its leading eigenpairs, which is the quantity
CCIPCA converges to.  Parity with the reference is statistical for this part.

All tensors may live on any device (the unit tests drive the reductions with gloo on CPU;
production uses the CUDA tensors the step kernel accumulated into, over NCCL).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

try:                                     # only the K leading eigenpairs are needed: LAPACK's selected-eigenpair driver
    from scipy.linalg import eigh as _subset_eigh
except ImportError:                      # pragma: no cover
    _subset_eigh = None


class WindowedAdaptation:
    """Doubling window schedule; same constructor and ``window_closed(m)`` cursor semantics
    as reference ``windowedadaptation.py:1-43``.  ``warmup <= windowsize`` disables
    adaptation (the reference raises IndexError on equality)."""

    def __init__(self, warmup, windowsize=25, windowscale=2):
        if windowsize < 1 or windowscale < 1:
            raise ValueError("windowsize and windowscale must be >= 1")
        self._warmup, self._windowsize, self._windowscale = int(warmup), int(windowsize), int(windowscale)
        self._calculate_windows()

    def _calculate_windows(self):
        self._closures = []
        self._idx = 0
        size, close = self._windowsize, self._windowsize
        if self._warmup > self._windowsize:
            while close <= self._warmup:
                self._closures.append(close)
                if close == self._warmup:
                    break
                size *= self._windowscale
                close = self._warmup if close + self._windowscale * size >= self._warmup else close + size
        self._num_windows = len(self._closures)

    @property
    def closures(self):
        return list(self._closures)

    def next_closure(self, m):
        """Smallest closure iteration > m, or None."""
        for c in self._closures:
            if c > m:
                return c
        return None

    def window_closed(self, m):
        if not self._closures:
            return False
        closed = m == self._closures[self._idx]
        if closed and self._idx < self._num_windows - 1:
            self._idx += 1
        return closed

    def reset(self):
        self._calculate_windows()


class OnlineMoments:
    """Pooled running mean / sample variance.  ``update(x)`` takes a batch (B, D); raw sums
    of (x - shift) are kept in fp64.  ``var()`` follows onlinemoments.py:20-23: ones while
    N <= 2, otherwise the n-1 normalised variance."""

    def __init__(self, D, device=None, shift=None):
        self.D = D
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.shift = (torch.zeros(D, dtype=torch.float64, device=self.device) if shift is None
                      else shift.to(self.device, torch.float64))
        self.reset()

    def reset(self, shift=None):
        if shift is not None:
            self.shift = shift.to(self.device, torch.float64).clone()
        self.N = 0
        self.s1 = torch.zeros(self.D, dtype=torch.float64, device=self.device)
        self.s2 = torch.zeros(self.D, dtype=torch.float64, device=self.device)

    def update(self, x):
        x = torch.as_tensor(x, device=self.device).to(torch.float64).reshape(-1, self.D) - self.shift
        self.N += x.shape[0]
        self.s1 += x.sum(0)
        self.s2 += (x * x).sum(0)

    def add_sums(self, n, s1=None, s2=None):
        """Account for sums a kernel accumulated straight into ``self.s1`` / ``self.s2``."""
        self.N += int(n)
        if s1 is not None:
            self.s1 += s1
            self.s2 += s2

    def mean(self):
        if self.N == 0:
            return self.shift.clone()
        return self.shift + self.s1 / self.N

    def var(self):
        if self.N > 2:
            mu = self.s1 / self.N
            return (self.s2 - self.N * mu * mu) / (self.N - 1)
        return torch.ones(self.D, dtype=torch.float64, device=self.device)


class OnlinePCA:
    """Pooled replacement of CCIPCA (onlinepca.py): ``update(u)`` takes centred rows (B, D),
    ``values()`` / ``vectors()`` return the K leading eigenpairs of sum(u u^T)/n with the
    reference's ``+ tol`` convention on the values."""

    def __init__(self, D, K=1, l=0, tol=1e-10, device=None):
        self.D, self.K, self.l, self.tol = D, K, l, tol
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.reset()

    def reset(self):
        self.n = 0
        self.outer = torch.zeros(self.D, self.D, dtype=torch.float64, device=self.device)
        self._eig = None

    def update(self, u):
        u = torch.as_tensor(u, device=self.device).to(torch.float64).reshape(-1, self.D)
        self.n += u.shape[0]
        self.outer += u.T @ u
        self._eig = None

    def add_sums(self, n):
        self.n += int(n)
        self._eig = None

    def _solve(self):
        if self._eig is None:
            K = self.K
            if self.n < 1 or K == 0:
                self._eig = (np.zeros((self.D, K)), np.zeros(K))
            else:
                M = (self.outer / self.n).cpu().numpy()      # D x D, host eigh: identical bits on every rank
                M = 0.5 * (M + M.T)
                if not np.all(np.isfinite(M)):
                    self._eig = (np.zeros((self.D, K)), np.zeros(K))
                else:
                    if _subset_eigh is not None and 0 < K < self.D:
                        # dsyevx on the top-K window: 0.9 ms instead of 1.4 ms at D = 100, 3 ms instead of 19 ms at D = 256
                        # (the GPU idles while the host solves: 4 closures per warm-up)
                        w, V = _subset_eigh(M, subset_by_index=(self.D - K, self.D - 1), driver="evx")
                    else:
                        w, V = np.linalg.eigh(M)
                    order = np.argsort(w)[::-1][:K]
                    w, V = np.clip(w[order], 0.0, None), V[:, order]
                    for j in range(V.shape[1]):              # sign convention: largest |component| positive
                        k = int(np.argmax(np.abs(V[:, j])))
                        if V[k, j] < 0:
                            V[:, j] = -V[:, j]
                    if V.shape[1] < K:                       # K > D cannot happen for KLHR (J < D); pad anyway
                        V = np.pad(V, ((0, 0), (0, K - V.shape[1])))
                        w = np.pad(w, (0, K - w.shape[0]))
                    self._eig = (V, w)
        return self._eig

    def values(self):
        return self._solve()[1] + self.tol

    def vectors(self):
        return self._solve()[0]


class Smoother:
    """Robbins-Monro smoother of the over-relaxation count K; same API as reference
    ``smoother.py:3-20``.  ``update(d)`` accepts the POOLED mean of the reference's +-1 signal
    (klhr.py:220-221: +1 whenever the chain moved, since ``_msjd`` stays 0)."""

    def __init__(self, x0, kappa=-0.75):
        self._initial = x0
        self._x = x0
        self._kappa = kappa
        self._count = 0

    def update(self, d):
        self._count += 1
        k = self._count ** self._kappa
        self._x = k * (self._x + d) + (1 - k) * self._x

    def optimum(self):
        return self._x

    def reset(self):
        self._count = 0
        self._x = self._initial


def _tree_sum(rows):
    """Canonical pairwise tree over the rank index: ((r0 + r1) + (r2 + r3)) + ... -- the continuation of the tree
    the kernels use over chain slices (csrc/klhr_api.cu:outer_reduce_kernel), so that equal power-of-two shards
    reproduce the single-process sums bit for bit."""
    rows = list(rows)
    while len(rows) > 1:
        nxt = [rows[i] + rows[i + 1] for i in range(0, len(rows) - 1, 2)]
        if len(rows) % 2:
            nxt.append(rows[-1])
        rows = nxt
    return rows[0]


def allreduce_adaptation(moments, pca, group=None, extra=()):
    """Sum the raw adaptation state over all ranks: ONE collective per window closure on a flat fp64 buffer
    [N_moments, n_pca, s1 (D), s2 (D), outer (D*D), extra...].  The buffer is all-gathered and the ranks' rows are
    added in a canonical tree order on every rank (identical bits everywhere, and identical to the single-process
    sums for aligned power-of-two shards); NCCL's own all-reduce leaves the order of the additions unspecified.
    No-op without an initialised process group or with world size 1.  ``extra`` tensors are summed in place too."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    parts = []
    for mom in moments:
        parts += [mom.s1, mom.s2]
    parts += [pca.outer.reshape(-1)] + [e.reshape(-1) for e in extra]
    dev = parts[0].device
    counts = torch.tensor([float(m.N) for m in moments] + [float(pca.n)], dtype=torch.float64, device=dev)
    flat = torch.cat([counts] + [p_.to(torch.float64) for p_ in parts])
    world = dist.get_world_size(group)
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat, group=group)
    flat = _tree_sum(gathered)
    k = len(moments) + 1
    for i, mom in enumerate(moments):
        mom.N = int(round(float(flat[i])))
    pca.n = int(round(float(flat[len(moments)])))
    for p_ in parts:
        n = p_.numel()
        p_.copy_(flat[k:k + n].reshape(p_.shape))
        k += n
    pca._eig = None
