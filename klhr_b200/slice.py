"""``Slice`` -- univariate slice sampling (stepping out + shrinkage) along KLHR's adapted random
directions, for a batch of chains on B200.

Drop-in for reference ``slice.py:12-176`` (an ``ALGORITHM`` of ``experiment_accuracy.py:56-64``,
``experiment_ar1.py:69-79``, ``experiment_funnel.py:42-52``): same class name, constructor keywords and
defaults (slice.py:14-32).  The direction law and the windowed adaptation are KLHR's (slice.py:148-176 ==
klhr.py:143-153,196-223), so the class reuses ``KLHR``'s host logic and swaps the transition kernel for
``klhr_slice_run`` (csrc/klhr_slice.cuh).

Reference quirks: only ``m = inf`` is supported -- the reference's finite-``m`` branch raises NameError
(slice.py:108,124); J is not clipped to D - 1 (slice.py:41); ``overrelaxed`` is accepted and ignored
(slice.py:30); every draw moves, so ``acceptance_probability`` is 1 after the first draw (slice.py:143-144).
``grad_evals`` counts the line evaluations (the reference leaves a TODO, slice.py:66).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .klhr import KLHR


class Slice(KLHR):
    _family = "gauss"          # only used for the (unused) fit configuration
    _adapt_K = False

    def __init__(self, bsmodel, theta=None, seed=None, w=1, m=np.inf, lower=-np.inf, upper=np.inf, J=2, l=4,
                 initscale=0.1, warmup=1_000, windowsize=50, windowscale=2, tol=1e-12, scale_dir_cov=False,
                 overrelaxed=False, eigen_method_one=True, max_init_tries=100, *, chains=1, dtype=torch.float64,
                 device=None, process_group=None, chain_offset=None, pca_stride=None, shrink_trace=24):
        if not np.isinf(m):
            raise NotImplementedError("Slice: only m = inf (unlimited stepping out) is supported; the reference's "
                                      "finite-m branch cannot run (slice.py:108,124)")
        if J > bsmodel.dim():
            raise ValueError("J must not exceed the model dimension")
        self._slice = engine.SliceConfig(w=float(w), lower=float(lower), upper=float(upper), tol=float(tol),
                                         cap=int(shrink_trace))
        self.w, self.m, self.lower, self.upper = w, m, lower, upper
        super().__init__(bsmodel, theta=theta, seed=seed, J=J, l=l, initscale=initscale, warmup=warmup,
                         windowsize=windowsize, windowscale=windowscale, tol=tol, scale_dir_cov=scale_dir_cov,
                         overrelaxed=False, eigen_method_one=eigen_method_one, max_init_tries=max_init_tries,
                         chains=chains, dtype=dtype, device=device, process_group=process_group,
                         chain_offset=chain_offset, pca_stride=pca_stride, moments_every_draw=False)

    def _clip_J(self, J):
        return J                                               # slice.py:41

    def _launch(self, steps, **kw):
        engine.slice_run(self.model, self._slice, self._theta, steps, self.seed, self._direction,
                         chain_offset=self._chain_offset, draw_offset=self._draw,
                         accept_count=self._accept_count, evals_total=self._evals_total, **kw)

    def fit(self, rho, z_init=None):
        raise NotImplementedError("Slice has no line fit")
