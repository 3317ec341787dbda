"""``KLHRSINH`` -- KL Hit-and-Run with the 4-parameter sinh-arcsinh line family.

Drop-in for reference ``klhr_sinh.py:14-289``: constructor keywords and defaults of
klhr_sinh.py:15-32 (``tol=1e-10``, ``scale_clip=300``, ``eigen_method_one=False``), no
clipping of ``J`` (:37), normalised eigenvalue weights in the method-two direction mean
(:210), ``overrelaxed=True`` (:30) with the over-relaxed proposal of :215-228 (binomial / beta
variates from the chain's Philox stream instead of SciPy's global RNG) and a fixed K (:279).
"""
from __future__ import annotations

import torch

from .klhr import KLHR


class KLHRSINH(KLHR):
    _family = "sinh"
    _eigen_weights_normalised = True           # klhr_sinh.py:210 uses p = evals / sum(evals)
    _adapt_K = False                           # klhr_sinh.py:279 has the K update commented out

    def __init__(self, bsmodel, theta=None, seed=None, N=8, K=10, J=2, l=4, initscale=0.1, warmup=1_000,
                 windowsize=50, windowscale=2, tol=1e-10, grad_clip=1e15, scale_clip=300,
                 scale_dir_cov=False, overrelaxed=True, eigen_method_one=False, max_init_tries=100, *,
                 chains=1, dtype=torch.float64, device=None, process_group=None, chain_offset=None,
                 pca_stride=None, moments_every_draw=False, fit_budget=None):
        if dtype != torch.float64:
            raise TypeError("the sinh-arcsinh family needs float64 (sinh/cosh of up to +-300, "
                            "klhr_sinh.py:100-110)")
        super().__init__(bsmodel, theta=theta, seed=seed, N=N, K=K, J=J, l=l, initscale=initscale,
                         warmup=warmup, windowsize=windowsize, windowscale=windowscale, tol=tol,
                         grad_clip=grad_clip, scale_clip=scale_clip, scale_dir_cov=scale_dir_cov,
                         overrelaxed=overrelaxed, eigen_method_one=eigen_method_one,
                         max_init_tries=max_init_tries, chains=chains, dtype=dtype, device=device,
                         process_group=process_group, chain_offset=chain_offset, pca_stride=pca_stride,
                         moments_every_draw=moments_every_draw,
                         fit_budget=fit_budget)

    def _clip_J(self, J):
        return J                                # klhr_sinh.py:37 does not clip

    def _kl_grad_clip(self):
        """klhr_sinh.py:158-161: ``KL`` clips every component of the model gradient at ``scale_clip`` (sic)."""
        return float(self._scale_clip)
