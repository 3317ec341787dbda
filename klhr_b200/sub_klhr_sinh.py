"""``SUBKLHRSINH`` -- KL Hit-and-Run with the 3-parameter sinh-arcsinh family (d = 1).

Drop-in for reference ``sub_klhr_sinh.py:14-284``: the same step as ``KLHRSINH`` with the tail-weight
parameter removed (eta = (m, log s, e), ``sub_klhr_sinh.py:92-97``).  On the device this is the sinh
kernel with ``KLHR_FIT_FIX_D``: d = 1 exactly and its row/column dropped from the Newton system.
``fit`` returns 3-vectors like the reference.
"""
from __future__ import annotations

from .klhr_sinh import KLHRSINH


class SUBKLHRSINH(KLHRSINH):
    def __init__(self, bsmodel, *args, **kwargs):
        super().__init__(bsmodel, *args, **kwargs)
        self._fit.fix_d = True

    def _kl_grad_clip(self):
        """sub_klhr_sinh.py:152-154: this variant clips at ``grad_clip`` (default 1e15: inactive)."""
        return float(self._grad_clip)

    def fit(self, rho, z_init=None):
        eta = super().fit(rho, z_init=z_init)
        return eta[..., [0, 1, 3]]

    def KL(self, eta, rho):
        """3-parameter ``KL`` (reference sub_klhr_sinh.py): eta = (m, log s, e)."""
        import numpy as np
        import torch
        e = torch.as_tensor(np.asarray(eta.detach().cpu() if torch.is_tensor(eta) else eta, dtype=np.float64)).reshape(-1, 3)
        e4 = torch.stack([e[:, 0], e[:, 1], torch.zeros_like(e[:, 0]), e[:, 2]], dim=1)
        f, g = super().KL(e4, rho)
        return f, g[..., [0, 1, 3]]
