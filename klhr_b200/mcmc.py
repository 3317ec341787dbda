"""Sampler base with the surface of reference ``mcmc.py:3-46``, for a batch of chains.

Differences from the reference, all additive:
  * ``chains=B`` independent chains are advanced together; ``theta`` is ``(B, D)`` on the
    CUDA device (``(D,)`` views for ``chains == 1`` are returned as NumPy like the reference);
  * ``sample(M)`` returns ``(M, D)`` float64 NumPy for one chain (row 0 is the start point,
    mcmc.py:33-36) and a ``(M, B, D)`` device tensor for a batch;
  * ``sample_constrained`` works (the reference's is broken).
"""
from __future__ import annotations

import numpy as np
import torch


class MCMCBase:
    def __init__(self, model, stepsize, theta=None, seed=None, chains=1, dtype=torch.float64, device=None):
        self.model = model
        self.D = self.model.dim()
        self.stepsize = stepsize
        self.chains = int(chains)
        if self.chains < 1:
            raise ValueError("chains must be >= 1")
        self.dtype = dtype
        self.device = torch.device(device) if device is not None else model.device
        if self.device.type != "cuda":
            raise RuntimeError("klhr_b200 samplers run on a CUDA device only (no CPU fallback)")
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0] >> 1)
        self.seed = int(seed)
        self.rng = np.random.default_rng(self.seed)          # host-side uses only (start points)
        if theta is None:
            th = self.rng.normal(scale=0.1, size=(self.chains, self.D))
        else:
            th = np.asarray(theta.detach().cpu() if torch.is_tensor(theta) else theta, dtype=np.float64)
            th = np.broadcast_to(th.reshape(-1, self.D), (self.chains, self.D))
        self._theta = torch.as_tensor(np.array(th, dtype=np.float64, order="C"), dtype=dtype, device=self.device).contiguous()

    # ``theta``: (D,) NumPy for one chain like the reference, else the live (B, D) device tensor
    @property
    def theta(self):
        if self.chains == 1:
            return self._theta[0].double().cpu().numpy()
        return self._theta

    @theta.setter
    def theta(self, value):
        v = torch.as_tensor(np.asarray(value.detach().cpu() if torch.is_tensor(value) else value,
                                       dtype=np.float64))
        self._theta.copy_(v.reshape(-1, self.D).expand(self.chains, self.D))

    def __iter__(self):
        return self

    def __next__(self):
        return self.draw()

    def draw(self):
        raise NotImplementedError

    def log_density(self, theta):
        return self.model.log_density(theta)

    def sample(self, M):
        raise NotImplementedError

    def sample_constrained(self, M):
        out = self.sample(M)
        return self.model.constrain(out)
