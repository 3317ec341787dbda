"""``BSModel`` shim -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Same surface as reference ``bsmodel.py:5-55`` (``dim``, ``log_density``,
``log_density_gradient``, ``constrain``/``unconstrain``, ``parameter_names``) but backed
by the analytic NumPy models of ``stan_models.py`` instead of BridgeStan, which is not
installable in this image.  Placing this directory ahead of /root/reference on
``sys.path`` lets the UNMODIFIED reference samplers import it by the bare name
``bsmodel`` (reference ``klhr.py:8``); ``make_golden.py`` does exactly that.

Error behaviour follows the reference wrapper: any failure or non-finite value yields
``-inf`` / zero gradient rather than an exception (reference ``bsmodel.py:15-30``).
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

try:  # imported both as ``oracle.bsmodel`` and as bare ``bsmodel`` (see docstring)
    from .stan_models import make_model
except ImportError:  # pragma: no cover - bare-name import path
    from stan_models import make_model


class BSModel:
    def __init__(self, stan_file="", data_file="", stepsize=1.0, warn=False, data=None):
        self._stan_file = stan_file
        self._data_file = data_file
        name = Path(str(stan_file)).stem
        if data is None:
            data = json.loads(Path(data_file).read_text())
        self.name = name
        self.data = data
        self.model = make_model(name, data)
        self.n_value_calls = 0
        self.n_grad_calls = 0

    def dim(self):
        return self.model.dim()

    def log_density(self, theta, **kws):
        self.n_value_calls += 1
        ld = float(self.model.lp(np.asarray(theta, dtype=np.float64)))
        return ld if np.isfinite(ld) else -np.inf

    def log_density_gradient(self, theta, **kws):
        self.n_grad_calls += 1
        theta = np.asarray(theta, dtype=np.float64)
        ld, g = self.model.lp_grad(theta)
        if not (np.isfinite(ld) and np.all(np.isfinite(g))):
            return -np.inf, np.zeros_like(theta)
        return float(ld), g

    def constrain(self, theta):
        theta = np.array(theta, dtype=np.float64)
        if self.name == "arK":
            theta[..., -1] = np.exp(theta[..., -1])
        if self.name == "earnings":
            theta[..., 2:] = np.exp(theta[..., 2:])
        return theta

    def unconstrain(self, theta):
        theta = np.array(theta, dtype=np.float64)
        if self.name == "arK":
            theta[..., -1] = np.log(theta[..., -1])
        if self.name == "earnings":
            theta[..., 2:] = np.log(theta[..., 2:])
        return theta

    def parameter_names(self):
        m = self.model
        if self.name == "funnel":
            return ["double_log_sigma"] + [f"alpha.{i + 1}" for i in range(m.Da)]
        if self.name == "arK":
            return ["alpha"] + [f"beta.{i + 1}" for i in range(m.K)] + ["sigma"]
        if self.name == "earnings":
            return ["beta.1", "beta.2", "sigma", "s"]
        if self.name == "rosenbrock":
            return [f"v.{i + 1}" for i in range(m.Dh)] + [f"theta.{i + 1}" for i in range(m.Dh)]
        return [f"y.{i + 1}" for i in range(m.dim())]
