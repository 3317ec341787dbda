"""Fraction of sinh-family fits of the fixed-iteration oracle that land within 1e-5 of the UNMODIFIED reference's
fit on the same (theta, rho, variates), with and without the elementwise gradient clip of klhr_sinh.py:158-161
(tests/golden tapes; CPU only).  Also how often the clip fires inside the oracle's iteration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import batched, stan_models
from oracle.batched import FitConfig

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name in ("funnel_d2_sinh", "funnel_d2_sinh_tight", "funnel_d2_sinh_scaledir_method1", "funnel_d2_subsinh_tight",
             "ark_t200_sinh", "earnings_sinh", "rosenbrock_d4_sinh", "rosenbrock_d4_subsinh"):
    t = dict(np.load(os.path.join(G, name + ".npz")))
    meta, data = json.loads(str(t.pop("meta_json"))), json.loads(str(t.pop("data_json")))
    model = stan_models.make_model(meta["model"], data)
    row = [f"{name:34s}"]
    for clip in (False, True):
        cfg = FitConfig.for_family("sinh", fix_d=meta["family"] == "subsinh")
        if not clip:
            cfg.grad_clip = 0.0
        calls = [0, 0]
        inner = batched.line_eval

        def counting(m, th, rh, y, grad_clip=0.0, _inner=inner):
            if grad_clip:
                yy = np.asarray(y)
                yy = yy[:, None] if yy.ndim == 1 else yy
                with np.errstate(all="ignore"):
                    g = m.lp_grad(th[:, None, :] + yy[..., None] * rh[:, None, :])[1]
                calls[0] += int(np.any(np.abs(g) > grad_clip, axis=-1).sum())
                calls[1] += yy.size
            return _inner(m, th, rh, y, grad_clip)
        batched.line_eval = counting
        try:
            out = batched.step(model, t["theta0"], t["rho"], t["z_init"], t["z_prop"], t["u"], cfg,
                               init4=t.get("init4"), xw=(t["x_nodes"], t["w_nodes"]))
        finally:
            batched.line_eval = inner
        s = np.exp(t["eta"][:, 1])
        em = np.abs(out["eta"][:, 0] - t["eta"][:, 0]) / s
        es = np.abs(out["eta"][:, 1:] - t["eta"][:, 1:]).max(axis=1)
        good = (em <= 1e-5) & (es <= 1e-5)
        row.append(f"clip={'on ' if clip else 'off'} within1e-5={good.mean():.4f} conv={out['converged'].mean():.4f} "
                   f"flags_equal={(out['accept'] == t['accept']).mean():.4f}"
                   + (f" clipped_evals={calls[0]}/{calls[1]}" if clip else ""))
    print(" | ".join(row))
