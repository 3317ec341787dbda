"""Development aid: the chain kernel's thread-per-chain D-phase (KLHR_CHAIN_SERIAL=1, default) against the octet
D-phase it replaces (KLHR_CHAIN_SERIAL=0) on identical inputs.  `python tools/dev/serial_dphase_ab.py run OUT.npz`
in two processes with the two settings, then `... cmp A.npz B.npz`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

CASES = [("funnel", {"D": 1}, "sinh"), ("funnel", {"D": 10}, "sinh")]

if sys.argv[1] == "run":
    import torch, klhr_b200 as kb
    dev = torch.device("cuda", 0)
    out = {}
    for name, data, family in CASES:
        model = kb.BSModel(stan_file=f"stan/{name}.stan", data=data, device=dev)
        D = model.dim()
        fit = kb.FitConfig(family="sinh", tol=1e-10, scale_clip=300.0, n2=48, kmax=32, grad_clip=300.0)
        rng = np.random.default_rng(5)
        B, S = 65536, 30
        th0 = rng.normal(size=(B, D)) * np.r_[3.0, np.ones(D - 1) * 4.0]
        cols = np.zeros((2, D)); cols[0, 0] = 2.0; cols[1, D - 1] = 3.0
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
        direction = kb.Direction(mean_cols=up(cols), sd=up(np.r_[3.0, np.ones(D - 1) * 9.0]), cdf=up(np.array([0.3, 0.6, 1.0])),
                                 n_zero_cols=1)
        th = up(th0)
        tr = kb.Trace(S, B, D, 4, torch.float64, dev, variates=True, rho=True)
        kb.run(model, fit, th, S, 3, direction, trace=tr)
        torch.cuda.synchronize()
        k = f"{name}{D}"
        out[k + "_eta"] = tr.eta.cpu().numpy(); out[k + "_acc"] = tr.accept.cpu().numpy()
        out[k + "_ev"] = tr.evals.cpu().numpy(); out[k + "_rho"] = tr.rho.cpu().numpy(); out[k + "_th"] = th.cpu().numpy()
    np.savez(sys.argv[2], **out)
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for name, data, _ in CASES:
        k = f"{name}{data['D'] + 1}"
        S = a[k + "_eta"].shape[0]
        print(k, "evals/draw", a[k + "_ev"].mean(), b[k + "_ev"].mean(), "acc", a[k + "_acc"].mean(), b[k + "_acc"].mean())
        for s in (0, 1, 2, 5, 10, S - 1):
            de = np.abs(a[k + "_eta"][s] - b[k + "_eta"][s]).max(axis=1)
            dr = np.abs(a[k + "_rho"][s] - b[k + "_rho"][s]).max(axis=1)
            print(f"  draw {s:2d}: rho max diff {dr.max():.2e}  chains with |d eta| > 1e-9: {(de > 1e-9).mean():.5f}  > 1e-6: {(de > 1e-6).mean():.5f}"
                  f"  accept flags differ: {(a[k + '_acc'][s] != b[k + '_acc'][s]).mean():.6f}  evals differ: {(a[k + '_ev'][s] != b[k + '_ev'][s]).mean():.5f}")
