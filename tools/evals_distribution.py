"""Per-chain fit cost vs the lock-step cost of a 32-chain warp (development aid): mean evaluations per chain-draw
against the mean of the per-tile maximum."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, klhr_b200 as kb
dev = torch.device("cuda", 0)
for name, data, family in (("funnel", {"D": 1}, "sinh"), ("funnel", {"D": 10}, "sinh"), ("funnel", {"D": 1}, "gauss"),
                           ("rosenbrock", {"D": 2}, "gauss")):
    model = kb.BSModel(stan_file=f"stan/{name}.stan", data=data, device=dev)
    base = dict(family="sinh", tol=1e-10, scale_clip=300.0, n2=48, kmax=32) if family == "sinh" else dict(family="gauss")
    fit = kb.FitConfig(**base)
    B, S, D = 32768, 8, model.dim()
    th = (torch.randn(B, D, dtype=torch.float64, device=dev) * 0.5).contiguous()
    kb.run(model, fit, th, 200, 1)                       # towards stationarity
    tr = kb.Trace(S, B, D, fit.n_eta, torch.float64, dev, variates=False, rho=False)
    kb.run(model, fit, th, S, 1, draw_offset=200, trace=tr)
    torch.cuda.synchronize()
    ev = tr.evals.cpu().numpy().astype(np.float64)       # (S, B)
    tiles = ev.reshape(S, B // 32, 32)
    q = np.percentile(ev, [50, 90, 99, 99.9])
    print(f"{name} {data} {family}: mean {ev.mean():.1f}  p50/p90/p99/p99.9 {q}  max {ev.max():.0f}  "
          f"mean of tile max {tiles.max(-1).mean():.1f}  ratio {tiles.max(-1).mean() / ev.mean():.2f}")
