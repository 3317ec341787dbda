"""Stage-2 evaluation cap (kmax) sweep for the sinh family on funnel (development aid): cost of the lock-step
warp (mean of the per-tile maximum), fraction of fits that stop on the cap, acceptance rate, throughput."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, klhr_b200 as kb
dev = torch.device("cuda", 0)
for data in ({"D": 1}, {"D": 10}):
    model = kb.BSModel(stan_file="stan/funnel.stan", data=data, device=dev)
    D = model.dim()
    for kmax in (12, 16, 20, 24, 32, 48):
        fit = kb.FitConfig(family="sinh", tol=1e-10, scale_clip=300.0, n2=48, kmax=kmax)
        B, S = 65536, 8
        th = (torch.randn(B, D, dtype=torch.float64, device=dev) * 0.5).contiguous()
        acc = torch.zeros(B, dtype=torch.int64, device=dev)
        kb.run(model, fit, th, 200, 1)
        tr = kb.Trace(S, B, D, 4, torch.float64, dev, variates=False, rho=False)
        kb.run(model, fit, th, S, 1, draw_offset=200, trace=tr)
        torch.cuda.synchronize()
        ev = tr.evals.cpu().numpy().astype(np.float64)
        a = tr.accept.double().mean().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); kb.run(model, fit, th, 50, 1, draw_offset=300); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        tiles = ev.reshape(S, B // 32, 32)
        print(f"funnel {data} kmax={kmax:3d}: mean evals {ev.mean():6.1f} tile-max {tiles.max(-1).mean():6.1f} "
              f"p99 {np.percentile(ev, 99):5.0f} accept {a:.4f}  {B * 50 / ms / 1e3:.1f} Mdraws/s")
