"""Stage-1 tolerance (gtol1) x stage-2 cap (kmax) sweep (development aid): lock-step cost, acceptance, throughput."""
import sys, os, json; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, klhr_b200 as kb
dev = torch.device("cuda", 0)
ark = json.load(open(os.path.join(os.path.dirname(__file__), "ark10k.json")))
cases = [("funnel", {"D": 1}, "sinh"), ("funnel", {"D": 10}, "sinh"), ("funnel", {"D": 1}, "gauss"),
         ("rosenbrock", {"D": 2}, "gauss"), ("arK", ark, "gauss")]
for name, data, family in cases:
    model = kb.BSModel(stan_file=f"stan/{name}.stan", data=data, device=dev)
    D = model.dim()
    for gtol1 in (1e-8, 1e-6, 1e-4):
        for kmax in ((32, 16) if family == "sinh" else (0,)):
            base = dict(family="sinh", tol=1e-10, scale_clip=300.0, n2=48, kmax=kmax) if family == "sinh" else dict(family="gauss")
            fit = kb.FitConfig(**base, gtol1=gtol1)
            B, S = 65536, 8
            th = (torch.randn(B, D, dtype=torch.float64, device=dev) * 0.5).contiguous()
            kb.run(model, fit, th, 200, 1)
            tr = kb.Trace(S, B, D, fit.n_eta, torch.float64, dev, variates=False, rho=False)
            kb.run(model, fit, th, S, 1, draw_offset=200, trace=tr)
            torch.cuda.synchronize()
            ev = tr.evals.cpu().numpy().astype(np.float64)
            a = tr.accept.double().mean().item()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); kb.run(model, fit, th, 50, 1, draw_offset=300); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            tiles = ev.reshape(S, B // 32, 32)
            print(f"{name} D={D} {family} gtol1={gtol1:g} kmax={kmax}: mean evals {ev.mean():6.1f} tile-max {tiles.max(-1).mean():6.1f} "
                  f"accept {a:.4f}  {B * 50 / ms / 1e3:.1f} Mdraws/s", flush=True)
