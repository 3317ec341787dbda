"""Quick device timing of klhr_run for one configuration (development aid; not the bench).
KLHR_SM100_LIB selects an alternative build of the library."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import klhr_b200 as kb

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ill-normal")
ap.add_argument("--data", default='{"D": 100}')
ap.add_argument("--family", default="gauss")
ap.add_argument("--chains", type=int, default=65536)
ap.add_argument("--draws", type=int, default=200)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--tag", default="")
ap.add_argument("--octet", action="store_true")
ap.add_argument("--tile", action="store_true", help="tile kernel where the lane kernel would be chosen")
ap.add_argument("--peaks", action="store_true", help="print the peak micro-kernels (klhr_peak_probe) and exit")
ap.add_argument("--slice", action="store_true", help="time klhr_slice_run instead of klhr_run")
ap.add_argument("--direction", action="store_true", help="eigen_method_one law: 2 mean columns + zero column")
a = ap.parse_args()
dev = torch.device("cuda", 0)
if a.peaks:
    from klhr_b200 import engine
    for k in engine.PROBE_KINDS:
        print(f"peak {k:14s} {engine.peak_probe(k, dev):.4e} ops/s")
    sys.exit(0)
dt = torch.float64 if a.dtype == "f64" else torch.float32
data = json.load(open(a.data[1:])) if a.data.startswith("@") else json.loads(a.data)
model = kb.BSModel(stan_file=f"stan/{a.model}.stan", data=data, device=dev)
base = dict(family="sinh", tol=1e-10, scale_clip=300.0, n2=48, kmax=32) if a.family == "sinh" else dict(family="gauss")
fit = kb.FitConfig(**base).for_dtype(dt)
fit.force_octet = a.octet
fit.force_tile = a.tile
D = model.dim()
th = (torch.randn(a.chains, D, dtype=torch.float64, device=dev) * 0.5).to(dt).contiguous()
acc = torch.zeros(a.chains, dtype=torch.int64, device=dev)
ev = torch.zeros(1, dtype=torch.int64, device=dev)
direction = None
if a.direction:
    cols = torch.randn(2, D, dtype=torch.float64, device=dev).to(dt).contiguous()
    direction = kb.Direction(mean_cols=cols, sd=torch.ones(D, dtype=dt, device=dev),
                             cdf=torch.tensor([0.4, 0.7, 1.0], dtype=dt, device=dev), n_zero_cols=1)
if a.slice:
    scfg = kb.SliceConfig()
    _run = lambda n, off: kb.slice_run(model, scfg, th, n, 1, direction, draw_offset=off, accept_count=acc, evals_total=ev)
else:
    _run = lambda n, off: kb.run(model, fit, th, n, 1, direction, draw_offset=off, accept_count=acc, evals_total=ev)
_run(50, 0)
torch.cuda.synchronize()
acc.zero_(); ev.zero_()
times = []
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _run(a.draws, 50 + r * a.draws)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = float(np.median(times))
n = a.chains * a.draws
info = kb.launch_info(model, fit, dtype=dt, device=dev)
print(f"{a.tag or os.environ.get('KLHR_SM100_LIB', 'default'):32s} {a.model} {a.family} {a.dtype} D={D} B={a.chains}: "
      f"{n / ms / 1e6:8.3f} Mdraws/ms = {n / (ms * 1e-3):.3e} draws/s  ms={ms:.2f} acc={float(acc.double().mean()) / (a.draws * a.reps):.4f} "
      f"evals/draw={ev.item() / (n * a.reps):.1f} regs={info['regs']} ctas/sm={info['ctas_per_sm']} finite={bool(torch.isfinite(th).all())}")
