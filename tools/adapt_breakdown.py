"""Development aid: where the warm-up (adaptation) phase spends its wall time.  Synchronises after every part, so
the total is an upper bound of the pipelined run."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import klhr_b200 as kb
dev = torch.device("cuda", 0)
model = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": 100}, device=dev)
for rep in range(2):
    s = kb.KLHR(model, seed=1, chains=65536, warmup=1000, device=dev)
    T = collections.defaultdict(float); N = collections.Counter()
    def wrap(name):
        f = getattr(s, name)
        def g(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = f(*a, **k)
            torch.cuda.synchronize(); T[name] += time.perf_counter() - t0; N[name] += 1
            return r
        setattr(s, name, g)
    for nm in ("_launch", "_snapshot_update", "_close_window", "_refresh_direction"):
        wrap(nm)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.run(1000)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print(f"rep {rep}: total {tot*1e3:.1f} ms  " + "  ".join(f"{k}: {v*1e3:.1f} ms / {N[k]}" for k, v in T.items()))
s2 = kb.KLHR(model, seed=1, chains=65536, warmup=1000, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter(); s2.run(1000); torch.cuda.synchronize()
print(f"unsynchronised total {1e3*(time.perf_counter()-t0):.1f} ms")
