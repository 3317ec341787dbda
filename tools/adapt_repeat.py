import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import klhr_b200 as kb
dev = torch.device("cuda", 0)
model = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": 100}, device=dev)
for rep in range(6):
    s = kb.KLHR(model, seed=1, chains=65536, warmup=1000, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter(); s.run(1000); torch.cuda.synchronize()
    t1 = time.perf_counter(); s.run(1000); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: adapt {1e3*(t1-t0):.1f} ms, next 1000 draws {1e3*(t2-t1):.1f} ms")
