"""BASELINE.json configs[4]: stan/arK, synthetic series T = 10 000, K = 5, chains sharded over the ranks with the
pooled windowed-adaptation all-reduce.  Run under torchrun (one rank per GPU); prints throughput and the posterior
means next to the generating coefficients.  Development aid / worked multi-GPU example, not the bench."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import klhr_b200 as kb
from klhr_b200.diagnostics import chain_summary

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
data = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ark10k.json")))
model = kb.BSModel(stan_file="stan/arK.stan", data=data, device=dev)
s = kb.KLHR(model, seed=7, chains=chains, warmup=1000)
t0 = time.time()
s.run(1000)                                  # 4 window closures -> 4 all-reduces
torch.cuda.synchronize()
t_adapt = time.time() - t0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.run(500)
e0.record()
s.run(1000)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
s1, s2 = s.run(200, chain_stats=True)
summ = chain_summary(s1, s2, 200)
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
if rank == 0:
    mean = summ["mean"].cpu().numpy()
    print(f"arK T=10000 K=5: {world} GPU(s) x {chains} chains; adaptation phase {t_adapt:.2f} s; "
          f"{world * chains * 1000 / (ms * 1e-3):.3e} chain-draws/s; acceptance {s.acceptance_probability:.3f}")
    print("posterior means [alpha, beta1..5, log sigma]:", [round(float(x), 4) for x in mean])
    print("generating values                           : [0.2, 0.05, -0.1, 0.15, -0.2, 0.6, %.4f]" % float(torch.log(torch.tensor(0.5))))
    print("closures:", s._windowedadaptation.closures, " _cov:", [float(f"{v:.2e}") for v in s._cov])
if world > 1:
    dist.destroy_process_group()
