"""Device time of KLHR.sample() (thinned-draw output) for the tile-kernel targets (development aid)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, klhr_b200 as kb
dev = torch.device("cuda", 0)
model = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": 100}, device=dev)
for force in (False, True):
    for thin in (1, 10):
        s = kb.KLHR(model, seed=1, chains=65536, warmup=0)
        s._fit.force_octet = force
        s.run(20)
        M = 41
        out = s.sample(M, thin=thin)                    # first call: lazy module load, allocation
        del out
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = s.sample(M, thin=thin); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"sample(M={M}, thin={thin}) force_octet={force}: {ms:.2f} ms  {65536 * (M - 1) * thin / ms / 1e6:.3f} Gdraws/s")
        del out
