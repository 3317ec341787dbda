"""Development aid: funnel dims=2, KLHRSINH -- posterior of x ~ N(0, 9) with the KL gradient clip on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import klhr_b200 as kb
from klhr_b200.diagnostics import chain_summary
dev = torch.device("cuda", 0)
model = kb.BSModel(stan_file="stan/funnel.stan", data={"D": 1}, device=dev)
for clip in (300.0, 0.0):
    for stride in (None, 10):
        s = kb.KLHRSINH(model, seed=9, chains=8192, warmup=400, overrelaxed=False, pca_stride=stride)
        s._fit.grad_clip = clip
        s.run(400); s.run(600)
        for rep in range(3):
            a0 = s._accept_count.clone()
            s1, s2 = s.run(1500, chain_stats=True)
            summ = chain_summary(s1, s2, 1500)
            acc = float((s._accept_count - a0).double().mean()) / 1500
            print(f"clip={clip} stride={stride} rep={rep} mean0={float(summ['mean'][0]):+.4f} mcse={float(summ['mcse_mean'][0]):.4f} "
                  f"var0={float(summ['var'][0]):.3f} acc={acc:.4f} cov={s._cov}")
