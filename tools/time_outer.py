import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, klhr_b200 as kb
dev = torch.device("cuda", 0)
for B, D in ((65536, 100), (16384, 256), (262144, 11)):
    th = torch.randn(B, D, dtype=torch.float64, device=dev)
    sh = torch.zeros(D, dtype=torch.float64, device=dev)
    outer = torch.zeros(D, D, dtype=torch.float64, device=dev); s1 = torch.zeros(D, dtype=torch.float64, device=dev)
    scr = kb.outer_scratch(th)
    for _ in range(3): kb.outer_accumulate(th, sh, outer, s1, scratch=scr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): kb.outer_accumulate(th, sh, outer, s1, scratch=scr)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"outer_accumulate B={B} D={D}: {ms*1e3:.1f} us  {2*B*D*D/2/ms/1e9:.2f} TFLOP/s (upper triangle)")
