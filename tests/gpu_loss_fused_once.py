"""One fused SSI + HDN-DR loss evaluation at the benchmark's shape (not a pytest file): for ncu launch lists.
usage: python tests/gpu_loss_fused_once.py [B H W]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import distill_any_depth_b200 as d
from distill_any_depth_b200 import synthetic

B, H, W = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32, 518, 518)
pred, gt, _ = synthetic.make_depth_pair(B, H, W, seed=7)
P, G = pred.cuda(), gt.cuda()
for _ in range(3):
    ssi, hdn = d.ssi_hdn_dr(P, G, None, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ssi, hdn = d.ssi_hdn_dr(P, G, None, 3)
e1.record()
torch.cuda.synchronize()
print(f"fused SSI + HDN-DR B={B} {H}x{W}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per evaluation; ssi={float(ssi):.7f} hdn={float(hdn):.7f}")
