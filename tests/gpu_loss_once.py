"""ncu / timing target: SSI + fused HDN-DR at the bench shape (B=32, 518x518).  usage: python tests/gpu_loss_once.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import distill_any_depth_b200 as d
from distill_any_depth_b200 import synthetic, losses

B, H = 32, 518
pred, gt, mask = synthetic.make_depth_pair(B, H, H, seed=7)
pred, gt = pred.cuda(), gt.cuda()
full = torch.ones_like(gt, dtype=torch.bool)
for _ in range(3):
    a = losses._ssi(pred, gt, full, False)[0]
    b = losses.hdn_loss_dr(pred, gt, None, 3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    a = losses._ssi(pred, gt, full, False)[0]
e1.record()
torch.cuda.synchronize()
t_ssi = e0.elapsed_time(e1) / 10
e0.record()
for _ in range(10):
    b = losses.hdn_loss_dr(pred, gt, None, 3)
e1.record()
torch.cuda.synchronize()
print(f"ssi {t_ssi:.3f} ms  hdn_dr {e0.elapsed_time(e1) / 10:.3f} ms  values {float(a):.6f} {float(b):.6f}")
