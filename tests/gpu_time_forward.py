"""Per-launch device-time listing of one forward (DAD_DEBUG_TIME=1 serialises launches and prints each with its label).
usage: python tests/gpu_time_forward.py [preset B H]"""
import os
import sys
os.environ["DAD_DEBUG_TIME"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import distill_any_depth_b200 as d
from distill_any_depth_b200 import synthetic
preset, B, H = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("vitl", 32, 518)
m = d.DepthAnythingV2(**synthetic.MODEL_PRESETS[preset]).cuda()
x = torch.randn(B, 3, H, H, device="cuda")
m(x)
torch.cuda.synchronize()
print("=== timed forward", file=sys.stderr, flush=True)
m(x)
torch.cuda.synchronize()
