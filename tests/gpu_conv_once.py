"""ncu target: implicit-GEMM conv launches at a decoder shape (refinenet1 RCU conv, B=8)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from distill_any_depth_b200 import _lib as L
from test_gpu_kernels import pack_conv_weight

B, H, W, C, Co, taps = (int(v) for v in sys.argv[1:7]) if len(sys.argv) >= 7 else (8, 148, 148, 256, 256, 9)
lib = L.load()
x = torch.randn(B, H, W, C, device="cuda").bfloat16()
w = pack_conv_weight(torch.randn(Co, C, 3, 3, device="cuda") * 0.05, torch.bfloat16)
bias = torch.randn(Co, device="cuda")
out = torch.empty(B, H, W, Co, device="cuda")
for _ in range(4):
    L.check(lib.dad_conv_nhwc(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(out), B, H, W, C, Co, taps, 0, L.stream_ptr()))
torch.cuda.synchronize()
print("done")
