"""ncu target: a few launches of the tcgen05 attention kernel at the ViT-L 518x518 shape (B=8)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distill_any_depth_b200 import _lib as L

B, N, heads = 8, 1370, 16
lib = L.load()
qkv = torch.randn(B * N, 3 * heads * 64, device="cuda").bfloat16()
out = torch.empty(B * N, heads * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    L.check(lib.dad_attention(L.ptr(qkv), L.ptr(out), B, N, heads, 0, L.stream_ptr()))
torch.cuda.synchronize()
print("done")
