"""ncu target: a few launches of the tcgen05 GEMM engine on encoder-shaped problems.
usage: python tests/gpu_gemm_once.py [M N K]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distill_any_depth_b200 import _lib as L

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (43840, 3072, 1024)
kind = sys.argv[4] if len(sys.argv) > 4 else "plain"
lib = L.load()
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
out = torch.empty(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
gamma = torch.ones(N, device="cuda")
outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    if kind == "plain":
        L.check(lib.dad_gemm(L.ptr(A), L.ptr(W), None, L.ptr(out), M, N, K, 0, L.stream_ptr()))
    elif kind == "bias":
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(outb), 1, 0, M, N, K, 0, L.stream_ptr()))
    elif kind == "gelu":
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(outb), 1, 1, M, N, K, 0, L.stream_ptr()))
    else:
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gamma), L.ptr(out), 0, L.ptr(out), 0, 0, M, N, K, 0, L.stream_ptr()))
torch.cuda.synchronize()
print("done", M, N, K)
