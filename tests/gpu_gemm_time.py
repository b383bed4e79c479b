"""Device time of the encoder GEMMs (ViT-L 518^2 B=32 shapes) through dad_gemm_ex (not a pytest file)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distill_any_depth_b200 import _lib as L

lib = L.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 43840
for name, N, K, kind in (("qkv", 3072, 1024, "bias"), ("proj", 1024, 1024, "res"), ("fc1", 4096, 1024, "gelu"), ("fc2", 1024, 4096, "res")):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    gamma = torch.full((N,), 1e-3, device="cuda")
    out32 = torch.zeros(M, N, device="cuda")
    outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def run():
        if kind == "bias":
            L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(outb), 1, 0, M, N, K, 0, L.stream_ptr()))
        elif kind == "gelu":
            L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(outb), 1, 1, M, N, K, 0, L.stream_ptr()))
        else:
            L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gamma), L.ptr(out32), 0, L.ptr(out32), 0, 0, M, N, K, 0,
                                    L.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:5s} M={M} N={N} K={K} {kind:5s}: {ms * 1e3:7.1f} us  {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s", flush=True)
