"""Device time + accuracy of the tcgen05 attention kernel at the ViT-L shapes (not a pytest file).
usage: DAD_ATT_VARIANT={5,2,3} DAD_ATT_POLY5={0,2,3,4,5} python tests/gpu_attn_time.py [B N heads [scale]]
(scale = std of the random q/k/v; 0.5 gives the no-rescale regime of a random-init model, 1.5 forces rescales)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distill_any_depth_b200 import _lib as L

B, N, heads = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32, 1370, 16)
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
lib = L.load()
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B * N, 3 * heads * 64, device="cuda", generator=g) * scale).bfloat16()
out = torch.empty(B * N, heads * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    L.check(lib.dad_attention(L.ptr(qkv), L.ptr(out), B, N, heads, 0, L.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
it = 20
for _ in range(it):
    L.check(lib.dad_attention(L.ptr(qkv), L.ptr(out), B, N, heads, 0, L.stream_ptr()))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / it
flops = 4.0 * B * heads * N * N * 64
# accuracy on the first image against an fp32 softmax reference (q is pre-scaled in this layout)
q, k, v = qkv[:N].float().view(N, 3, heads, 64).permute(1, 2, 0, 3)
ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).permute(1, 0, 2).reshape(N, heads * 64)
err = (out[:N].float() - ref).abs().max().item() / ref.abs().max().item()
print(f"variant={os.environ.get('DAD_ATT_VARIANT', '5')} poly={os.environ.get('DAD_ATT_POLY5', '-')} scale={scale} B={B} N={N} h={heads}: {ms:.3f} ms  "
      f"{flops / ms / 1e9:.0f} TFLOP/s  max_err/max_ref={err:.2e}", flush=True)
