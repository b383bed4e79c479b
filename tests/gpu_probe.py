"""Diagnostic probe for a GPU box (not a pytest file): runs each engine in its own subprocess so a
faulting kernel cannot take the others down, and writes detailed error maps / timings to
gpurun_out/probe_*.log.   usage: python tests/gpu_probe.py [step ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def step_gemm():
    import torch
    from test_gpu_kernels import run_gemm
    for (M, N, K) in [(128, 32, 64), (128, 256, 64), (128, 256, 256), (256, 512, 128), (1370, 1024, 1024)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        out = run_gemm(A, W, None, 0)
        ref = A.float() @ W.float().t()
        err = (out - ref).abs()
        print(f"gemm_tc M={M} N={N} K={K}: max_err={err.max().item():.3e} ref_max={ref.abs().max().item():.3f} "
              f"nan={int(torch.isnan(out).sum())}", flush=True)
        if err.max().item() > 1e-2 or torch.isnan(out).any():
            bad = (err > 1e-2) | torch.isnan(out)
            rows = bad.any(1).nonzero().flatten()[:16].tolist()
            cols = bad.any(0).nonzero().flatten()[:32].tolist()
            print("  bad rows (first 16):", rows, "bad cols (first 32):", cols, "frac bad:", bad.float().mean().item())
            # is the output a permutation / K-chunk subset of the reference?  test partial-K hypotheses
            for kk in range(0, K, 16):
                part = A[:, kk:kk + 16].float() @ W[:, kk:kk + 16].float().t()
                print(f"   corr with K-chunk {kk}: {torch.nn.functional.cosine_similarity(out.nan_to_num().flatten(), part.flatten(), dim=0).item():.3f}")
            print("  out[0,:8]", out[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist())


def step_gemm_perf():
    import torch
    from distill_any_depth_b200 import _lib as L
    lib = L.load()
    res = {}
    for (M, N, K) in [(43840, 3072, 1024), (43840, 1024, 1024), (43840, 4096, 1024), (43840, 1024, 4096), (8192, 8192, 8192)]:
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        out = torch.empty(M, N, device="cuda")
        for _ in range(3):
            L.check(lib.dad_gemm(L.ptr(A), L.ptr(W), None, L.ptr(out), M, N, K, 0, L.stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            L.check(lib.dad_gemm(L.ptr(A), L.ptr(W), None, L.ptr(out), M, N, K, 0, L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * M * N * K / ms / 1e9
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        Af, Wf = A, W
        for _ in range(3):
            torch.matmul(Af, Wf.t())
        t0.record()
        for _ in range(10):
            torch.matmul(Af, Wf.t())
        t1.record()
        torch.cuda.synchronize()
        ms_ref = t0.elapsed_time(t1) / 10
        res[f"{M}x{N}x{K}"] = dict(ms=ms, tflops=tf, cublas_ms=ms_ref, cublas_tflops=2.0 * M * N * K / ms_ref / 1e9)
        print(f"gemm_tc {M}x{N}x{K}: {ms:.3f} ms {tf:.0f} TFLOP/s (fp32 out) | cuBLAS bf16 {ms_ref:.3f} ms "
              f"{2.0 * M * N * K / ms_ref / 1e9:.0f} TFLOP/s", flush=True)
    json.dump(res, open(os.path.join(OUT, "probe_gemm_perf.json"), "w"), indent=1)


def step_forward_perf():
    import torch
    import distill_any_depth_b200 as d
    from distill_any_depth_b200 import synthetic
    res = {}
    for preset, B, H in (("vitl", 8, 518), ("vitl", 32, 518), ("vitb", 16, 392)):
        kw = synthetic.MODEL_PRESETS[preset]
        m = d.DepthAnythingV2(**kw).cuda()
        x = torch.randn(B, 3, H, H, device="cuda")
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[f"{preset}_B{B}_{H}"] = dict(ms=ms, img_s=B / ms * 1e3)
        print(f"forward {preset} B={B} {H}x{H}: {ms:.2f} ms  {B / ms * 1e3:.1f} img/s", flush=True)
        del m
    json.dump(res, open(os.path.join(OUT, "probe_forward_perf.json"), "w"), indent=1)


def step_gemm_kinds():
    """Throughput of the fused encoder epilogues at ViT-L B=32 shapes."""
    import torch
    from distill_any_depth_b200 import _lib as L
    lib = L.load()
    M = 43840
    cases = [("qkv bias->bf16", 3072, 1024, "bias"), ("fc1 bias+gelu->bf16", 4096, 1024, "gelu"),
             ("proj res fp32", 1024, 1024, "res"), ("fc2 res fp32", 1024, 4096, "res")]
    for name, N, K, kind in cases:
        A = torch.randn(M, K, device="cuda").bfloat16()
        W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda"); gamma = torch.ones(N, device="cuda")
        outb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        outf = torch.zeros(M, N, device="cuda")
        def run():
            if kind == "bias":
                L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(outb), 1, 0, M, N, K, 0, L.stream_ptr()))
            elif kind == "gelu":
                L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(outb), 1, 1, M, N, K, 0, L.stream_ptr()))
            else:
                L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gamma), L.ptr(outf), 0, L.ptr(outf), 0, 0, M, N, K, 0, L.stream_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:24s} M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)


def step_conv_perf():
    """Implicit-GEMM conv throughput at the decoder shapes of ViT-L 518x518 (B=32); fp32 out + bias (generic kind)."""
    import torch
    from distill_any_depth_b200 import _lib as L
    from test_gpu_kernels import pack_conv_weight
    lib = L.load()
    cases = [("refinenet1 rcu 148^2 256->256", 32, 148, 148, 256, 256, 9), ("refinenet2 rcu 74^2", 32, 74, 74, 256, 256, 9),
             ("output_conv1 296^2 256->128", 32, 296, 296, 256, 128, 9), ("head 518^2 128->32", 32, 518, 518, 128, 32, 9),
             ("layer3_rn 37^2 1024->256", 32, 37, 37, 1024, 256, 9), ("out_conv 1x1 148^2", 32, 148, 148, 256, 256, 1),
             ("layer1_rn 148^2 256->256", 32, 148, 148, 256, 256, 9)]
    for name, B, H, W, C, Co, taps in cases:
        k = 3 if taps == 9 else 1
        x = torch.randn(B, H, W, C, device="cuda").bfloat16()
        w = pack_conv_weight(torch.randn(Co, C, k, k, device="cuda") * 0.05, torch.bfloat16)
        bias = torch.randn(Co, device="cuda")
        out = torch.empty(B, H, W, Co, device="cuda")
        run = lambda: L.check(lib.dad_conv_nhwc(L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(out), B, H, W, C, Co, taps, 0, L.stream_ptr()))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * B * H * W * Co * C * taps
        print(f"{name:32s} {ms:7.3f} ms  {fl / ms / 1e9:6.0f} TFLOP/s", flush=True)


STEPS = {"conv_perf": step_conv_perf, "gemm_kinds": step_gemm_kinds, "gemm": step_gemm, "gemm_perf": step_gemm_perf, "forward_perf": step_forward_perf}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        STEPS[sys.argv[2]]()
        sys.exit(0)
    steps = sys.argv[1:] or list(STEPS)
    for s in steps:
        t = time.time()
        log = os.path.join(OUT, f"probe_{s}.log")
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", s], stdout=f,
                                   stderr=subprocess.STDOUT, timeout=600)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        print(f"[probe] {s}: rc={rc} {time.time() - t:.1f}s", flush=True)
        print(open(log).read()[-3000:], flush=True)
