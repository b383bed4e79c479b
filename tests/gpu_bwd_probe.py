"""GPU probe (not a pytest file): bf16 training path vs the fp32 engine / oracle autograd, and step timings.
usage: python tests/gpu_bwd_probe.py [preset B H W] """
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
import distill_any_depth_b200 as d
from distill_any_depth_b200 import synthetic, _lib

def splitk_check():
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(0)
    for (M, N, K, ks) in [(768, 768, 1570, 4), (32, 576, 20000, 16), (3072, 768, 1570, 2), (96, 384, 333, 3)]:
        lda = (K + 63) // 64 * 64
        A = torch.zeros(M, lda); A[:, :K] = torch.randn(M, K, generator=g)
        W = torch.zeros(N, lda); W[:, :K] = torch.randn(N, K, generator=g)
        Ab, Wb = A.cuda().bfloat16().contiguous(), W.cuda().bfloat16().contiguous()
        out0 = torch.randn(M, N, generator=g).cuda()
        out = out0.clone()
        z, o = torch.zeros(8192, device="cuda"), torch.ones(8192, device="cuda")
        _lib.check(lib.dad_gemm_splitk(_lib.ptr(Ab), _lib.ptr(Wb), _lib.ptr(z), _lib.ptr(o), _lib.ptr(out), M, N, K, lda, ks,
                                       _lib.stream_ptr()), "splitk")
        torch.cuda.synchronize()
        ref = out0 + Ab.float() @ Wb.float().t()
        err = float((out - ref).abs().max() / ref.abs().max())
        print(f"splitk M={M} N={N} K={K} ks={ks}: rel err {err:.2e}", flush=True)

def grads_of(m, x, wd, wf):
    for p in m.parameters(): p.grad = None
    depth, feat = m(x)
    ((depth * wd).sum() + (feat * wf).sum()).backward()
    return depth.detach(), feat.detach(), {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}

def main():
    preset, B, H, W = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("vits", 2, 70, 98)
    splitk_check()
    kw = synthetic.MODEL_PRESETS[preset]
    sd = synthetic.make_state_dict(seed=0, **kw)
    x = synthetic.make_images(B, H, W, seed=77).cuda()
    g = torch.Generator().manual_seed(5)
    D = sd["pretrained.cls_token"].shape[-1]
    wd = torch.randn(B, 1, H, W, generator=g).cuda()
    wf = (torch.randn(B, (H // 14) * (W // 14), D, generator=g) * 0.05).cuda()
    m = d.DepthAnythingV2(**kw); m.load_state_dict(sd, strict=True); m = m.cuda()
    m.precision = "fp32"
    d32, f32, g32 = grads_of(m, x, wd, wf)
    m.precision = "bf16"; m.bf16_backward = True
    d16, f16, g16 = grads_of(m, x, wd, wf)
    print("depth rel err", float((d16 - d32).abs().max() / d32.abs().max()), "feat", float((f16 - f32).abs().max() / f32.abs().max()))
    rows = []
    for k, r in g32.items():
        if r is None:
            assert g16[k] is None; continue
        a = g16[k]
        l2 = float((a - r).norm() / (r.norm() + 1e-30)); mx = float((a - r).abs().max() / (r.abs().max() + 1e-30))
        cos = float((a * r).sum() / (a.norm() * r.norm() + 1e-30))
        rows.append((l2, mx, cos, k))
    rows.sort(reverse=True)
    for r in rows[:12]: print("worst  l2 %.3e  max %.3e  cos %.5f  %s" % r)
    print("median l2 %.3e" % sorted(r[0] for r in rows)[len(rows) // 2], " nan:", sum(1 for r in rows if r[0] != r[0]))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump([dict(l2=r[0], mx=r[1], cos=r[2], k=r[3]) for r in rows], open("gpurun_out/bwd_bf16_probe.json", "w"))
    # how far is PyTorch's own bf16 autocast of the reference graph from its fp32 gradients? (the yardstick for the numbers above)
    if os.environ.get("PROBE_AUTOCAST", "1") == "1" and B * H * W <= 4 * 518 * 518:
        def ograds(autocast):
            leaves = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                dd, ff = oracle.depth_anything_forward(x, leaves, kw["encoder"])
            ((dd.float() * wd).sum() + (ff.float() * wf).sum()).backward()
            return {k: v.grad for k, v in leaves.items()}
        torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
        o32, o16 = ograds(False), ograds(True)
        l2s = sorted((float((o16[k] - o32[k]).norm() / (o32[k].norm() + 1e-30)), k) for k in o32 if o32[k] is not None)
        print("torch autocast(bf16) vs torch fp32 on the oracle graph: median l2 %.3e, worst %.3e (%s)" % (l2s[len(l2s) // 2][0], l2s[-1][0], l2s[-1][1]))
        ours = sorted((float((g32[k] - o32[k]).norm() / (o32[k].norm() + 1e-30)), k) for k in o32 if o32[k] is not None)
        print("our fp32 engine vs torch fp32 (GPU): median l2 %.3e, worst %.3e (%s)" % (ours[len(ours) // 2][0], ours[-1][0], ours[-1][1]))
    # timings
    if os.environ.get("DAD_DEBUG_TIME"):
        m.precision = "bf16"
        print("=== per-launch timing of one bf16 train step", file=sys.stderr, flush=True)
        grads_of(m, x, wd, wf)
        torch.cuda.synchronize()
        return
    import ctypes
    lib = _lib.load()
    names = ["gemm_tc", "gemm_simt", "attention", "layernorm", "elementwise", "loss"]
    for prec in ("fp32", "bf16"):
        m.precision = prec
        for it in range(3):
            torch.cuda.synchronize(); t0 = time.time()
            grads_of(m, x, wd, wf)
            torch.cuda.synchronize(); t1 = time.time()
        print(f"train step (fwd+bwd) {prec}: {1e3 * (t1 - t0):.1f} ms for B={B} {H}x{W} {preset}", flush=True)
        lib.dad_profile_enable(1)
        grads_of(m, x, wd, wf)
        torch.cuda.synchronize()
        for ci, nm in enumerate(names):
            ms, wk, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
            lib.dad_profile_get(ci, ctypes.byref(ms), ctypes.byref(wk), ctypes.byref(n))
            if n.value:
                print(f"   {nm:12s} {ms.value:9.2f} ms  {n.value:5d} launches  work {wk.value:.3e}  -> {wk.value / max(ms.value, 1e-9) / 1e9:.1f} G(flop|B)/s")
        lib.dad_profile_enable(0)

if __name__ == "__main__":
    main()
