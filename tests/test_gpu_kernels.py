"""GPU parity of the individual engines through the C ABI: tcgen05 GEMM / implicit-GEMM conv,
FFMA verification engine, attention (bf16 tensor-core and fp32).  References are plain PyTorch
fp32 ops on the same (bf16-rounded where applicable) inputs."""
import ctypes
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from distill_any_depth_b200 import _lib as L
    return L


def run_gemm(A, W, bias, mode):
    L = _lib()
    lib = L.load()
    M, K = A.shape
    N = W.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda")
    L.check(lib.dad_gemm(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(out), M, N, K, mode, L.stream_ptr()), "dad_gemm")
    torch.cuda.synchronize()
    return out


GEMM_SHAPES = [(128, 256, 64), (256, 256, 128), (1370, 1024, 1024), (1000, 384, 640), (2740, 3072, 1024),
               (333, 48, 768), (1369, 96, 768), (4096, 4096, 512), (130, 32, 64), (785, 1152, 384)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tc_matches_fp32_matmul(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = run_gemm(A, W, bias, 0)
    ref = A.float() @ W.float().t() + bias
    err = (out - ref).abs().max().item()
    assert torch.isfinite(out).all(), "unwritten / non-finite outputs"
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 384, 640), (777, 48, 100)])
def test_gemm_simt_matches_fp32_matmul(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = run_gemm(A, W, bias, 1)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    assert (out - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


# the last three shapes have more 256 x 256 tiles than SM pairs with a short partial wave: 75 tiles (tail 1 -> four 64-wide
# items), 104 tiles (tail 30 -> 128-wide halves), 2 x 80 tiles (tail 12 -> quarters) - the tail-splitting schedule of gemm_tc2
@pytest.mark.parametrize("M,N,K", [(1370, 1024, 1024), (2740, 3072, 1024), (785, 384, 384), (1000, 768, 3072),
                                   (19200, 256, 128), (26500, 256, 192), (20480, 512, 64)])
@pytest.mark.parametrize("kind", ["bias_bf16", "gelu_bf16", "res_f32", "res_bf16_relu"])
def test_gemm_tc_fused_epilogues(M, N, K, kind):
    """The encoder epilogues (compile-time specialised) and a generic run-time one."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(N, device="cuda", generator=g)
    acc = A.float() @ W.float().t() + bias
    if kind == "bias_bf16":
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(out), 1, 0, M, N, K, 0, L.stream_ptr()))
        ref, tol = acc, 1e-2
    elif kind == "gelu_bf16":
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(out), 1, 1, M, N, K, 0, L.stream_ptr()))
        ref, tol = F.gelu(acc), 1e-2
    elif kind == "res_f32":
        res = torch.randn(M, N, device="cuda", generator=g)
        out = res.clone()  # in-place residual update, as the encoder does
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), L.ptr(gamma), L.ptr(out), 0, L.ptr(out), 0, 0, M, N, K, 0, L.stream_ptr()))
        ref, tol = res + gamma * acc, 2e-3
    else:
        res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        L.check(lib.dad_gemm_ex(L.ptr(A), L.ptr(W), L.ptr(bias), None, L.ptr(res), 1, L.ptr(out), 1, 2, M, N, K, 0, L.stream_ptr()))
        ref, tol = res.float() + torch.relu(acc), 1e-2
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= tol * max(1.0, ref.abs().max().item()), err


def pack_conv_weight(w, dtype):
    """[Co, Ci, kh, kw] -> [Co, taps * Cp] (tap-major, channels zero-padded to a multiple of 64)."""
    Co, Ci, kh, kw = w.shape
    Cp = (Ci + 63) // 64 * 64
    p = torch.zeros(Co, kh * kw, Cp, device=w.device)
    p[:, :, :Ci] = w.permute(0, 2, 3, 1).reshape(Co, kh * kw, Ci)
    return p.reshape(Co, -1).to(dtype).contiguous()


CONV_SHAPES = [(2, 19, 37, 64, 64, 9), (1, 37, 37, 256, 256, 9), (2, 24, 40, 96, 48, 9), (1, 74, 74, 128, 32, 9),
               (3, 10, 10, 192, 64, 1), (1, 148, 148, 64, 128, 9), (2, 8, 16, 48, 64, 9),
               # 2-CTA halo kernel (conv_tc2.cu): odd tile counts, channel tails, several chunks, N = 128 / 256 / 512
               (3, 19, 19, 256, 256, 9), (1, 8, 8, 72, 128, 9), (2, 33, 50, 136, 256, 9), (1, 74, 74, 512, 256, 9),
               (5, 16, 8, 64, 512, 9)]


@pytest.mark.parametrize("B,H,W,C,Co,taps", CONV_SHAPES)
@pytest.mark.parametrize("mode", [0, 1])
def test_conv_nhwc_matches_conv2d(B, H, W, C, Co, taps, mode):
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    k = 3 if taps == 9 else 1
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    w = torch.randn(Co, C, k, k, device="cuda", generator=g) * 0.05
    bias = torch.randn(Co, device="cuda", generator=g)
    dt = torch.bfloat16 if mode == 0 else torch.float32
    xq, wq = x.to(dt), w.to(dt)
    x_nhwc = xq.permute(0, 2, 3, 1).contiguous()
    wp = pack_conv_weight(wq.float(), dt)
    out = torch.full((B, H, W, Co), float("nan"), device="cuda")
    L.check(lib.dad_conv_nhwc(L.ptr(x_nhwc), L.ptr(wp), L.ptr(bias), L.ptr(out), B, H, W, C, Co, taps, mode,
                              L.stream_ptr()), "dad_conv_nhwc")
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(xq.double(), wq.double(), bias.double(), padding=k // 2).float().permute(0, 2, 3, 1)
    assert torch.isfinite(out).all(), "unwritten / non-finite outputs"
    tol = 2e-3 if mode == 0 else 2e-5
    assert (out - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,C,Co", [(2, 37, 37, 128, 128), (1, 28, 28, 768, 768), (3, 9, 14, 64, 256), (1, 5, 5, 384, 64)])
def test_conv_stride2_matches_conv2d(B, H, W, C, Co):
    """resize_layers[3] (dpt.py:101-106): 3x3 stride-2 pad-1 convolution as an implicit GEMM (TMA element strides)."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    x = torch.randn(B, C, H, W, device="cuda", generator=g)
    w = torch.randn(Co, C, 3, 3, device="cuda", generator=g) * 0.05
    bias = torch.randn(Co, device="cuda", generator=g)
    xq, wq = x.bfloat16(), w.bfloat16()
    x_nhwc = xq.permute(0, 2, 3, 1).contiguous()
    wp = pack_conv_weight(wq.float(), torch.bfloat16)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.full((B, Ho, Wo, Co), float("nan"), device="cuda")
    L.check(lib.dad_conv_nhwc_ex(L.ptr(x_nhwc), L.ptr(wp), L.ptr(bias), L.ptr(out), B, H, W, C, Co, 9, 2, 0,
                                 L.stream_ptr()), "dad_conv_nhwc_ex")
    torch.cuda.synchronize()
    ref = F.conv2d(xq.double(), wq.double(), bias.double(), stride=2, padding=1).float().permute(0, 2, 3, 1)
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())


# 5 = attention_tc5.cu (the default), 2 / 3 = the round-1 kernels kept for A/B measurements
@pytest.mark.parametrize("M,N,K", [(128, 128, 4096), (768, 2304, 1570), (96, 768, 3140), (200, 72, 777)])
def test_gemm_splitk_mn_major_operands(M, N, K):
    """dW = dY^T X read straight from dY [K, M] and X [K, N] (padded row pitches): MN-major UMMA operands, no transposes."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    lda, ldw = (M + 15) // 8 * 8, (N + 23) // 8 * 8
    A = torch.randn(K, lda, device="cuda", generator=g).bfloat16()
    W = torch.randn(K, ldw, device="cuda", generator=g).bfloat16()
    out0 = torch.randn(M, N, device="cuda", generator=g)
    out = out0.clone()
    zeros, ones = torch.zeros(16384, device="cuda"), torch.ones(16384, device="cuda")
    L.check(lib.dad_gemm_splitk_mn(L.ptr(A), L.ptr(W), L.ptr(zeros), L.ptr(ones), L.ptr(out), M, N, K, lda, ldw, 3,
                                   L.stream_ptr()), "dad_gemm_splitk_mn")
    torch.cuda.synchronize()
    ref = out0.double() + A[:, :M].double().t() @ W[:, :N].double()
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,W,Co,Ci", [(2, 16, 16, 128, 128), (2, 37, 29, 64, 96), (1, 9, 70, 256, 32), (3, 8, 8, 32, 200)])
def test_conv_wgrad_from_nhwc_matches_autograd(B, H, W, Co, Ci):
    """3x3 / stride 1 / padding 1 weight gradient from the NHWC tensors (no im2col, no transposes) against autograd of
    torch.nn.functional.conv2d in float64; ragged patches, channel counts below and between the 64 / 128 tile sizes."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(B * H + Ci)
    X = torch.randn(B, H, W, Ci, device="cuda", generator=g).bfloat16()
    dY = torch.randn(B, H, W, Co, device="cuda", generator=g).bfloat16()
    CiP = (Ci + 127) // 128 * 128
    out = torch.zeros(Co, 9 * CiP, device="cuda")
    zeros, ones = torch.zeros(16384, device="cuda"), torch.ones(16384, device="cuda")
    L.check(lib.dad_conv_wgrad(L.ptr(dY), L.ptr(X), L.ptr(zeros), L.ptr(ones), L.ptr(out), B, H, W, Co, Ci, 2, L.stream_ptr()),
            "dad_conv_wgrad")
    torch.cuda.synchronize()
    w = torch.zeros(Co, Ci, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(X.double().permute(0, 3, 1, 2), w, padding=1)
    y.backward(dY.double().permute(0, 3, 1, 2))
    got = out.view(Co, 9, CiP)[:, :, :Ci].permute(0, 2, 1).reshape(Co, Ci, 3, 3).double()
    assert (got - w.grad).abs().max().item() <= 1e-4 * w.grad.abs().max().item()
    assert out.view(Co, 9, CiP)[:, :, Ci:].abs().max().item() == 0 if CiP > Ci else True


@pytest.mark.parametrize("M,rows,K,ld,offs", [(64, 64, 5000, 128, [0, 64, -64]), (256, 200, 20000, 256, [-232, -232, -232, 0, 0, 0, 232, 232, 232]),
                                              (128, 128, 3000, 128, [8, -8, 16, -4096])])
def test_gemm_shifted_views_match_torch(M, rows, K, ld, offs):
    """The 3x3 weight gradient's GEMM (train.inl conv_wgrad): out[m, t*ld + r] += sum_k A[m, k] W[r, k + off_t], zero outside
    the matrix, rows >= `rows` of every view zero.  Offsets are multiples of 8 elements (TMA box starts are 16-byte aligned)."""
    import ctypes
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(5)
    Kp = (K + 63) // 64 * 64
    A = torch.zeros(M, Kp, device="cuda", dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = torch.zeros(rows, Kp, device="cuda", dtype=torch.bfloat16)
    W[:, :K] = torch.randn(rows, K, device="cuda", generator=g).bfloat16()
    taps = len(offs)
    out = torch.zeros(M, taps * ld, device="cuda")
    zeros, ones = torch.zeros(16384, device="cuda"), torch.ones(16384, device="cuda")
    L.check(lib.dad_gemm_shifted(L.ptr(A), L.ptr(W), L.ptr(zeros), L.ptr(ones), L.ptr(out), M, rows, K, Kp, taps, ld,
                                 (ctypes.c_int * taps)(*offs), 4, L.stream_ptr()), "dad_gemm_shifted")
    torch.cuda.synchronize()
    ref = torch.zeros(M, taps * ld, device="cuda", dtype=torch.float64)
    Ad, Wd = A.double(), W.double()
    for t, o in enumerate(offs):
        Ws = torch.zeros_like(Wd)
        if 0 <= o < Kp:
            Ws[:, :Kp - o] = Wd[:, o:]
        elif -Kp < o < 0:
            Ws[:, -o:] = Wd[:, :Kp + o]
        ref[:, t * ld:t * ld + rows] = Ad @ Ws.t()
    assert (out.double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    with pytest.raises(ValueError):   # an odd offset is rejected on the host instead of faulting in TMA
        L.check(lib.dad_gemm_shifted(L.ptr(A), L.ptr(W), L.ptr(zeros), L.ptr(ones), L.ptr(out), M, rows, K, Kp, 1, ld,
                                     (ctypes.c_int * 1)(1), 4, L.stream_ptr()), "dad_gemm_shifted")


@pytest.mark.parametrize("rows,D,period", [(2740, 1024, None), (4 * 1369, 1024, (1369, 1370, 1)), (2049, 768, None),
                                            (2 * 1024, 384, (1024, 1025, 1)), (2100, 1536, None), (100, 1024, None)])
@pytest.mark.parametrize("mode", [0, 1])
def test_layernorm_kernel(rows, D, period, mode):
    """dinov2.py:213-214 / :304-305 (torch.nn.LayerNorm): torch.layer_norm within fp32 / bf16 rounding; the period arguments
    drop the class-token row of every image, as the shared final norm of the intermediate layers does."""
    L = _lib()
    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    out_period, in_period, in_offset = period if period else (rows, rows, 0)
    in_rows = (rows // out_period) * in_period if period else rows
    x = torch.randn(in_rows, D, device="cuda", generator=g) * 3 + torch.randn(in_rows, 1, device="cuda", generator=g)
    w = torch.randn(D, device="cuda", generator=g)
    b = torch.randn(D, device="cuda", generator=g)
    out = torch.full((rows, D), float("nan"), device="cuda", dtype=torch.bfloat16 if mode == 0 else torch.float32)
    out32 = torch.full((rows, D), float("nan"), device="cuda")
    L.check(lib.dad_layernorm(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(out), L.ptr(out32), rows, D, out_period, in_period,
                              in_offset, 1e-6, mode, L.stream_ptr()), "dad_layernorm")
    torch.cuda.synchronize()
    xs = x.view(-1, in_period, D)[:, in_offset:in_offset + out_period].reshape(rows, D) if period else x
    ref = torch.nn.functional.layer_norm(xs, (D,), w, b, 1e-6)
    assert (out32 - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    tol = 8e-3 if mode == 0 else 2e-5
    assert (out.float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


_ATT_VARIANTS = {5: "tc5", 6: "tc6_three_buffers", 7: "tc7_two_groups", 8: "tc7_split_issuers", 2: "pipelined2cta", 3: "serial4cta"}


@pytest.fixture(params=list(_ATT_VARIANTS), ids=list(_ATT_VARIANTS.values()))
def att_variant(request, monkeypatch):
    """All tcgen05 attention kernels (attention_tc5.cu / attention_tc.cu / attention_tc3.cu) go through the same tests."""
    monkeypatch.setenv("DAD_ATT_VARIANT", str(request.param))
    return request.param


@pytest.mark.parametrize("B,N,heads", [(1, 64, 1), (2, 785, 6), (1, 1370, 16), (1, 26, 2), (1, 200, 3), (1, 5477, 16), (2, 5477, 16)])
@pytest.mark.parametrize("mode", [0, 1])
def test_attention_matches_softmax_reference(B, N, heads, mode, att_variant):
    L = _lib()
    lib = L.load()
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(N + heads)
    qkv = torch.randn(B * N, 3 * D, device="cuda", generator=g)
    qkv[:, :D] *= 0.5  # q arrives pre-scaled from the qkv GEMM; any values are fine for the kernel contract
    dt = torch.bfloat16 if mode == 0 else torch.float32
    qq = qkv.to(dt).contiguous()
    out = torch.full((B * N, D), float("nan"), device="cuda", dtype=dt)
    L.check(lib.dad_attention(L.ptr(qq), L.ptr(out), B, N, heads, mode, L.stream_ptr()), "dad_attention")
    torch.cuda.synchronize()
    r = qq.double().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((r[0] @ r[1].transpose(-2, -1)).softmax(-1) @ r[2]).transpose(1, 2).reshape(B * N, D).float()
    assert torch.isfinite(out.float()).all()
    tol = 2e-2 if mode == 0 else 2e-5
    assert (out.float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())


def test_attention_lazy_rescale_path(att_variant):
    """Late key tiles with much larger logits force the reference-maximum rescale of O / l (TMEM round trip)."""
    L = _lib()
    lib = L.load()
    B, N, heads = 2, 600, 2
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(B, N, 3, heads, 64, device="cuda", generator=g)
    qkv[:, :, 0] *= 0.7
    qkv[:, 300:420, 1] *= 3.0    # logits jump by >> 2^8 in the 3rd / 4th key tile for many rows
    qkv[:, 520:, 1] *= 6.0       # and again in the last (partial) tile
    qq = qkv.reshape(B * N, 3 * D).bfloat16().contiguous()
    out = torch.full((B * N, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(lib.dad_attention(L.ptr(qq), L.ptr(out), B, N, heads, 0, L.stream_ptr()), "dad_attention")
    torch.cuda.synchronize()
    r = qq.double().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((r[0] @ r[1].transpose(-2, -1)).softmax(-1) @ r[2]).transpose(1, 2).reshape(B * N, D).float()
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 3e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("variant", [5, 6, 7, 8])
@pytest.mark.parametrize("poly", [0, 2, 3, 4, 5])
def test_attention_tc5_overflow_falls_back_to_exact_passes(poly, variant, monkeypatch):
    """attention_tc5 exponentiates against the FIRST key tile's row maximum; later logits more than 127 log2 units above
    it make the tensor-core row sum non-finite, and the CTA must then redo its work in-kernel (max pass + exact pass).
    Rows that overflow and rows that do not share CTAs here; every FMA-pipe / MUFU split is exercised."""
    monkeypatch.setenv("DAD_ATT_VARIANT", str(variant))
    monkeypatch.setenv("DAD_ATT_POLY5", str(poly))
    L = _lib()
    lib = L.load()
    B, N, heads = 2, 700, 2
    D = heads * 64
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn(B, N, 3, heads, 64, device="cuda", generator=g)
    qkv[:, :, 0] *= 0.7
    qkv[:, 200:330, 1] *= 14.0   # logit std ~ 8 * 0.7 * 14 = 78 (113 log2 units): far beyond the first tile's maximum
    qkv[0, :, 0, 1] *= 0.05      # image 0 / head 1: small logits everywhere -> no overflow in those CTAs
    qq = qkv.reshape(B * N, 3 * D).bfloat16().contiguous()
    out = torch.full((B * N, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(lib.dad_attention(L.ptr(qq), L.ptr(out), B, N, heads, 0, L.stream_ptr()), "dad_attention")
    torch.cuda.synchronize()
    r = qq.double().reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = r[0] @ r[1].transpose(-2, -1)
    assert ((s[..., 64:].max(-1).values - s[..., :64].max(-1).values) * 1.4427 > 130).any(), "test does not overflow"
    ref = (s.softmax(-1) @ r[2]).transpose(1, 2).reshape(B * N, D).float()
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 3e-2 * max(1.0, ref.abs().max().item())
