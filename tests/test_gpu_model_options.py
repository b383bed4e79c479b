"""GPU parity of the model OPTIONS: the ``use_clstoken`` readout (dpt.py:116-122, 153-156), the
ViT-g / SwiGLU encoder (dinov2.py:381-395, dinov2_layers/swiglu_ffn.py) and ``use_bn`` (util/blocks.py:49-51, eval mode),
forward-only, through the same front-end and C ABI as the main path.  Same tolerances as test_gpu_model.py: relative depth error <= 1e-4 (fp32 engine) / 2e-2 (bf16)."""
import pytest
import torch

import oracle
from distill_any_depth_b200 import synthetic
from helpers import rel_depth_err, sub
from oracle.make_golden_options import OPTION_CASES
from test_gpu_model import report

pytestmark = pytest.mark.gpu


def build(kw, seed, head_bias=0.25):
    import distill_any_depth_b200 as d
    sd = synthetic.make_state_dict(seed=seed, head_bias=head_bias, **kw)
    m = d.DepthAnythingV2(**kw)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("case", OPTION_CASES, ids=lambda c: c[0])
def test_option_forward_matches_reference_fixture(case, golden_model_options):
    """Both engines against the live reference's recorded outputs (one model build per case: ViT-g is 1.26e9 parameters)."""
    name, kw, B, H, W, ws, xs, hb = case
    g = golden_model_options
    m, _ = build(kw, ws, hb)
    x = synthetic.make_images(B, H, W, seed=xs).cuda()
    d_ref = torch.from_numpy(g[name + "/depth_sub"])
    f_ref = torch.from_numpy(g[name + "/feat_sub"])
    den = d_ref.abs().clamp(min=0.1 * float(g[name + "/depth_stats"][1]))
    fmax = max(1.0, float(g[name + "/feat_stats"][1]))
    for precision, tol, ftol in (("fp32", 1e-4, 1e-3), ("bf16", 2e-2, 6e-2)):
        m.precision = precision
        with torch.no_grad():
            depth, feat = m(x)
        torch.cuda.synchronize()
        assert depth.shape == (B, 1, H, W)
        assert feat.shape == (B, (H // 14) * (W // 14), oracle.VIT_CONFIGS[kw["encoder"]]["embed_dim"])
        rel = ((sub(depth.cpu()) - d_ref).abs() / den).max().item()
        f_err = (sub(feat.cpu()) - f_ref).abs().max().item()
        report(f"{name}/{precision}", dict(rel_depth=rel, feat_abs=f_err))
        assert rel <= tol, (precision, rel)
        assert f_err <= ftol * fmax, (precision, f_err)


def test_clstoken_full_maps_and_batch_rows():
    """Full-map comparison with the oracle at a size where every readout GEMM has several row tiles and the two images'
    class tokens differ (the concat must pick row b's token for image b)."""
    kw = dict(synthetic.MODEL_PRESETS["vits"], use_clstoken=True)
    m, sd = build(kw, 6)
    x = synthetic.make_images(3, 154, 210, seed=78)
    with torch.no_grad():
        d_ref, f_ref = oracle.depth_anything_forward(x, sd, "vits")
        for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
            m.precision = precision
            depth, feat = m(x.cuda())
            rel = rel_depth_err(depth.cpu(), d_ref).max().item()
            report(f"vits_clstoken_154x210/{precision}", dict(rel_depth=rel))
            assert rel <= tol, (precision, rel)
        # the readout really is in the path: dropping it changes the map
        plain, _ = build(synthetic.MODEL_PRESETS["vits"], 6)
        plain.precision = "fp32"
        m.precision = "fp32"
        assert (plain(x.cuda())[0] - m(x.cuda())[0]).abs().max().item() > 1e-3


def test_vitg_giant_head_engines_agree():
    """ViT-g with the Depth-Anything-V2 giant head (features 384, out_channels 4 x 1536) at 8 x 70 x 98 (288 token rows:
    the 2-CTA GEMM kernel takes the encoder GEMMs): the tcgen05 path against the fp32 FFMA engine, and the fp32 engine
    against the oracle on the first two images (images are independent)."""
    kw = synthetic.MODEL_PRESETS["vitg"]
    m, sd = build(kw, 7)
    x = synthetic.make_images(8, 70, 98, seed=79)
    with torch.no_grad():
        m.precision = "fp32"
        d32, f32 = m(x.cuda())
        m.precision = "bf16"
        d16, f16 = m(x.cuda())
        d_ref, f_ref = oracle.depth_anything_forward(x[:2], sd, "vitg")
    rel32 = rel_depth_err(d32[:2].cpu(), d_ref).max().item()
    rel16 = rel_depth_err(d16.cpu(), d32.cpu()).max().item()
    f_err = ((f32[:2].cpu() - f_ref).abs().max() / f_ref.abs().max()).item()
    report("vitg_giant_70x98", dict(rel_fp32_vs_oracle=rel32, rel_bf16_vs_fp32=rel16, feat_rel=f_err))
    assert rel32 <= 1e-4, rel32
    assert f_err <= 1e-4, f_err
    assert rel16 <= 2e-2, rel16


def test_use_bn_tracks_running_statistics_updates():
    """The folded convolutions are rebuilt when a BatchNorm BUFFER changes (buffers are part of the weight signature)."""
    kw = dict(synthetic.MODEL_PRESETS["vits"], use_bn=True)
    m, sd = build(kw, 8)
    m.precision = "fp32"
    x = synthetic.make_images(1, 70, 98, seed=1240)
    with torch.no_grad():
        d0, _ = m(x.cuda())
        bn = m.depth_head.scratch.refinenet1.resConfUnit2.bn2
        bn.running_mean.add_(0.5)
        d1, _ = m(x.cuda())
        sd2 = dict(sd)
        sd2["depth_head.scratch.refinenet1.resConfUnit2.bn2.running_mean"] = bn.running_mean.cpu().clone()
        d_ref, _ = oracle.depth_anything_forward(x, sd2, "vits")
    assert (d1 - d0).abs().max().item() > 1e-3
    assert rel_depth_err(d1.cpu(), d_ref).max().item() <= 1e-4
    m.train()
    with pytest.raises(NotImplementedError):
        m(x.cuda())


def test_options_are_forward_only():
    """The training backward covers the Mlp encoder without the readout: asking for gradients must fail loudly."""
    kw = dict(synthetic.MODEL_PRESETS["vits"], use_clstoken=True)
    m, _ = build(kw, 6)
    m.precision = "fp32"
    x = synthetic.make_images(1, 70, 70, seed=80).cuda()
    with pytest.raises(NotImplementedError):
        m(x)
    with torch.no_grad():
        depth, _ = m(x)
    assert not depth.requires_grad
