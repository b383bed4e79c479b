"""GPU parity of the whole forward (DepthAnythingV2 / DepthAnything front-ends -> C ABI) against
the CPU oracle and the committed reference fixtures.
Tolerances (north_star): per-pixel relative depth error <= 2e-2 in bf16 mode, <= 1e-4 in the
fp32 verification mode (helpers.rel_depth_err documents the denominator floor)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from distill_any_depth_b200 import synthetic
from helpers import rel_depth_err, sub
from oracle.make_golden import MODEL_CASES

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def build(preset, seed, teacher=False, head_bias=0.25):
    import distill_any_depth_b200 as d
    kw = synthetic.MODEL_PRESETS[preset]
    sd = synthetic.make_state_dict(seed=seed, head_bias=head_bias, **kw)
    if teacher:
        from oracle.make_golden import student_to_teacher_keys
        m = d.DepthAnything(**kw)
        m.load_state_dict(student_to_teacher_keys(sd), strict=True)
    else:
        m = d.DepthAnythingV2(**kw)
        m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd, kw


def report(name, payload):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity_model.jsonl"), "a") as f:
        f.write(json.dumps(dict(case=name, **payload)) + "\n")


@pytest.mark.parametrize("case", MODEL_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_matches_reference_fixture(case, precision, golden_model):
    name, preset, B, H, W, ws, xs, teacher, hb = case
    m, sd, kw = build(preset, ws, teacher, hb)
    m.precision = precision
    x = synthetic.make_images(B, H, W, seed=xs)
    with torch.no_grad():
        depth, feat = m(x.cuda())
    torch.cuda.synchronize()
    assert depth.shape == (B, 1, H, W)
    assert feat.shape == (B, (H // 14) * (W // 14), oracle.VIT_CONFIGS[kw["encoder"]]["embed_dim"])
    d_ref = torch.from_numpy(golden_model[name + "/depth_sub"])
    f_ref = torch.from_numpy(golden_model[name + "/feat_sub"])
    # fixtures hold a sub-sample; the denominator floor uses the full-map max recorded with it
    dmax = float(golden_model[name + "/depth_stats"][1])
    d_got = sub(depth.cpu())
    den = d_ref.abs().clamp(min=0.1 * dmax)
    rel = ((d_got - d_ref).abs() / den).max().item()
    f_err = (sub(feat.cpu()) - f_ref).abs().max().item()
    report(f"{name}/{precision}", dict(rel_depth=rel, feat_abs=f_err))
    if precision == "fp32":
        assert rel <= 1e-4, rel
        assert f_err <= 1e-3 * max(1.0, float(golden_model[name + "/feat_stats"][1])), f_err
    else:
        assert rel <= 2e-2, rel
        assert f_err <= 6e-2 * max(1.0, float(golden_model[name + "/feat_stats"][1])), f_err


@pytest.mark.parametrize("preset,B,H,W", [("vits", 2, 154, 210), ("vitb", 2, 392, 392)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_matches_oracle_full_maps(preset, B, H, W, precision):
    """Full-map comparison against the oracle (seconds on CPU) with per-stage captures so a
    regression is localised (tokens -> first block -> last block -> decoder)."""
    m, sd, kw = build(preset, 3)
    m.precision = precision
    x = synthetic.make_images(B, H, W, seed=77)
    D = oracle.VIT_CONFIGS[kw["encoder"]]["embed_dim"]
    T = 1 + (H // 14) * (W // 14)
    caps = {k: torch.zeros(B * T * D, device="cuda") for k in ("tokens", "block0", "block_last")}
    with torch.no_grad():
        depth, feat = m._run(x.cuda(), captures=caps)
        d_ref, f_ref, pre = oracle.depth_anything_forward(x, sd, kw["encoder"], return_pre_relu=True)
        _, raw = oracle.vit_intermediate_layers(x, sd, kw["encoder"], return_all=True)
    torch.cuda.synchronize()
    stage = {}
    for k, ref in (("tokens", raw[0]), ("block0", raw[1]), ("block_last", raw[-1])):
        got = caps[k].cpu().reshape(ref.shape)
        stage[k] = ((got - ref).abs().max() / ref.abs().max()).item()
    rel = rel_depth_err(depth.cpu(), d_ref).max().item()
    f_err = ((feat.cpu() - f_ref).abs().max() / f_ref.abs().max()).item()
    sig = (pre - pre.mean()).abs().mean().item()
    sig_err = (depth.cpu() - d_ref).abs().mean().item() / sig
    report(f"{preset}_{H}x{W}/{precision}", dict(rel_depth=rel, feat_rel=f_err, mean_err_over_signal=sig_err, **stage))
    if precision == "fp32":
        assert max(stage.values()) <= 2e-5, stage
        assert rel <= 1e-4, rel
    else:
        assert rel <= 2e-2, rel
        assert sig_err <= 5e-2, sig_err


def test_teacher_equals_student_and_api_contract():
    import distill_any_depth_b200 as d
    kw = synthetic.MODEL_PRESETS["vitl"]
    sd = synthetic.make_state_dict(seed=2, **kw)
    from oracle.make_golden import student_to_teacher_keys
    s = d.DepthAnythingV2(**kw); s.load_state_dict(sd, strict=True); s.cuda()
    t = d.DepthAnything(**kw); t.load_state_dict(student_to_teacher_keys(sd), strict=True); t.cuda()
    x = synthetic.make_images(1, 70, 98, seed=5).cuda()
    ds, fs = s(x)
    dt, ft = t(x)
    assert torch.equal(ds, dt) and torch.equal(fs, ft)  # same kernels, same weights: bit-identical (F6)
    ds2, _ = s(x)
    assert torch.equal(ds, ds2)  # run-to-run determinism
    with pytest.raises(AssertionError):
        s(torch.zeros(1, 3, 75, 70, device="cuda"))
    with pytest.raises(KeyError):
        d.DepthAnythingV2(encoder="vitx")
    with pytest.raises(NotImplementedError):
        d.DepthAnything(use_registers=True)
    # weight updates are picked up (optimizer-style in-place change)
    with torch.no_grad():
        s.depth_head.scratch.output_conv2[2].bias.add_(0.5)
    ds3, _ = s(x)
    assert (ds3 - ds).abs().max().item() > 0.1
