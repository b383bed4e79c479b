"""BASELINE.json configurations at their full sizes.

  * the headline sizes are compared DIRECTLY with the oracle and with the live reference's fixture
    (tests/golden/golden_model_fullsize.npz, generator oracle/make_golden_fullsize.py): ViT-L 518x518 (B=2) and
    ViT-L 1036x1036 (B=1), full depth map and feature map, in the fp32 verification mode (<= 1e-4) and in the
    bf16 tensor-core mode (<= 2e-2 per-pixel relative depth).  The oracle needs ~1 s (518) / ~20 s (1036) per
    image on the GPU box's host cores; the chain is  device == oracle == live reference  at these sizes;
  * images are independent: a batch of 32 must reproduce, bit for bit, what each
    image gives alone - this is what makes sharding by rank exact;
  * the composed distillation step (config 4) must equal its parts.
"""
import pytest
import torch

import os

import numpy as np

import oracle
from distill_any_depth_b200 import synthetic
from helpers import rel_depth_err, sub
from oracle.make_golden_fullsize import FULLSIZE_CASES, DEPTH_SUB, FEAT_SUB

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _inference_mode():
    # forward-only checks: keep the fp32 reference forwards on the inference path (no activation tape)
    with torch.no_grad():
        yield


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


@pytest.fixture(scope="module")
def golden_fullsize():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return dict(np.load(os.path.join(root, "tests", "golden", "golden_model_fullsize.npz")))


_ORACLE_CACHE = {}


def _oracle_forward(name):
    """Oracle forward of a FULLSIZE case on the host cores (cached: the fp32 and bf16 tests share it)."""
    if name not in _ORACLE_CACHE:
        _, preset, B, H, W, ws, xs, hb = next(c for c in FULLSIZE_CASES if c[0] == name)
        kw = synthetic.MODEL_PRESETS[preset]
        sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
        x = synthetic.make_images(B, H, W, seed=xs)
        torch.set_num_threads(os.cpu_count())
        d, f = oracle.depth_anything_forward(x, sd, kw["encoder"])
        _ORACLE_CACHE[name] = (x, d, f)
    return _ORACLE_CACHE[name]


@pytest.mark.parametrize("name", [c[0] for c in FULLSIZE_CASES])
def test_oracle_matches_live_reference_fixture_at_headline_size(name, golden_fullsize):
    """oracle == live reference (depth_anything_v2/dpt.py:211-225) at 518^2 / 1036^2 on THIS box's host BLAS."""
    _, d, f = _oracle_forward(name)
    g = golden_fullsize
    np.testing.assert_allclose(sub(d, DEPTH_SUB).numpy(), g[name + "/depth_sub"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(sub(f, FEAT_SUB).numpy(), g[name + "/feat_sub"], rtol=1e-4, atol=5e-5)
    st = g[name + "/depth_stats"]
    assert abs(d.double().pow(2).sum().item() - st[2]) <= 1e-4 * st[2]


@pytest.mark.parametrize("precision,tol,ftol", [("fp32", 1e-4, 1e-4), ("bf16", 2e-2, 6e-2)])
@pytest.mark.parametrize("name", [c[0] for c in FULLSIZE_CASES])
def test_headline_sizes_match_oracle_full_map(name, precision, tol, ftol, golden_fullsize):
    """ViT-L 518x518 (B=2) and 1036x1036 (B=1, 5477 tokens): every pixel of the depth map and every feature against the
    oracle, and the same sub-sample against the live reference's fixture (dinov2_layers/attention.py:49-62 at N=5477)."""
    _, preset, B, H, W, ws, xs, hb = next(c for c in FULLSIZE_CASES if c[0] == name)
    m, _, _ = build(preset, ws, head_bias=hb)
    x, d_or, f_or = _oracle_forward(name)
    m.precision = precision
    depth, feat = m(x.cuda())
    torch.cuda.synchronize()
    assert depth.shape == d_or.shape and feat.shape == f_or.shape
    rel = rel_depth_err(depth.cpu(), d_or).max().item()
    f_err = ((feat.cpu() - f_or).abs().max() / f_or.abs().max()).item()
    d_fix = torch.from_numpy(golden_fullsize[name + "/depth_sub"])
    dmax = float(golden_fullsize[name + "/depth_stats"][1])
    rel_fix = ((sub(depth.cpu(), DEPTH_SUB) - d_fix).abs() / d_fix.abs().clamp(min=0.1 * dmax)).max().item()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_fullsize.jsonl"), "a") as fh:
        fh.write('{"case": "%s", "precision": "%s", "rel_depth_vs_oracle": %.3e, "rel_depth_vs_reference_fixture": %.3e, '
                 '"feat_rel_vs_oracle": %.3e}\n' % (name, precision, rel, rel_fix, f_err))
    assert rel <= tol, rel
    assert rel_fix <= tol, rel_fix
    assert f_err <= ftol, f_err


def test_config3_vitl_518_batch32_batch_independence_and_bf16_tolerance():
    """BASELINE configs[2]: ViT-L 518x518 batch 32 bf16 (the bench workload)."""
    m, _, _ = build("vitl", 1)
    x = synthetic.make_images(32, 518, 518, seed=1234).cuda()
    m.precision = "bf16"
    d32, f32 = m(x)
    assert d32.shape == (32, 1, 518, 518) and f32.shape == (32, 37 * 37, 1024)
    assert torch.isfinite(d32).all() and torch.isfinite(f32).all()
    assert float(d32.max()) > 0  # live output
    for i in (0, 17, 31):
        di, fi = m(x[i:i + 1].contiguous())
        assert torch.equal(di[0], d32[i]), f"image {i}: batch-32 output differs from the single-image output"
        assert torch.equal(fi[0], f32[i])
    m.precision = "fp32"
    dref, fref = m(x[:2].contiguous())
    rel = rel_depth_err(d32[:2], dref).max().item()
    assert rel <= 2e-2, rel
    assert ((f32[:2] - fref).abs().max() / fref.abs().max()).item() <= 6e-2


def test_config5_vitl_1036_highres():
    """BASELINE configs[4]: ViT-L 1036x1036 (5476 patch tokens), attention-dominated."""
    m, _, _ = build("vitl", 1)
    x = synthetic.make_images(2, 1036, 1036, seed=99).cuda()
    m.precision = "bf16"
    d2, f2 = m(x)
    assert d2.shape == (2, 1, 1036, 1036) and f2.shape == (2, 74 * 74, 1024)
    d1, _ = m(x[1:2].contiguous())
    assert torch.equal(d1[0], d2[1])
    m.precision = "fp32"
    dref, _ = m(x[:1].contiguous())
    rel = rel_depth_err(d2[:1], dref).max().item()
    assert rel <= 2e-2, rel


def test_config2_vitb_392_batch16_ssi_grad_losses():
    """BASELINE configs[1]: student ViT-B 392x392 batch 16 forward + SSI / gradient loss; losses checked against
    the oracle evaluated on the device's own depth maps (loss tolerance 1e-3 relative)."""
    import distill_any_depth_b200 as d
    m, _, _ = build("vitb", 3)
    x = synthetic.make_images(16, 392, 392, seed=5).cuda()
    m.precision = "bf16"
    depth, _ = m(x)
    _, gt, mask = synthetic.make_depth_pair(16, 392, 392, seed=6)
    ssi = d.SSILoss()(depth, gt.cuda(), mask.cuda())
    grad = d.gradient_preservation_loss(depth)
    dc = depth.cpu()
    ssi_ref = oracle.SSILoss()(dc, gt, mask)
    grad_ref = oracle.gradient_preservation_loss(dc)
    assert abs(float(ssi) - float(ssi_ref)) <= 1e-3 * abs(float(ssi_ref))
    assert abs(float(grad) - float(grad_ref)) <= 1e-3 * abs(float(grad_ref))
    m.precision = "fp32"
    dref, _ = m(x[:2].contiguous())
    assert rel_depth_err(depth[:2], dref).max().item() <= 2e-2


def test_config4_distillation_step_equals_its_parts():
    """BASELINE configs[3] per-GPU shard: ViT-L teacher + ViT-B student at 392x392, 16 images, SC / LG / feature /
    gradient / HDN-DR losses (tools/train_distillation.py:1509-1560).  The composed step must equal the public
    loss functions applied to the same forwards, and those equal the oracle on the same maps."""
    import distill_any_depth_b200 as d
    student, _, _ = build("vitb", 3)
    teacher, _, _ = build("vitl", 2, teacher=True, head_bias=0.6)
    x = synthetic.make_images(16, 392, 392, seed=11).cuda()
    out = d.distillation_step_losses(student, teacher, x, x)
    sd, sf = student(x)
    td, tf = teacher(x)
    parts = dict(sc_loss=d.distillation_loss(sd, td, "hybrid"), lg_loss=d.distillation_loss(sd, sd, "hybrid"),
                 feat_loss=d.feature_distillation_loss(sf, tf), grad_loss=d.gradient_preservation_loss(sd),
                 hdn_loss=d.compute_hdn_loss(d.SSILoss(), sd, td, d.get_contexts_dr(3, td, None)))
    for k, v in parts.items():
        assert abs(float(out[k]) - float(v)) <= 1e-6 * max(abs(float(v)), 1e-6), (k, float(out[k]), float(v))
    total = 0.5 * parts["sc_loss"] + 0.5 * parts["lg_loss"] + 1.0 * parts["feat_loss"] + 0.2 * parts["grad_loss"] \
        + 0.8 * parts["hdn_loss"]
    assert abs(float(out["batch_loss"]) - float(total)) <= 1e-5 * abs(float(total))
    assert float(out["lg_loss"]) == 0.0  # identical global / local inputs: identity distillation (:1524-1529)
    dedup = d.distillation_step_losses(student, teacher, x, x, dedup_student=True)
    assert float(dedup["batch_loss"]) == float(out["batch_loss"])
    # oracle on the device's maps (4 images keep the CPU side to seconds)
    sc, tc = sd[:4].cpu(), td[:4].cpu()
    ref = dict(sc=oracle.distillation_loss(sc, tc, "hybrid"), grad=oracle.gradient_preservation_loss(sc),
               hdn=oracle.compute_hdn_loss(oracle.SSILoss(), sc, tc, oracle.get_contexts_dr(3, tc, None)),
               feat=oracle.feature_distillation_loss(sf[:4].cpu(), tf[:4].cpu()))
    got = dict(sc=d.distillation_loss(sd[:4].contiguous(), td[:4].contiguous(), "hybrid"),
               grad=d.gradient_preservation_loss(sd[:4].contiguous()),
               hdn=d.hdn_loss_dr(sd[:4].contiguous(), td[:4].contiguous(), None, 3),
               feat=d.feature_distillation_loss(sf[:4].contiguous(), tf[:4].contiguous()))
    for k in ref:
        assert abs(float(got[k]) - float(ref[k])) <= 1e-3 * max(abs(float(ref[k])), 1e-6), (k, float(got[k]), float(ref[k]))


def test_cuda_graph_capture_replays_bit_identically():
    """The whole step (forward + SSI + fused HDN-DR + gradient loss) is capturable: nothing allocates or synchronises
    inside the library.  Replay must reproduce the eager forward bit for bit (losses to 1e-6), also on new input data."""
    import distill_any_depth_b200 as d
    m, _, _ = build("vitb", 3)
    x0 = synthetic.make_images(4, 392, 392, seed=5).cuda()
    x1 = synthetic.make_images(4, 392, 392, seed=6).cuda()
    _, gt, mask = synthetic.make_depth_pair(4, 392, 392, seed=7)
    gt, mask = gt.cuda(), mask.cuda()

    def step(x):
        depth, feat = m(x)
        return depth, feat, d.SSILoss()(depth, gt, mask), d.hdn_loss_dr(depth, gt, None, 3), d.gradient_preservation_loss(depth)

    eager0 = [t.clone() for t in step(x0)]
    eager1 = [t.clone() for t in step(x1)]
    cap = d.capture(step, x0)
    for x, ref in ((x0, eager0), (x1, eager1), (x0, eager0)):
        out = cap(x)
        torch.cuda.synchronize()
        assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])   # forward: bit-identical
        for a, b in zip(out[2:], ref[2:]):   # losses: fp64 atomics accumulate in arrival order
            assert abs(float(a) - float(b)) <= 1e-6 * abs(float(b))
    with pytest.raises(ValueError):
        cap(x0[:2])


def test_cuda_graph_replay_sees_weight_updates_including_the_head_bias():
    """graph.py contract: after a weight update, one eager forward repacks the operands in place and the captured graph
    then computes with the NEW weights.  output_conv2.2.bias used to travel by value inside the captured kernel
    parameters (stale after an optimiser step); it is a device pointer now."""
    import distill_any_depth_b200 as d
    for precision in ("bf16", "fp32"):
        m, _, _ = build("vits", 0)
        m.precision = precision
        x = synthetic.make_images(2, 70, 98, seed=3).cuda()
        cap = d.capture(lambda t: m(t), x)
        before = [t.clone() for t in cap(x)]
        with torch.no_grad():
            m.depth_head.scratch.output_conv2[2].bias.add_(0.75)
            m.depth_head.scratch.output_conv2[0].weight.mul_(1.5)
            m.pretrained.blocks[0].mlp.fc1.bias.add_(0.05)
        eager = [t.clone() for t in m(x)]          # repacks in place
        after = cap(x)
        torch.cuda.synchronize()
        assert not torch.equal(before[0], eager[0])
        assert torch.equal(after[0], eager[0]) and torch.equal(after[1], eager[1]), precision
        assert float((eager[0] - before[0]).abs().min()) > 0.1   # the bias shift reached every pixel


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_losses_run_on_the_tensors_device_not_the_current_one():
    """Tensors on cuda:1 while cuda:0 is current (the reference's .to(device) usage): every loss entry point must launch
    on the tensors' device and stream."""
    import distill_any_depth_b200 as d
    pred, gt, mask = synthetic.make_depth_pair(3, 56, 84, seed=8)
    feats = synthetic.make_features(3, 49, 96, seed=1)
    feats_t = synthetic.make_features(3, 49, 128, seed=2)
    ref = dict(ssi=oracle.SSILoss()(pred, gt, mask), grad=oracle.gradient_preservation_loss(pred),
               hdn=oracle.compute_hdn_loss(oracle.SSILoss(), pred, gt, oracle.get_contexts_dr(3, gt, mask)),
               dist=oracle.distillation_loss(pred, gt, "hybrid"), feat=oracle.feature_distillation_loss(feats, feats_t))
    dev = torch.device("cuda:1")
    torch.cuda.set_device(0)
    p, g, mk = pred.to(dev), gt.to(dev), mask.to(dev)
    got = dict(ssi=d.SSILoss()(p, g, mk), grad=d.gradient_preservation_loss(p),
               hdn=d.compute_hdn_loss(d.SSILoss(), p, g, d.get_contexts_dr(3, g, mk)),
               dist=d.distillation_loss(p, g, "hybrid"), feat=d.feature_distillation_loss(feats.to(dev), feats_t.to(dev)))
    assert torch.cuda.current_device() == 0
    for k in ref:
        assert got[k].device == dev
        assert abs(float(got[k]) - float(ref[k])) <= 1e-3 * max(abs(float(ref[k])), 1e-6), k
