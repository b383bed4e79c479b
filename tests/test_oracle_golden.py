"""CPU: the oracle restatement reproduces the LIVE reference's outputs recorded in
tests/golden/ by oracle/make_golden.py (the oracle's pin)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.make_golden import MODEL_CASES, LOSS_CASES
from oracle.make_golden_options import OPTION_CASES
from distill_any_depth_b200 import synthetic
from helpers import sub


@pytest.mark.parametrize("case", [c for c in MODEL_CASES if c[0] != "vits_518"], ids=lambda c: c[0])
def test_model_matches_reference_fixture(case, golden_model):
    name, preset, B, H, W, ws, xs, _teacher, hb = case
    kw = synthetic.MODEL_PRESETS[preset]
    sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
    x = synthetic.make_images(B, H, W, seed=xs)
    with torch.no_grad():
        d, f = oracle.depth_anything_forward(x, sd, kw["encoder"])
    np.testing.assert_allclose(sub(d).numpy(), golden_model[name + "/depth_sub"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(sub(f).numpy(), golden_model[name + "/feat_sub"], rtol=1e-4, atol=2e-5)
    st = golden_model[name + "/depth_stats"]
    assert abs(d.mean().item() - st[0]) <= 1e-5 * max(1, abs(st[0]))
    assert abs(d.double().pow(2).sum().item() - st[2]) <= 1e-4 * st[2]


def test_vitl_518_matches_reference_fixture_at_headline_size():
    """The bench workload's size (ViT-L 518x518): oracle vs the live reference's recorded outputs
    (oracle/make_golden_fullsize.py; the 1036x1036 case is replayed by the GPU suite, where the oracle runs anyway)."""
    from oracle.make_golden_fullsize import FULLSIZE_CASES, DEPTH_SUB, FEAT_SUB
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_model_fullsize.npz")))
    name, preset, B, H, W, ws, xs, hb = FULLSIZE_CASES[0]
    kw = synthetic.MODEL_PRESETS[preset]
    sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
    x = synthetic.make_images(B, H, W, seed=xs)[:1]
    with torch.no_grad():
        d, f = oracle.depth_anything_forward(x, sd, kw["encoder"])
    np.testing.assert_allclose(sub(d, DEPTH_SUB).numpy(), g[name + "/depth_sub"][:1], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(sub(f, FEAT_SUB).numpy(), g[name + "/feat_sub"][:1], rtol=1e-4, atol=5e-5)


@pytest.mark.parametrize("case", OPTION_CASES, ids=lambda c: c[0])
def test_option_branches_match_reference_fixture(case, golden_model_options):
    """use_clstoken readout and ViT-g / SwiGLU against the live reference's recorded outputs."""
    name, kw, B, H, W, ws, xs, hb = case
    sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
    x = synthetic.make_images(B, H, W, seed=xs)
    with torch.no_grad():
        d, f = oracle.depth_anything_forward(x, sd, kw["encoder"])
    g = golden_model_options
    np.testing.assert_allclose(sub(d).numpy(), g[name + "/depth_sub"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(sub(f).numpy(), g[name + "/feat_sub"], rtol=1e-4, atol=2e-5)
    st = g[name + "/depth_stats"]
    assert abs(d.double().pow(2).sum().item() - st[2]) <= 1e-4 * st[2]


@pytest.mark.parametrize("case", LOSS_CASES, ids=lambda c: c[0])
def test_losses_match_reference_fixture(case, golden_losses):
    name, B, H, W, seed = case
    pred, gt, mask = synthetic.make_depth_pair(B, H, W, seed=seed)
    full = torch.ones_like(mask)
    fs = synthetic.make_features(B, 49, 96, seed=seed + 100)
    ft = synthetic.make_features(B, 49, 128, seed=seed + 200)
    pred[0, 0, 0, :8] = pred[0, 0, 1, :8]
    gt[0, 0, 2, 3] = gt[0].max()
    ssi = oracle.SSILoss()
    got = {}
    for tag, mk in (("mask", mask), ("full", full)):
        pa, ga = oracle.masked_shift_and_scale(pred, gt, mk)
        got[f"align_pred_{tag}"], got[f"align_gt_{tag}"] = sub(pa), sub(ga)
        got[f"ssi_{tag}"] = ssi(pred, gt, mk)
        got[f"ssi_dense_{tag}"] = sub(ssi(pred, gt, mk, dense=True))
        ctx = oracle.get_contexts_dr(3, gt, mk)
        got[f"ctx_dr_count_{tag}"] = sub(ctx.sum(0).float())
        got[f"hdn_dr_{tag}"] = oracle.compute_hdn_loss(ssi, pred, gt, ctx)
        got[f"hdn_dp_{tag}"] = oracle.compute_hdn_loss(ssi, pred, gt, oracle.get_contexts_dp(3, gt, mk))
        if H == W:
            got[f"hdn_ds_{tag}"] = oracle.compute_hdn_loss(ssi, pred, gt, oracle.get_contexts_ds(3, mk))
    got["ctx_dr_none"] = sub(oracle.get_contexts_dr(3, gt, None).sum(0).float())
    got["grad"] = oracle.gradient_preservation_loss(pred)
    got["feat"] = oracle.feature_distillation_loss(fs, ft)
    got["feat_same"] = oracle.feature_distillation_loss(fs, fs * 0.5 + 0.1)
    for st in ("global", "hybrid", "local", "none"):
        got[f"distill_{st}"] = oracle.distillation_loss(pred, gt, st)
    got["norm_hybrid"] = sub(oracle.hybrid_normalize(pred, 4))
    got["norm_global"] = sub(oracle.global_normalize(pred))
    zero, half, empty = torch.zeros_like(gt), torch.full_like(gt, 0.5), torch.zeros_like(mask)
    for tag, gg, mk in (("allzero", zero, full), ("const", half, full), ("empty", gt, empty)):
        got[f"hdn_dr_{tag}"] = oracle.compute_hdn_loss(ssi, pred, gg, oracle.get_contexts_dr(3, gg, mk))
        got[f"ssi_{tag}"] = ssi(pred, gg, mk)
    n = 0
    for k, v in got.items():
        exp = golden_losses[f"{name}/{k}"]
        v = torch.as_tensor(v).detach().float().numpy()
        scale = max(float(np.abs(exp).max()), 1e-6)
        assert np.abs(v - exp).max() <= 2e-5 * scale, (k, float(np.abs(v - exp).max()), scale)
        n += 1
    assert n >= 30


def test_lower_median_and_edge_semantics():
    """Known answers for the selection rules the CUDA path must reproduce."""
    x = torch.tensor([[1., 2., 3., 4.]]).view(1, 1, 1, 4)
    m = torch.ones_like(x, dtype=torch.bool)
    p, g = oracle.masked_shift_and_scale(x, x, m)
    # lower median of [1,2,3,4] is 2; s = (1+0+1+2)/(4+1) = 0.8
    np.testing.assert_allclose(p.flatten().numpy(), (np.array([1, 2, 3, 4.]) - 2) / (0.8 + 1e-6), rtol=1e-6)
    # all-masked row -> t = 0, s = 0 -> x / 1e-6
    p, _ = oracle.masked_shift_and_scale(x, x, torch.zeros_like(m))
    np.testing.assert_allclose(p.flatten().numpy(), np.array([1, 2, 3, 4.]) / 1e-6, rtol=1e-6)
    # HDN-DR: max pixel belongs to no context, every other pixel to exactly 3
    d = torch.linspace(0.3, 0.9, 64).view(1, 1, 8, 8)
    c = oracle.get_contexts_dr(3, d, None).sum(0)
    assert c.flatten()[-1] == 0 and bool((c.flatten()[:-1] == 3).all())
    # constant 0 -> all 7 contexts; constant 0.5 -> none (SURVEY.md A.4 iii)
    assert bool((oracle.get_contexts_dr(3, torch.zeros(1, 1, 4, 4), None).sum(0) == 7).all())
    assert bool((oracle.get_contexts_dr(3, torch.full((1, 1, 4, 4), 0.5), None).sum(0) == 0).all())


def test_preprocess_oracle_matches_reference_fixture():
    """oracle/preprocess.py (cv2 / numpy restatement of dpt.py:237-262) against the committed outputs of the live
    reference class on seeded synthetic images (oracle/make_golden_preprocess.py)."""
    import numpy as np
    from oracle import preprocess as P
    from oracle.make_golden_preprocess import CASES, synthetic_image, sub as psub
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_preprocess.npz")))
    for name, h, w, size, keep, seed in CASES:
        t, hw = P.image2tensor(synthetic_image(h, w, seed), size, keep_aspect_ratio=keep)
        assert tuple(t.shape) == tuple(g[name + "/shape"]) and hw == (h, w)
        assert np.array_equal(psub(t.numpy()), g[name + "/tensor_sub"]), name
    assert P.get_size(640, 480, 518, 518) == (686, 518)
    assert P.get_size(1242, 375, 518, 518) == (1722, 518)


def test_resize_size_arithmetic_matches_the_reference_sweep():
    """Resize.get_size (util/transform.py:52-106) over 4050 (raw size, target, keep_aspect_ratio, method) cases recorded
    from the live reference (oracle/make_golden_sizes.py): the product's host logic for all three methods, the oracle's
    restatement for 'lower_bound' (the method on the path, dpt.py:240-250)."""
    from distill_any_depth_b200 import preprocess
    from oracle.make_golden_sizes import METHODS
    from oracle.preprocess import get_size as oracle_get_size
    rows = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_sizes.npz"))["rows"]
    assert len(rows) == 4050
    for w, h, target, keep, mi, nw, nh in rows.tolist():
        got = preprocess.get_size(w, h, target, target, keep_aspect_ratio=bool(keep), resize_method=METHODS[mi])
        assert got == (nw, nh), (w, h, target, keep, METHODS[mi], got, (nw, nh))
        if mi == 0:
            assert oracle_get_size(w, h, target, target, bool(keep)) == (nw, nh), (w, h, target, keep)
