"""GPU parity of the loss kernels (through the reference-named Python mirror and the C ABI)
against the CPU oracle on the same seeded inputs, plus the golden reference fixtures, edge
cases and size-independent properties at BASELINE sizes."""
import numpy as np
import pytest
import torch

import oracle
from distill_any_depth_b200 import synthetic
from helpers import rel_scalar, sub
from oracle.make_golden import LOSS_CASES

pytestmark = pytest.mark.gpu
TOL = 1e-3  # north_star: loss values within 1e-3 relative (observed ~1e-6)


def dad():
    import distill_any_depth_b200 as d
    return d


def _case(B, H, W, seed):
    pred, gt, mask = synthetic.make_depth_pair(B, H, W, seed=seed)
    pred[0, 0, 0, :8] = pred[0, 0, 1, :8]
    gt[0, 0, 2, 3] = gt[0].max()
    return pred, gt, mask


@pytest.mark.parametrize("case", LOSS_CASES, ids=lambda c: c[0])
def test_losses_match_golden_reference(case, golden_losses):
    d = dad()
    name, B, H, W, seed = case
    pred, gt, mask = _case(B, H, W, seed)
    full = torch.ones_like(mask)
    fs = synthetic.make_features(B, 49, 96, seed=seed + 100)
    ft = synthetic.make_features(B, 49, 128, seed=seed + 200)
    P, G = pred.cuda(), gt.cuda()
    ssi = d.SSILoss()
    got = {}
    for tag, mk in (("mask", mask), ("full", full)):
        M = mk.cuda()
        pa, ga = d.masked_shift_and_scale(P, G, M)
        got[f"align_pred_{tag}"], got[f"align_gt_{tag}"] = sub(pa.cpu()), sub(ga.cpu())
        got[f"ssi_{tag}"] = ssi(P, G, M)
        got[f"ssi_dense_{tag}"] = sub(ssi(P, G, M, dense=True).cpu())
        ctx = d.get_contexts_dr(3, G, M)
        got[f"ctx_dr_count_{tag}"] = sub(ctx.sum(0).float().cpu())
        got[f"hdn_dr_{tag}"] = d.compute_hdn_loss(ssi, P, G, ctx)                 # fused path
        got[f"hdn_dr_{tag}#generic"] = d.compute_hdn_loss(ssi, P, G, ctx.clone())  # explicit-context path
        cdp = d.get_contexts_dp(3, G, M)
        assert torch.equal(cdp.cpu(), oracle.get_contexts_dp(3, gt, mk)), "get_contexts_dp differs from the oracle"
        got[f"hdn_dp_{tag}"] = d.compute_hdn_loss(ssi, P, G, cdp)
        if H == W:
            cds = d.get_contexts_ds(3, M)
            assert torch.equal(cds.cpu(), oracle.get_contexts_ds(3, mk)), "get_contexts_ds differs from the oracle"
            got[f"hdn_ds_{tag}"] = d.compute_hdn_loss(ssi, P, G, cds)
    got["ctx_dr_none"] = sub(d.get_contexts_dr(3, G, None).sum(0).float().cpu())
    got["grad"] = d.gradient_preservation_loss(P)
    got["feat"] = d.feature_distillation_loss(fs.cuda(), ft.cuda())
    got["feat_same"] = d.feature_distillation_loss(fs.cuda(), (fs * 0.5 + 0.1).cuda())
    for st in ("global", "hybrid", "local", "none"):
        got[f"distill_{st}"] = d.distillation_loss(P, G, st)
    got["norm_hybrid"] = sub(d.hybrid_normalize(P, 4).cpu())
    got["norm_global"] = sub(d.global_normalize(P).cpu())
    zero, half, empty = torch.zeros_like(gt), torch.full_like(gt, 0.5), torch.zeros_like(mask)
    for tag, gg, mk in (("allzero", zero, full), ("const", half, full), ("empty", gt, empty)):
        got[f"hdn_dr_{tag}"] = d.compute_hdn_loss(ssi, P, gg.cuda(), d.get_contexts_dr(3, gg.cuda(), mk.cuda()))
        got[f"ssi_{tag}"] = ssi(P, gg.cuda(), mk.cuda())
    bad = []
    for k, v in got.items():
        exp = golden_losses[f"{name}/{k.split('#')[0]}"]
        v = torch.as_tensor(v).detach().float().cpu().numpy()
        scale = max(float(np.abs(exp).max()), 1e-6)
        err = float(np.abs(v - exp).max()) / scale
        tol = 2e-5 if v.ndim else TOL
        if k.startswith("ctx_"):
            tol = 0.0
        if not err <= tol:
            bad.append((k, err))
    assert not bad, bad


def test_contexts_dr_bit_exact_vs_oracle():
    d = dad()
    for B, H, W, seed in ((4, 98, 126, 3), (2, 392, 392, 4)):
        pred, gt, mask = synthetic.make_depth_pair(B, H, W, seed=seed)
        for mk in (None, mask):
            ref = oracle.get_contexts_dr(3, gt, mk)
            got = d.get_contexts_dr(3, gt.cuda(), None if mk is None else mk.cuda()).cpu()
            assert torch.equal(ref, got)


@pytest.mark.parametrize("B,H,W", [(16, 392, 392), (4, 518, 518)])
def test_losses_match_oracle_at_baseline_sizes(B, H, W):
    d = dad()
    pred, gt, mask = synthetic.make_depth_pair(B, H, W, seed=21)
    P, G, M = pred.cuda(), gt.cuda(), mask.cuda()
    ssi_o, ssi_g = oracle.SSILoss(), d.SSILoss()
    pairs = {
        "ssi_full": (ssi_g(P, G, torch.ones_like(M)), ssi_o(pred, gt, torch.ones_like(mask))),
        "ssi_mask": (ssi_g(P, G, M), ssi_o(pred, gt, mask)),
        "hdn_dr": (d.hdn_loss_dr(P, G, None, 3), oracle.compute_hdn_loss(ssi_o, pred, gt, oracle.get_contexts_dr(3, gt, None))),
        "hdn_dr_mask": (d.hdn_loss_dr(P, G, M, 3), oracle.compute_hdn_loss(ssi_o, pred, gt, oracle.get_contexts_dr(3, gt, mask))),
        "grad": (d.gradient_preservation_loss(P), oracle.gradient_preservation_loss(pred)),
        "hybrid": (d.distillation_loss(P, G, "hybrid"), oracle.distillation_loss(pred, gt, "hybrid")),
        "global": (d.distillation_loss(P, G, "global"), oracle.distillation_loss(pred, gt, "global")),
        "none": (d.distillation_loss(P, G, "none"), oracle.distillation_loss(pred, gt, "none")),
    }
    bad = {k: (float(a), float(b)) for k, (a, b) in pairs.items() if rel_scalar(a, b) > TOL}
    assert not bad, bad
    # aligned maps are exact up to the rounding of the MAD sum
    pa, ga = d.masked_shift_and_scale(P, G, M)
    po, go = oracle.masked_shift_and_scale(pred, gt, mask)
    assert (pa.cpu() - po).abs().max().item() <= 2e-5 * po.abs().max().item()
    assert (ga.cpu() - go).abs().max().item() <= 2e-5 * go.abs().max().item()


def test_feature_loss_at_config2_shape():
    d = dad()
    fs = synthetic.make_features(16, 784, 768, seed=5)
    ft = synthetic.make_features(16, 784, 1024, seed=6)
    got = d.feature_distillation_loss(fs.cuda(), ft.cuda())
    assert rel_scalar(got, oracle.feature_distillation_loss(fs, ft)) <= TOL
    got = d.feature_distillation_loss([fs.cuda(), None, fs.cuda()], [ft.cuda(), ft.cuda(), (fs * 2).cuda()])
    ref = oracle.feature_distillation_loss([fs, None, fs], [ft, ft, fs * 2])
    assert rel_scalar(got, ref) <= TOL


def test_median_properties_and_edge_cases():
    """Size-independent properties: the aligned map has (lower) median 0 and mean |.| over the mask
    equal to n/(n+1) * s/(s+1e-6); shift / positive-scale invariance of SSI; known answers."""
    d = dad()
    pred, gt, mask = synthetic.make_depth_pair(3, 224, 224, seed=9)
    P, G, M = pred.cuda(), gt.cuda(), mask.cuda()
    pa, _ = d.masked_shift_and_scale(P, G, M)
    for b in range(3):
        v = pa[b][M[b]].cpu()
        srt = v.sort().values
        assert srt[(len(srt) - 1) // 2].item() == 0.0
        n = len(v)
        assert abs(v.abs().mean().item() - n / (n + 1)) < 1e-3
    ssi = d.SSILoss()
    a = ssi(P, G, M)
    b = ssi(P * 3.0 + 2.0, G * 0.5 - 1.0, M)
    assert rel_scalar(a, b) < 1e-3
    # known answer: lower median of [1,2,3,4] is 2, s = 4/5
    x = torch.tensor([1., 2., 3., 4.], device="cuda").view(1, 1, 1, 4)
    m = torch.ones_like(x, dtype=torch.bool)
    p, _ = d.masked_shift_and_scale(x, x, m)
    np.testing.assert_allclose(p.flatten().cpu().numpy(), (np.array([1, 2, 3, 4.]) - 2) / (0.8 + 1e-6), rtol=1e-6)
    p, _ = d.masked_shift_and_scale(x, x, torch.zeros_like(m))
    np.testing.assert_allclose(p.flatten().cpu().numpy(), np.array([1, 2, 3, 4.]) / 1e-6, rtol=1e-6)
    # negative values and ties
    y = torch.tensor([-3., -1., -1., 5., 0., -1., 2.], device="cuda").view(1, 1, 1, 7)
    p, _ = d.masked_shift_and_scale(y, y, torch.ones_like(y, dtype=torch.bool))
    ref, _ = oracle.masked_shift_and_scale(y.cpu(), y.cpu(), torch.ones(1, 1, 1, 7, dtype=torch.bool))
    np.testing.assert_allclose(p.cpu().numpy(), ref.numpy(), rtol=1e-6)


def test_unsupported_options_raise():
    d = dad()
    with pytest.raises(TypeError):
        d.get_contexts_dp(3, torch.zeros(1, 1, 4, 4, device="cuda"), None)
    with pytest.raises(RuntimeError):  # non-square map: the reference's template broadcast fails too
        d.get_contexts_ds(3, torch.ones(1, 1, 4, 6, dtype=torch.bool, device="cuda"))
    with pytest.raises(NotImplementedError):
        d.feature_distillation_loss(torch.zeros(1, 4, 8, device="cuda"), torch.zeros(1, 5, 8, device="cuda"))
    with pytest.raises(ValueError):
        d.distillation_loss(torch.zeros(1, 1, 4, 4, device="cuda"), torch.zeros(1, 1, 4, 4, device="cuda"), "bogus")
    with pytest.raises(RuntimeError):
        d.gradient_preservation_loss(torch.zeros(1, 1, 4, 4))


@pytest.mark.parametrize("level", [1, 2, 3, 4])
@pytest.mark.parametrize("B,H,W", [(3, 37, 53), (2, 64, 64)])
def test_contexts_dp_ds_bit_exact(level, B, H, W):
    """HDN-DP / HDN-DS context builders (tools/train_distillation.py:578-673): bit-exact against the oracle,
    including ties at quantile boundaries, images with 0 / 1 / 2 valid pixels and 1-ulp-adjacent values."""
    d = dad()
    g = torch.Generator().manual_seed(100 * level + B)
    gt = torch.rand(B, 1, H, W, generator=g)
    gt[0, 0, :4] = gt[0, 0, 4:8]                       # ties
    gt[0, 0, 9, :16] = torch.nextafter(gt[0, 0, 8, :16], torch.tensor(2.0))  # adjacent floats
    mask = torch.rand(B, 1, H, W, generator=g) > 0.3
    if B > 2:
        mask[2] = False
        mask[2, 0, 3, 5] = True                         # single valid pixel
    mask[1, 0, ::2] = False
    for mk in (mask, torch.ones_like(mask), torch.zeros_like(mask)):
        got = d.get_contexts_dp(level, gt.cuda(), mk.cuda())
        exp = oracle.get_contexts_dp(level, gt, mk)
        assert got.shape == exp.shape and got.dtype == torch.bool
        assert torch.equal(got.cpu(), exp), (level, int((got.cpu() != exp).sum()))
        if H == W and level <= 3:
            gs = d.get_contexts_ds(level, mk.cuda())
            assert torch.equal(gs.cpu(), oracle.get_contexts_ds(level, mk))


def _fused_vs_parts(P, G, M, level, pred, gt, mask, tol=2e-5):
    """ssi_hdn_dr against (a) the two separate device paths and (b) the oracle."""
    d = dad()
    ssi_f, hdn_f, ps, ph = d.ssi_hdn_dr(P, G, M, level, want_partials=True)
    full = torch.ones_like(G, dtype=torch.bool) if M is None else M
    ssi_s = d.SSILoss()(P, G, full)
    hdn_s, ph_s = d.hdn_loss_dr(P, G, M, level, want_partials=True)
    assert rel_scalar(ssi_f, ssi_s) <= tol, ("ssi", float(ssi_f), float(ssi_s))
    assert rel_scalar(hdn_f, hdn_s) <= tol, ("hdn", float(hdn_f), float(hdn_s))
    assert float(ph[1]) == float(ph_s[1]), "HDN valid-location counts differ"   # integer work: exact
    mk = torch.ones_like(gt, dtype=torch.bool) if mask is None else mask
    ssi_o = oracle.SSILoss()(pred, gt, mk)
    hdn_o = oracle.compute_hdn_loss(oracle.SSILoss(), pred, gt, oracle.get_contexts_dr(level, gt, mask))
    assert rel_scalar(ssi_f, ssi_o) <= TOL and rel_scalar(hdn_f, hdn_o) <= TOL
    assert float(ps[1]) == float(mk.sum())
    return float(ssi_f), float(hdn_f)


@pytest.mark.parametrize("level", [1, 2, 3])
@pytest.mark.parametrize("B,H,W,seed,masked", [(2, 64, 64, 7, True), (3, 56, 84, 8, False), (1, 14, 14, 3, True),
                                              (2, 129, 71, 5, True)])
def test_fused_ssi_hdn_matches_the_separate_paths_and_the_oracle(level, B, H, W, seed, masked):
    """One shared sweep for SSILoss + HDN-DR (losses_fused.cu) == the two separate kernels' values == the oracle
    (tools/train_distillation.py:449-576, 686-707); ties, the maximum pixel (outside every half-open bin) and
    ragged sizes included."""
    pred, gt, mask = _case(B, H, W, seed)
    M = mask.cuda() if masked else None
    _fused_vs_parts(pred.cuda(), gt.cuda(), M, level, pred, gt, mask if masked else None)


def test_fused_ssi_hdn_degenerate_images():
    """Images whose depth-range thresholds do not nest exactly (constant gt, all-zero gt, no valid pixel, heavy ties,
    a huge dynamic range) take the row-per-unit path inside the same kernels; mixed with ordinary images in one batch."""
    d = dad()
    pred, gt, mask = _case(6, 40, 52, 21)
    gt[1] = 0.5                                  # constant: every context collapses to [0.5, 0.5 + 1e-30)
    gt[2] = 0.0                                  # all-zero: 1e-30 is NOT absorbed, every pixel is in every context
    mask[3] = False                              # no valid pixel in image 3
    gt[4] = torch.round(gt[4] * 3) / 3           # four distinct values: medians sit in heavy ties (list overflow path)
    pred[4] = torch.round(pred[4] * 2) / 2
    gt[5, 0, 0, 0] = 1e30                        # one outlier: almost everything in the first bin
    for M, mk in ((mask.cuda(), mask), (None, None)):
        _fused_vs_parts(pred.cuda(), gt.cuda(), M, 3, pred, gt, mk)
    # a batch of only-degenerate images, and level 2
    _fused_vs_parts(pred[1:3].cuda().contiguous(), gt[1:3].cuda().contiguous(), None, 2, pred[1:3], gt[1:3], None)


def test_fused_ssi_hdn_at_baseline_size_and_gradient_fallback():
    """BASELINE configs[2] shape (the benchmark's loss half): 4 x 518 x 518, full mask, level 3; with requires_grad the
    autograd-aware separate losses are evaluated (same values)."""
    d = dad()
    pred, gt, _ = synthetic.make_depth_pair(4, 518, 518, seed=7)
    P, G = pred.cuda(), gt.cuda()
    ssi_f, hdn_f = _fused_vs_parts(P, G, None, 3, pred, gt, None)
    Pg = P.clone().requires_grad_(True)
    ssi_g, hdn_g = d.ssi_hdn_dr(Pg, G, None, 3)
    assert ssi_g.requires_grad and hdn_g.requires_grad
    assert rel_scalar(ssi_g, ssi_f) <= 2e-5 and rel_scalar(hdn_g, hdn_f) <= 2e-5
    with pytest.raises(NotImplementedError):
        d.ssi_hdn_dr(P, G, None, 4)
