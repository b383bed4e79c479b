"""Pin the oracle: run the LIVE reference (``/root/reference``, stub-imported) on
seeded synthetic weights/inputs, compare the restatement with it, and write the
reference's outputs as small fixtures under ``tests/golden/``.

Run in the build container only:  ``python -m oracle.make_golden``

Fixtures hold *reference* outputs (not oracle outputs):
  golden_model.npz   depth / feature sub-samples for a handful of model cases
  golden_losses.npz  every loss scalar + sub-sampled aligned maps / contexts
Inputs are regenerated from seeds by ``distill_any_depth_b200.synthetic``.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
import oracle  # noqa: E402
from oracle.model import teacher_to_student_keys  # noqa: E402
from distill_any_depth_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, preset, B, H, W, weight seed, image seed, teacher-class?, head bias)
# The head bias keeps every depth map "live" (strictly positive): a relative-error metric is meaningless on
# pixels the final ReLU clips to zero.
MODEL_CASES = [
    ("vits_70x98", "vits", 2, 70, 98, 0, 1234, False, 0.25),
    ("vits_518", "vits", 1, 518, 518, 0, 1234, False, 0.25),     # BASELINE config 1
    ("vitb_112", "vitb", 2, 112, 112, 1, 1235, False, 0.25),
    ("vitl_teacher_70", "vitl", 1, 70, 70, 2, 1236, True, 0.6),
]
LOSS_CASES = [("l_2x64x64", 2, 64, 64, 7), ("l_3x56x84", 3, 56, 84, 8)]


def student_to_teacher_keys(sd):
    out = {}
    for k, v in sd.items():
        if k.startswith("pretrained.blocks."):
            k = "backbone.blocks.0." + k[len("pretrained.blocks."):]
        elif k.startswith("pretrained."):
            k = "backbone." + k[len("pretrained."):]
        out[k] = v
    return out


def sub(x, n=24):
    """Deterministic sub-sample: up to n evenly spaced indices along the last two dims."""
    h = torch.linspace(0, x.shape[-2] - 1, min(n, x.shape[-2])).round().long()
    w = torch.linspace(0, x.shape[-1] - 1, min(n, x.shape[-1])).round().long()
    return x[..., h, :][..., w].contiguous().numpy()


def main():
    assert refload.available(), "reference tree not found"
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    V2, Teacher = refload.load_models()
    L = refload.load_losses()
    os.makedirs(OUT, exist_ok=True)
    g, report = {}, {}

    for name, preset, B, H, W, ws, xs, teacher, hb in MODEL_CASES:
        kw = synthetic.MODEL_PRESETS[preset]
        sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
        x = synthetic.make_images(B, H, W, seed=xs)
        with torch.no_grad():
            if teacher:
                m = Teacher(**kw).eval()
                tsd = student_to_teacher_keys(sd)
                missing = m.load_state_dict(tsd, strict=True)
            else:
                m = V2(**kw).eval()
                missing = m.load_state_dict(sd, strict=True)
            d_ref, f_ref = m(x)
            d_or, f_or, pre = oracle.depth_anything_forward(x, sd, kw["encoder"], return_pre_relu=True)
        e_d = (d_ref - d_or).abs().max().item()
        e_f = (f_ref - f_or).abs().max().item()
        report[name] = dict(depth_max_abs=e_d, feat_max_abs=e_f, depth_mean=d_ref.mean().item(),
                            depth_min=d_ref.min().item(), depth_max=d_ref.max().item(),
                            zeros_frac=(d_ref == 0).float().mean().item())
        print(name, report[name], flush=True)
        assert e_d <= 1e-5 * max(1.0, d_ref.abs().max().item()), name
        assert e_f <= 1e-4, name
        g[name + "/depth_sub"] = sub(d_ref)
        g[name + "/feat_sub"] = sub(f_ref)
        g[name + "/depth_stats"] = np.array([d_ref.mean().item(), d_ref.abs().max().item(),
                                             d_ref.double().pow(2).sum().item()])
        g[name + "/feat_stats"] = np.array([f_ref.mean().item(), f_ref.abs().max().item(),
                                            f_ref.double().pow(2).sum().item()])
        del m
    np.savez_compressed(os.path.join(OUT, "golden_model.npz"), **g)

    g = {}
    ssi_ref, ssi_or = L.SSILoss(), oracle.SSILoss()
    for name, B, H, W, seed in LOSS_CASES:
        pred, gt, mask = synthetic.make_depth_pair(B, H, W, seed=seed)
        full = torch.ones_like(mask)
        fs = synthetic.make_features(B, 49, 96, seed=seed + 100)
        ft = synthetic.make_features(B, 49, 128, seed=seed + 200)
        # make a few exact ties / boundary values so lower-median and range tests are exercised
        pred[0, 0, 0, :8] = pred[0, 0, 1, :8]
        gt[0, 0, 2, 3] = gt[0].max()
        ref, orc = {}, {}
        for tag, mk in (("mask", mask), ("full", full)):
            pa, ga = L.masked_shift_and_scale(pred, gt, mk)
            po, go = oracle.masked_shift_and_scale(pred, gt, mk)
            ref[f"align_pred_{tag}"], orc[f"align_pred_{tag}"] = pa, po
            ref[f"align_gt_{tag}"], orc[f"align_gt_{tag}"] = ga, go
            ref[f"ssi_{tag}"], orc[f"ssi_{tag}"] = ssi_ref(pred, gt, mk), ssi_or(pred, gt, mk)
            ref[f"ssi_dense_{tag}"] = ssi_ref(pred, gt, mk, dense=True)
            orc[f"ssi_dense_{tag}"] = ssi_or(pred, gt, mk, dense=True)
            cr = L.get_contexts_dr(3, gt, mk)
            co = oracle.get_contexts_dr(3, gt, mk)
            assert torch.equal(cr, co), (name, tag, "contexts_dr")
            ref[f"ctx_dr_count_{tag}"] = cr.sum(0).float()
            orc[f"ctx_dr_count_{tag}"] = co.sum(0).float()
            ref[f"hdn_dr_{tag}"] = L.compute_hdn_loss(ssi_ref, pred, gt, cr)
            orc[f"hdn_dr_{tag}"] = oracle.compute_hdn_loss(ssi_or, pred, gt, co)
            cp_r, cp_o = L.get_contexts_dp(3, gt, mk), oracle.get_contexts_dp(3, gt, mk)
            assert torch.equal(cp_r, cp_o), (name, tag, "contexts_dp")
            ref[f"hdn_dp_{tag}"] = L.compute_hdn_loss(ssi_ref, pred, gt, cp_r)
            orc[f"hdn_dp_{tag}"] = oracle.compute_hdn_loss(ssi_or, pred, gt, cp_o)
            if H == W:
                cs_r, cs_o = L.get_contexts_ds(3, mk), oracle.get_contexts_ds(3, mk)
                assert torch.equal(cs_r, cs_o), (name, tag, "contexts_ds")
                ref[f"hdn_ds_{tag}"] = L.compute_hdn_loss(ssi_ref, pred, gt, cs_r)
                orc[f"hdn_ds_{tag}"] = oracle.compute_hdn_loss(ssi_or, pred, gt, cs_o)
        ref["ctx_dr_none"] = L.get_contexts_dr(3, gt, None).sum(0).float()
        orc["ctx_dr_none"] = oracle.get_contexts_dr(3, gt, None).sum(0).float()
        ref["grad"], orc["grad"] = L.gradient_preservation_loss(pred), oracle.gradient_preservation_loss(pred)
        ref["feat"], orc["feat"] = L.feature_distillation_loss(fs, ft), oracle.feature_distillation_loss(fs, ft)
        ref["feat_same"] = L.feature_distillation_loss(fs, fs * 0.5 + 0.1)
        orc["feat_same"] = oracle.feature_distillation_loss(fs, fs * 0.5 + 0.1)
        for st in ("global", "hybrid", "local", "none"):
            ref[f"distill_{st}"] = L.distillation_loss(pred, gt, st)
            orc[f"distill_{st}"] = oracle.distillation_loss(pred, gt, st)
        ref["norm_hybrid"], orc["norm_hybrid"] = L.hybrid_normalize(pred, 4), oracle.hybrid_normalize(pred, 4)
        ref["norm_global"], orc["norm_global"] = L.global_normalize(pred), oracle.global_normalize(pred)
        # degenerate inputs
        zero = torch.zeros_like(gt)
        half = torch.full_like(gt, 0.5)
        empty = torch.zeros_like(mask)
        for tag, gg, mk in (("allzero", zero, full), ("const", half, full), ("empty", gt, empty)):
            cr, co = L.get_contexts_dr(3, gg, mk), oracle.get_contexts_dr(3, gg, mk)
            assert torch.equal(cr, co), (name, tag)
            ref[f"hdn_dr_{tag}"] = L.compute_hdn_loss(ssi_ref, pred, gg, cr)
            orc[f"hdn_dr_{tag}"] = oracle.compute_hdn_loss(ssi_or, pred, gg, co)
            ref[f"ssi_{tag}"], orc[f"ssi_{tag}"] = ssi_ref(pred, gg, mk), ssi_or(pred, gg, mk)
        worst = 0.0
        for k in ref:
            r, o = ref[k].detach().float(), torch.as_tensor(orc[k]).detach().float()
            den = max(r.abs().max().item(), 1e-6)
            err = (r - o).abs().max().item() / den
            worst = max(worst, err)
            assert err <= 2e-5, (name, k, err)
            g[f"{name}/{k}"] = sub(r) if r.dim() >= 2 else r.numpy()
        report[name] = dict(worst_rel_err_vs_reference=worst,
                            scalars={k: float(v) for k, v in ref.items() if v.dim() == 0})
        print(name, "worst", worst, flush=True)
    np.savez_compressed(os.path.join(OUT, "golden_losses.npz"), **g)
    with open(os.path.join(OUT, "golden_report.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=report), f, indent=1, sort_keys=True)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
