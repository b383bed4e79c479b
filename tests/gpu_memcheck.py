"""compute-sanitizer target (not a pytest file): one small pass through every CUDA entry point of the hot path -
both precision modes, both attention kernels, ViT-B (N = 128 decoder, generic ConvT scatter) and a ViT-L-width head
(N = 256, 2-CTA convs, TMA ConvT scatter), all losses, pre / post-processing.
usage (GPU box): compute-sanitizer --tool memcheck python tests/gpu_memcheck.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import distill_any_depth_b200 as d
from distill_any_depth_b200 import synthetic, preprocess


def model(preset, **over):
    kw = dict(synthetic.MODEL_PRESETS[preset])
    kw.update(over)
    m = d.DepthAnythingV2(**kw)
    m.load_state_dict(synthetic.make_state_dict(seed=1, **kw), strict=True)
    return m.cuda().eval()


x = synthetic.make_images(2, 70, 98, seed=3).cuda()
for preset, over in (("vits", {}), ("vitb", {}), ("vits", dict(features=256, out_channels=[256, 512, 1024, 1024]))):
    m = model(preset, **over)
    for prec in ("fp32", "bf16"):
        for variant in ("2", "3"):
            os.environ["DAD_ATT_VARIANT"] = variant
            m.precision = prec
            depth, feat = m(x)
    torch.cuda.synchronize()
    print("forward ok", preset, over, float(depth.mean()), flush=True)

pred, gt, mask = synthetic.make_depth_pair(3, 56, 84, seed=9)
P, G, M = pred.cuda(), gt.cuda(), mask.cuda()
ssi = d.SSILoss()
vals = [ssi(P, G, M), ssi(P, G, M, dense=True).sum(), d.compute_hdn_loss(ssi, P, G, d.get_contexts_dr(3, G, M)),
        d.compute_hdn_loss(ssi, P, G, d.get_contexts_dr(3, G, M).clone()),
        d.compute_hdn_loss(ssi, P, G, d.get_contexts_dp(3, G, M)), d.gradient_preservation_loss(P),
        d.feature_distillation_loss(synthetic.make_features(3, 49, 96).cuda(), synthetic.make_features(3, 49, 128).cuda())]
sq = torch.rand(2, 1, 64, 64, device="cuda")
vals.append(d.compute_hdn_loss(ssi, sq, sq * 0.5 + 0.1, d.get_contexts_ds(3, torch.ones_like(sq, dtype=torch.bool))))
for st in ("none", "global", "hybrid"):
    vals.append(d.distillation_loss(P, G, st))
const = torch.full_like(G, 0.0)   # degenerate: candidate lists overflow, the streaming select runs
vals.append(d.compute_hdn_loss(ssi, P, const, d.get_contexts_dr(3, const, None)))
vals.append(ssi(P, const, M))
torch.cuda.synchronize()
print("losses ok", [round(float(v), 5) for v in vals], flush=True)

raw = np.random.default_rng(0).integers(0, 256, (60, 90, 3), dtype=np.uint8)
t, hw = preprocess.image_to_tensor(raw, 70, device="cuda")
back = preprocess.resize_depth(torch.rand(1, 1, t.shape[2], t.shape[3], device="cuda"), hw)
nm = preprocess.normalize_minmax(back)
torch.cuda.synchronize()
print("preprocess ok", tuple(t.shape), tuple(back.shape), float(nm.max()), flush=True)
