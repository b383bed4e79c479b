"""Pin the oracle's OPTION branches against the LIVE reference: the ``use_clstoken`` readout
(dpt.py:116-122, 153-156), the ViT-g / SwiGLU encoder (dinov2.py:381-395, dinov2_layers/swiglu_ffn.py) and ``use_bn``
(util/blocks.py:49-51, eval mode).

Run in the build container only:  ``python -m oracle.make_golden_options``

Writes ``tests/golden/golden_model_options.npz`` (reference outputs, sub-sampled, same layout as golden_model.npz) and
``tests/golden/golden_options_report.json``.  Weights / inputs are regenerated from seeds by
``distill_any_depth_b200.synthetic`` (ViT-g: 1.26e9 parameters, ~5 GB of host memory while the case runs).
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
from oracle.make_golden import sub  # noqa: E402
from distill_any_depth_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, model kwargs, B, H, W, weight seed, image seed, head bias)
OPTION_CASES = [
    ("vits_clstoken_70x98", dict(synthetic.MODEL_PRESETS["vits"], use_clstoken=True), 2, 70, 98, 3, 1237, 0.25),
    ("vitb_clstoken_112", dict(synthetic.MODEL_PRESETS["vitb"], use_clstoken=True), 1, 112, 112, 5, 1239, 0.25),
    ("vits_bn_70x98", dict(synthetic.MODEL_PRESETS["vits"], use_bn=True), 2, 70, 98, 8, 1240, 0.25),
    # ViT-g encoder on the ViT-B head shape (keeps the case small); the giant head preset is a GPU-only test
    ("vitg_70x98", dict(encoder="vitg", features=128, out_channels=[96, 192, 384, 768]), 1, 70, 98, 4, 1238, 0.25),
]


def main():
    assert refload.available(), "reference tree not found"
    torch.set_num_threads(os.cpu_count())
    V2, _ = refload.load_models()
    os.makedirs(OUT, exist_ok=True)
    g, report = {}, {}
    for name, kw, B, H, W, ws, xs, hb in OPTION_CASES:
        sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
        x = synthetic.make_images(B, H, W, seed=xs)
        with torch.no_grad():
            m = V2(**kw).eval()
            m.load_state_dict(sd, strict=True)
            d_ref, f_ref = m(x)
            del m
            d_or, f_or = oracle.depth_anything_forward(x, sd, kw["encoder"])
        e_d = (d_ref - d_or).abs().max().item()
        e_f = (f_ref - f_or).abs().max().item()
        report[name] = dict(depth_max_abs=e_d, feat_max_abs=e_f, depth_mean=d_ref.mean().item(), depth_min=d_ref.min().item(),
                            depth_max=d_ref.max().item(), zeros_frac=(d_ref == 0).float().mean().item())
        print(name, report[name], flush=True)
        assert e_d <= 1e-5 * max(1.0, d_ref.abs().max().item()), name
        assert e_f <= 1e-4, name
        g[name + "/depth_sub"] = sub(d_ref)
        g[name + "/feat_sub"] = sub(f_ref)
        g[name + "/depth_stats"] = np.array([d_ref.mean().item(), d_ref.abs().max().item(), d_ref.double().pow(2).sum().item()])
        g[name + "/feat_stats"] = np.array([f_ref.mean().item(), f_ref.abs().max().item(), f_ref.double().pow(2).sum().item()])
        del sd
    np.savez_compressed(os.path.join(OUT, "golden_model_options.npz"), **g)
    with open(os.path.join(OUT, "golden_options_report.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=report), f, indent=1, sort_keys=True)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
