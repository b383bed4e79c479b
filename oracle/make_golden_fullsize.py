"""Pin the oracle at the HEADLINE sizes: run the LIVE reference (``/root/reference``, stub-imported) on ViT-L at
518x518 (BASELINE configs[2], the bench workload) and 1036x1036 (configs[4], 5477 tokens), compare the restatement
with it over the FULL maps, and write the reference's outputs as a fixture under ``tests/golden/``.

Run in the build container only:  ``python -m oracle.make_golden_fullsize``   (about 2 minutes on 8 cores)

The fixture holds *reference* outputs: an evenly spaced 96x96 sub-sample of each depth map, a 64x64 sub-sample of
the feature map and full-map statistics.  The GPU tests (tests/test_gpu_fullsize.py) compare the device's FULL map
with the oracle evaluated on the same box (seconds of CPU per image) and the oracle's sub-sample with this fixture,
which closes the chain  device == oracle == live reference  at the sizes the benchmark runs at.
Reference: depth_anything_v2/dpt.py:211-225, dinov2.py:271-321, dinov2_layers/attention.py:49-62.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402
import oracle  # noqa: E402
from distill_any_depth_b200 import synthetic  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, preset, B, H, W, weight seed, image seed, head bias) - the seeds tests/test_gpu_fullsize.py uses
FULLSIZE_CASES = [
    ("vitl_518", "vitl", 2, 518, 518, 1, 1234, 0.25),     # first two images of the batch-32 bench / test input
    ("vitl_1036", "vitl", 1, 1036, 1036, 1, 99, 0.25),
]
DEPTH_SUB, FEAT_SUB = 96, 64


def sub(x, n):
    h = torch.linspace(0, x.shape[-2] - 1, min(n, x.shape[-2])).round().long()
    w = torch.linspace(0, x.shape[-1] - 1, min(n, x.shape[-1])).round().long()
    return x[..., h, :][..., w].contiguous().numpy()


def main():
    assert refload.available(), "reference tree not found"
    torch.set_num_threads(os.cpu_count())
    V2, _ = refload.load_models()
    os.makedirs(OUT, exist_ok=True)
    g, report = {}, {}
    for name, preset, B, H, W, ws, xs, hb in FULLSIZE_CASES:
        kw = synthetic.MODEL_PRESETS[preset]
        sd = synthetic.make_state_dict(seed=ws, head_bias=hb, **kw)
        x = synthetic.make_images(B, H, W, seed=xs)
        with torch.no_grad():
            m = V2(**kw).eval()
            m.load_state_dict(sd, strict=True)
            t0 = time.time()
            d_ref, f_ref = m(x)
            t1 = time.time()
            d_or, f_or = oracle.depth_anything_forward(x, sd, kw["encoder"])
            t2 = time.time()
        e_d = (d_ref - d_or).abs().max().item()
        e_f = (f_ref - f_or).abs().max().item()
        report[name] = dict(depth_max_abs=e_d, feat_max_abs=e_f, depth_mean=d_ref.mean().item(),
                            depth_min=d_ref.min().item(), depth_max=d_ref.max().item(),
                            zeros_frac=(d_ref == 0).float().mean().item(),
                            reference_seconds=round(t1 - t0, 1), oracle_seconds=round(t2 - t1, 1),
                            threads=torch.get_num_threads())
        print(name, report[name], flush=True)
        assert e_d <= 1e-5 * max(1.0, d_ref.abs().max().item()), name
        assert e_f <= 1e-4, name
        g[name + "/depth_sub"] = sub(d_ref, DEPTH_SUB)
        g[name + "/feat_sub"] = sub(f_ref, FEAT_SUB)
        g[name + "/depth_stats"] = np.array([d_ref.mean().item(), d_ref.abs().max().item(),
                                             d_ref.double().pow(2).sum().item(), d_ref.min().item()])
        g[name + "/feat_stats"] = np.array([f_ref.mean().item(), f_ref.abs().max().item(),
                                            f_ref.double().pow(2).sum().item()])
        del m
    np.savez_compressed(os.path.join(OUT, "golden_model_fullsize.npz"), **g)
    with open(os.path.join(OUT, "golden_fullsize_report.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=report), f, indent=1, sort_keys=True)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
