"""Pin oracle/preprocess.py against the LIVE reference (build container only) and write
``tests/golden/golden_preprocess.npz``:   python -m oracle.make_golden_preprocess

Inputs are synthetic uint8 images from a seeded numpy PCG64 stream (`synthetic_image`), so the test can rebuild
them anywhere; the stored outputs are sub-sampled reference tensors (`DepthAnythingV2.image2tensor` of the reference
class, dpt.py:237-262, and its transform classes with keep_aspect_ratio=False as tools/testers/infer.py:173-177)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess as P, refload  # noqa: E402

CASES = [("img_480x640", 480, 640, 518, True, 21), ("img_375x1242", 375, 1242, 518, True, 22),
         ("img_700x500", 700, 500, 392, True, 23), ("img_518x518", 518, 518, 518, True, 24),
         ("img_300x400_square", 300, 400, 392, False, 25)]


def synthetic_image(h, w, seed):
    """Smooth-ish random uint8 BGR image (sum of a few sinusoids + noise), same on any box."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        for _ in range(4):
            fx, fy, ph = rng.uniform(0.005, 0.08), rng.uniform(0.005, 0.08), rng.uniform(0, 6.28)
            img[..., c] += np.sin(xx * fx + yy * fy + ph)
    img = (img - img.min()) / (img.max() - img.min()) * 215 + rng.uniform(0, 40, img.shape)
    return np.clip(np.round(img), 0, 255).astype(np.uint8)


def sub(x, n=32):
    h = np.linspace(0, x.shape[-2] - 1, min(n, x.shape[-2])).round().astype(int)
    w = np.linspace(0, x.shape[-1] - 1, min(n, x.shape[-1])).round().astype(int)
    return np.ascontiguousarray(x[..., h, :][..., w])


def main():
    assert refload.available()
    refload.install_stubs()
    sys.path.insert(0, refload.REF)
    import cv2
    from distillanydepth.depth_anything_v2.dpt import DepthAnythingV2 as V2
    from distillanydepth.depth_anything_v2.util.transform import Resize, NormalizeImage, PrepareForNet
    from torchvision.transforms import Compose
    m = V2(encoder="vits", features=64, out_channels=[48, 96, 192, 384])
    g, rep = {}, {}
    for name, h, w, size, keep, seed in CASES:
        raw = synthetic_image(h, w, seed)
        if keep:
            ref, hw = m.image2tensor(raw, size)
            ref = ref.cpu()
        else:  # tools/testers/infer.py:173-177 (keep_aspect_ratio=False) on an RGB float image
            tf = Compose([Resize(size, size, resize_target=False, keep_aspect_ratio=False, ensure_multiple_of=14,
                                 resize_method="lower_bound", image_interpolation_method=cv2.INTER_CUBIC),
                          NormalizeImage(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]), PrepareForNet()])
            ref = torch.from_numpy(tf({"image": raw[..., ::-1] / 255})["image"]).unsqueeze(0)
            hw = (h, w)
        orc, hw2 = P.image2tensor(raw, size, keep_aspect_ratio=keep)
        assert tuple(hw) == tuple(hw2) and ref.shape == orc.shape, (name, ref.shape, orc.shape)
        err = (ref - orc).abs().max().item()
        assert err == 0.0, (name, err)
        g[name + "/tensor_sub"] = sub(ref.numpy())
        g[name + "/shape"] = np.array(ref.shape)
        g[name + "/stats"] = np.array([ref.double().mean().item(), ref.abs().max().item()])
        rep[name] = dict(shape=list(ref.shape), max_abs_diff_oracle_vs_reference=err)
        print(name, tuple(ref.shape), "oracle == reference", flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_preprocess.npz"), **g)
    with open(os.path.join(ROOT, "tests", "golden", "golden_preprocess_report.json"), "w") as f:
        json.dump(rep, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
