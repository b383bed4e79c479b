"""GPU pre- / post-processing against the oracle (cv2 / numpy / torch on the CPU) and the
committed reference fixtures.  Tolerances: the cubic resize accumulates in float64 like OpenCV, the result is
rounded to fp32 once - outputs must agree to 2e-6 absolute (values are O(1)); bilinear depth resize and the
min-max normalisation to 1e-6 relative."""
import os

import numpy as np
import pytest
import torch

import oracle
from distill_any_depth_b200 import synthetic
from oracle import preprocess as P
from oracle.make_golden_preprocess import CASES, synthetic_image, sub as psub

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_preprocess.npz")


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_image_to_tensor_matches_oracle_and_fixture(case):
    from distill_any_depth_b200 import preprocess
    name, h, w, size, keep, seed = case
    raw = synthetic_image(h, w, seed)
    got, hw = preprocess.image_to_tensor(raw, size, device="cuda", bgr=True, keep_aspect_ratio=keep)
    ref, hw2 = P.image2tensor(raw, size, keep_aspect_ratio=keep)
    assert hw == hw2 == (h, w) and tuple(got.shape) == tuple(ref.shape)
    err = (got.cpu() - ref).abs().max().item()
    assert err <= 2e-6, err
    g = dict(np.load(GOLD))
    assert np.abs(psub(got.cpu().numpy()) - g[name + "/tensor_sub"]).max() <= 2e-6


def test_image_to_tensor_edge_cases():
    from distill_any_depth_b200 import preprocess
    rng = np.random.Generator(np.random.PCG64(5))
    for h, w, size in [(14, 14, 518), (37, 1000, 70), (1036, 1554, 1036), (519, 517, 518)]:
        raw = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)   # white noise: the worst case for cubic overshoot
        got, _ = preprocess.image_to_tensor(raw, size, device="cuda")
        ref, _ = P.image2tensor(raw, size)
        assert tuple(got.shape) == tuple(ref.shape)
        assert got.shape[2] % 14 == 0 and got.shape[3] % 14 == 0
        assert (got.cpu() - ref).abs().max().item() <= 4e-6
    with pytest.raises(ValueError):
        preprocess.image_to_tensor(np.zeros((4, 4), dtype=np.uint8), 518, device="cuda")
    with pytest.raises(RuntimeError):
        preprocess.image_to_tensor(np.zeros((4, 4, 3), dtype=np.uint8), 518, device="cpu")


def test_depth_postprocessing():
    from distill_any_depth_b200 import preprocess
    g = torch.Generator().manual_seed(3)
    d = torch.rand(2, 1, 518, 686, generator=g) * 7 - 1
    for size in [(480, 640), (1, 1), (518, 686), (1000, 37)]:
        got = preprocess.resize_depth(d.cuda(), size)
        ref = P.resize_depth(d, size)
        assert tuple(got.shape) == tuple(ref.shape)
        assert (got.cpu() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    got = preprocess.normalize_minmax(d.cuda())
    ref = P.normalize_minmax(d)
    assert (got.cpu() - ref).abs().max().item() <= 1e-6
    assert float(got.min()) == 0.0 and float(got.max()) == 1.0


def test_infer_image_end_to_end():
    """raw BGR image -> GPU preprocessing -> forward (fp32 verification mode) -> GPU resize back, against the same
    pipeline assembled from the oracle's pieces (dpt.py:227-235)."""
    import distill_any_depth_b200 as d
    kw = synthetic.MODEL_PRESETS["vits"]
    sd = synthetic.make_state_dict(seed=0, **kw)
    m = d.DepthAnythingV2(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = "fp32"
    raw = synthetic_image(120, 200, 31)
    got = m.infer_image(raw, input_size=154)
    x, (h, w) = P.image2tensor(raw, 154)
    with torch.no_grad():
        dref, _ = oracle.depth_anything_forward(x, sd, "vits")
    ref = P.resize_depth(dref, (h, w))[0, 0].numpy()
    assert got.shape == (120, 200) and got.dtype == np.float32
    den = np.maximum(np.abs(ref), 0.1 * np.abs(ref).max())
    assert (np.abs(got - ref) / den).max() <= 1e-4


@pytest.mark.parametrize("cmap", ["Spectral", "Spectral_r"])
def test_colorize_depth_maps_matches_oracle(cmap):
    """colorize_depth_maps (utils/image_util.py:69-118) + the uint8 HWC image of tools/testers/infer.py:137-140: the LUT
    index is integer work - bit-exact against the numpy restatement, including x == 0, x == 1, bin edges k / 256, values
    outside [min, max], a validity mask and the degenerate min == max branch."""
    from distill_any_depth_b200 import preprocess
    lut = preprocess.colormap_lut(cmap)
    rng = np.random.Generator(np.random.PCG64(3))
    d = rng.random((2, 61, 45), dtype=np.float32)
    d[0, 0, :8] = [0.0, 1.0, 0.5, 0.25, 255.0 / 256.0, 1.0 / 256.0, -0.3, 1.7]
    d[1, 3, :4] = np.nextafter(np.float32([0.5, 0.25, 0.75, 1.0]), np.float32(0))
    mask = rng.random((2, 61, 45)) > 0.2
    dt, mt = torch.from_numpy(d).cuda(), torch.from_numpy(mask).cuda()
    for lo, hi, vm, vmt in ((0.0, 1.0, None, None), (0.1, 0.8, mask, mt), (None, None, None, None), (0.3, 0.3, mask, mt)):
        got, got8 = preprocess.colorize_depth_maps(dt[:, None], lo, hi, cmap=cmap, valid_mask=vmt, as_uint8_hwc=True)
        ref = P.colorize_depth_maps(d[:, None], lo, hi, lut=lut, valid_mask=vm)
        assert tuple(got.shape) == ref.shape == (2, 3, 61, 45)
        assert np.array_equal(got.cpu().numpy(), ref.astype(np.float32)), (lo, hi)
        ref8 = (ref * 255).astype(np.uint8).transpose(0, 2, 3, 1)
        # the uint8 image is numpy's float64 product truncated (a host-computed uint8 table indexed on the device)
        assert np.array_equal(got8.cpu().numpy(), ref8), (lo, hi)
    single = preprocess.colorize_depth_maps(dt[0], 0.0, 1.0, cmap=cmap)   # [H, W] input -> [1, 3, H, W]
    assert tuple(single.shape) == (1, 3, 61, 45)
    with pytest.raises(NotImplementedError):
        preprocess.colorize_depth_maps(dt, 0.0, 1.0, cmap="magma")
