"""Oracle: the reference's CPU pre- / post-processing, restated with the same third-party calls it makes
(cv2.cvtColor / cv2.resize(INTER_CUBIC) on float64, numpy float64 normalisation, torch F.interpolate).

Reference: ``distillanydepth/depth_anything_v2/dpt.py:227-262`` (image2tensor, infer_image),
``depth_anything_v2/util/transform.py:5-148`` (Resize / NormalizeImage / PrepareForNet),
``tools/testers/infer.py:125-147`` (min-max normalisation of the prediction).

TEST INFRASTRUCTURE: see oracle/__init__.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _constrain(x, m, min_val=0, max_val=None):
    """transform.py:52-61."""
    y = (np.round(x / m) * m).astype(int)
    if max_val is not None and y > max_val:
        y = (np.floor(x / m) * m).astype(int)
    if y < min_val:
        y = (np.ceil(x / m) * m).astype(int)
    return int(y)


def get_size(width, height, tw, th, keep_aspect_ratio=True, multiple_of=14):
    """transform.py:63-106, 'lower_bound' (the only method on the path)."""
    sh, sw = np.float64(th) / height, np.float64(tw) / width
    if keep_aspect_ratio:
        if sw > sh:
            sh = sw
        else:
            sw = sh
    return _constrain(sw * width, multiple_of, min_val=tw), _constrain(sh * height, multiple_of, min_val=th)


def image2tensor(raw_image, input_size=518, keep_aspect_ratio=True):
    """dpt.py:237-262 -> (float32 [1, 3, nh, nw] CPU tensor, (h, w))."""
    import cv2
    h, w = raw_image.shape[:2]
    image = cv2.cvtColor(raw_image, cv2.COLOR_BGR2RGB) / 255.0
    nw, nh = get_size(image.shape[1], image.shape[0], input_size, input_size, keep_aspect_ratio)
    image = cv2.resize(image, (nw, nh), interpolation=cv2.INTER_CUBIC)
    image = (image - [0.485, 0.456, 0.406]) / [0.229, 0.224, 0.225]
    image = np.ascontiguousarray(np.transpose(image, (2, 0, 1))).astype(np.float32)
    return torch.from_numpy(image).unsqueeze(0), (h, w)


def resize_depth(depth, size):
    """dpt.py:233 (``depth`` given as [B, 1, H, W])."""
    return F.interpolate(depth, size, mode="bilinear", align_corners=True)


def normalize_minmax(depth):
    """tools/testers/infer.py:135, per image."""
    flat = depth.reshape(depth.shape[0], -1)
    mn, mx = flat.min(1).values, flat.max(1).values
    shp = (-1,) + (1,) * (depth.dim() - 1)
    return (depth - mn.reshape(shp)) / (mx - mn).reshape(shp)


def colorize_depth_maps(depth_map, min_depth=None, max_depth=None, lut=None, valid_mask=None):
    """Restatement of ``colorize_depth_maps`` (distillanydepth/utils/image_util.py:69-118) with matplotlib's
    ``Colormap.__call__`` spelled out (matplotlib is not installed in this image: PARITY UNPINNED for the colormap table
    itself - ``lut`` is the 256 x 3 float64 table of ``distill_any_depth_b200.preprocess.colormap_lut``, which restates
    matplotlib's ``_create_lookup_table`` on the published ColorBrewer Spectral data).  numpy in, ``[B, 3, H, W]`` float64
    out; ``xa = depth * N; xa[xa == N] = N - 1; astype(int)`` as in matplotlib/colors.py."""
    depth = np.array(depth_map, copy=True).squeeze()
    if depth.ndim < 3:
        depth = depth[np.newaxis, :, :]
    if min_depth != max_depth:
        depth = ((depth - min_depth) / (max_depth - min_depth)).clip(0, 1)
    else:
        depth = depth * 0.
    N = lut.shape[0]
    xa = np.array(depth, copy=True)
    xa *= N
    xa[xa == N] = N - 1
    idx = np.clip(xa, 0, N - 1).astype(int)
    img = np.rollaxis(lut[idx], 3, 1).copy()
    if valid_mask is not None:
        vm = np.asarray(valid_mask).squeeze()
        vm = vm[np.newaxis, np.newaxis] if vm.ndim < 3 else vm[:, np.newaxis]
        vm = np.repeat(np.broadcast_to(vm, (img.shape[0], 1) + img.shape[2:]), 3, axis=1)
        img[~vm] = 0
    return img
