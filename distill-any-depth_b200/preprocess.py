"""GPU pre- / post-processing either side of the forward: the reference does these per image on the
CPU with cv2 / numpy (``depth_anything_v2/dpt.py:227-262``, ``util/transform.py:5-148``, ``tools/testers/infer.py:125-147``);
here the uint8 image goes to the device as is (3x fewer H2D bytes than the fp32 CHW tensor) and one kernel does
``/255 -> cv2.resize(INTER_CUBIC) -> (x - mean) / std -> CHW`` in the reference's float64 arithmetic.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _constrain(x, multiple_of, min_val=0, max_val=None):
    """``Resize.constrain_to_multiple_of`` (util/transform.py:52-61); np.round = round-half-to-even."""
    y = int(np.round(x / multiple_of) * multiple_of)
    if max_val is not None and y > max_val:
        y = int(math.floor(x / multiple_of) * multiple_of)
    if y < min_val:
        y = int(math.ceil(x / multiple_of) * multiple_of)
    return y


def get_size(width, height, target_width, target_height, keep_aspect_ratio=True, ensure_multiple_of=14,
             resize_method="lower_bound"):
    """``Resize.get_size`` (util/transform.py:63-106) -> ``(new_width, new_height)``."""
    scale_height = target_height / height
    scale_width = target_width / width
    if keep_aspect_ratio:
        if resize_method == "lower_bound":
            if scale_width > scale_height:
                scale_height = scale_width
            else:
                scale_width = scale_height
        elif resize_method == "upper_bound":
            if scale_width < scale_height:
                scale_height = scale_width
            else:
                scale_width = scale_height
        elif resize_method == "minimal":
            if abs(1 - scale_width) < abs(1 - scale_height):
                scale_height = scale_width
            else:
                scale_width = scale_height
        else:
            raise ValueError(f"resize_method {resize_method} not implemented")
    if resize_method == "lower_bound":
        nh = _constrain(scale_height * height, ensure_multiple_of, min_val=target_height)
        nw = _constrain(scale_width * width, ensure_multiple_of, min_val=target_width)
    elif resize_method == "upper_bound":
        nh = _constrain(scale_height * height, ensure_multiple_of, max_val=target_height)
        nw = _constrain(scale_width * width, ensure_multiple_of, max_val=target_width)
    elif resize_method == "minimal":
        nh = _constrain(scale_height * height, ensure_multiple_of)
        nw = _constrain(scale_width * width, ensure_multiple_of)
    else:
        raise ValueError(f"resize_method {resize_method} not implemented")
    return nw, nh


def image_to_tensor(raw_image, input_size=518, device="cuda", bgr=True, keep_aspect_ratio=True,
                    mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """uint8 ``[h, w, 3]`` image (numpy or tensor; BGR as ``cv2.imread`` gives it when ``bgr``) ->
    ``(tensor [1, 3, nh, nw] fp32 on device, (h, w))`` as ``DepthAnythingV2.image2tensor`` (dpt.py:237-262)."""
    if isinstance(raw_image, np.ndarray):
        if raw_image.dtype != np.uint8 or raw_image.ndim != 3 or raw_image.shape[2] != 3:
            raise ValueError("expected a uint8 [h, w, 3] image")
        raw = torch.from_numpy(np.ascontiguousarray(raw_image))
    else:
        raw = raw_image
        if raw.dtype != torch.uint8 or raw.dim() != 3 or raw.shape[2] != 3:
            raise ValueError("expected a uint8 [h, w, 3] image")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("the B200 preprocessing path runs on CUDA only (no CPU fallback)")
    raw = raw.to(dev, non_blocking=True).contiguous()
    h, w = int(raw.shape[0]), int(raw.shape[1])
    nw, nh = get_size(w, h, input_size, input_size, keep_aspect_ratio=keep_aspect_ratio)
    if out is None:
        out = torch.empty(1, 3, nh, nw, dtype=torch.float32, device=dev)
    m3 = (ctypes.c_double * 3)(*mean)
    s3 = (ctypes.c_double * 3)(*std)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().dad_preprocess_image(_lib.ptr(raw), h, w, 3 * w, 1 if bgr else 0, nh, nw, m3, s3,
                                                    _lib.ptr(out), _lib.stream_ptr()), "image_to_tensor")
    return out, (h, w)


def resize_depth(depth, size):
    """``F.interpolate(depth, size, mode="bilinear", align_corners=True)`` for ``[B, 1, H, W]`` fp32 maps (dpt.py:233)."""
    if not depth.is_cuda:
        raise RuntimeError("resize_depth: CUDA tensors only (no CPU fallback)")
    if depth.dim() != 4 or depth.shape[1] != 1:
        raise ValueError("resize_depth expects [B, 1, H, W]")
    d = depth.detach().float().contiguous()
    B, _, H, W = d.shape
    h, w = int(size[0]), int(size[1])
    out = torch.empty(B, 1, h, w, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().dad_resize_depth(_lib.ptr(d), B, H, W, h, w, _lib.ptr(out), _lib.stream_ptr()), "resize_depth")
    return out


def normalize_minmax(depth):
    """Per image ``(d - d.min()) / (d.max() - d.min())`` (tools/testers/infer.py:135), any ``[B, ...]`` fp32 tensor."""
    if not depth.is_cuda:
        raise RuntimeError("normalize_minmax: CUDA tensors only (no CPU fallback)")
    d = depth.detach().float().contiguous()
    B = d.shape[0]
    L = d.numel() // B
    out = torch.empty_like(d)
    ws = torch.empty(8 * B, dtype=torch.uint8, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().dad_minmax_normalize(_lib.ptr(d), B, L, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                                    _lib.stream_ptr()), "normalize_minmax")
    return out


# ------------------------------------------------------------------------------------------ colourise (8f N2)
# matplotlib's "Spectral" colormap data (matplotlib/_cm.py `_Spectral_data`: the 11-class ColorBrewer Spectral scheme,
# RGB / 255) restated here because matplotlib is not a dependency of this package.
_SPECTRAL = ((158, 1, 66), (213, 62, 79), (244, 109, 67), (253, 174, 97), (254, 224, 139), (255, 255, 191),
             (230, 245, 152), (171, 221, 164), (102, 194, 165), (50, 136, 189), (94, 79, 162))
_lut_cache = {}


def colormap_lut(cmap="Spectral", N=256):
    """256 x 3 float64 lookup table of a ``LinearSegmentedColormap.from_list`` colormap, computed as matplotlib's
    ``colors._create_lookup_table`` does (nodes at ``linspace(0, 1, len(colors))``, linear interpolation in float64,
    clip to [0, 1]); ``name_r`` reverses the node order (``Colormap.reversed``)."""
    name, rev = (cmap[:-2], True) if cmap.endswith("_r") else (cmap, False)
    if name != "Spectral":
        raise NotImplementedError(f"colormap {cmap!r}: only Spectral / Spectral_r (the reference's choice, "
                                  "tools/testers/infer.py:137) are built in")
    cols = np.asarray(_SPECTRAL, dtype=np.float64) / 255.0
    if rev:
        cols = cols[::-1]
    x = np.linspace(0.0, 1.0, len(cols)) * (N - 1)
    xind = (N - 1) * np.linspace(0.0, 1.0, N)
    ind = np.searchsorted(x, xind)[1:-1]
    dist = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
    lut = np.empty((N, 3), dtype=np.float64)
    for c in range(3):
        y = cols[:, c]
        lut[:, c] = np.concatenate([[y[0]], dist * (y[ind] - y[ind - 1]) + y[ind - 1], [y[-1]]])
    return np.clip(lut, 0.0, 1.0)


def _device_lut(cmap, device):
    key = (cmap, str(device))
    t = _lut_cache.get(key)
    if t is None:
        lut64 = colormap_lut(cmap)
        t = (torch.from_numpy(lut64.astype(np.float32)).to(device).contiguous(),
             torch.from_numpy((lut64 * 255).astype(np.uint8)).to(device).contiguous())   # numpy's float64 product, truncated
        _lut_cache[key] = t
    return t


def colorize_depth_maps(depth_map, min_depth=None, max_depth=None, cmap="Spectral", valid_mask=None, as_uint8_hwc=False):
    """``colorize_depth_maps`` (distillanydepth/utils/image_util.py:69-118) on the device: ``[(B,) H, W]`` (any singleton
    dims squeezed, as upstream) -> float ``[B, 3, H, W]`` in [0, 1]; ``min_depth == max_depth`` (including both ``None``)
    colours everything with the first LUT entry, as upstream's ``depth * 0``.  With ``as_uint8_hwc`` the same pass also
    returns ``(rgb * 255).astype(uint8)`` as ``[B, H, W, 3]`` (tools/testers/infer.py:139-140).  The result stays on the
    GPU (upstream returns a CPU tensor)."""
    if not isinstance(depth_map, torch.Tensor) or not depth_map.is_cuda:
        raise RuntimeError("colorize_depth_maps: CUDA tensors only (no CPU fallback)")
    assert depth_map.dim() >= 2, "Invalid dimension"
    d = depth_map.detach().squeeze().float().contiguous()
    if d.dim() < 3:
        d = d[None]
    if d.dim() != 3:
        raise ValueError("colorize_depth_maps expects [(B,) H, W] after squeezing")
    B, H, W = d.shape
    valid = None
    if valid_mask is not None:
        valid = valid_mask.detach().to(d.device).squeeze()
        valid = (valid[None] if valid.dim() < 3 else valid).expand(B, H, W)
        valid = (valid != 0).contiguous().view(torch.uint8)
    degenerate = 1 if min_depth == max_depth else 0
    lo, hi = (0.0, 0.0) if degenerate else (float(min_depth), float(max_depth))
    lut, lut8 = _device_lut(cmap, d.device)
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=d.device)
    out8 = torch.empty(B, H, W, 3, dtype=torch.uint8, device=d.device) if as_uint8_hwc else None
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().dad_colorize_depth(_lib.ptr(d), _lib.ptr(valid), B, H * W, lo, hi, degenerate, _lib.ptr(lut),
                                                  _lib.ptr(lut8), _lib.ptr(out), _lib.ptr(out8), _lib.stream_ptr()),
                   "colorize_depth_maps")
    return (out, out8) if as_uint8_hwc else out
