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
