"""Host-side mirror of the reference's loss functions (``tools/train_distillation.py:173-707``):
same names, positional signatures and return conventions, computed by the sm_100a kernels in
``csrc/losses.cu`` through the C ABI.  No host sync, no per-image Python loop; every function
returns device tensors.  CUDA tensors only - there is no CPU fallback.
"""
import functools
import weakref

import torch
import torch.nn as nn

from . import _lib

_STRATEGY = {"none": 0, "global": 1, "hybrid": 2, "local": 2}
_ws_cache = {}


def _first_cuda_device(objs):
    for a in objs:
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (list, tuple)):
            d = _first_cuda_device(a)
            if d is not None:
                return d
    return None


def _on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its first tensor argument current: the C ABI launches on the current device
    and ``_lib.stream_ptr()`` / ``_workspace`` use that device's current stream, so tensors living on ``cuda:1`` while
    ``cuda:0`` is current (the reference's ``.to(device)`` usage) get the right device, stream and scratch buffer."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _first_cuda_device(list(args) + list(kwargs.values()))
        if dev is None:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def _workspace(device, rows, K):
    need = int(_lib.load().dad_loss_workspace_bytes(int(rows), int(K)))
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)  # one scratch buffer per stream
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 loss path has no CPU fallback")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _mask_u8(m, like, name="mask_valid"):
    if m is None:
        return None
    if m.shape != like.shape:
        m = m.expand_as(like)
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8)
    return (m != 0).contiguous().view(torch.uint8)


def _rows_L(t):
    if t.dim() < 3:
        raise ValueError("expected a [B, C, H, W] map")
    rows = 1
    for s in t.shape[:-2]:
        rows *= s
    return rows, t.shape[-2] * t.shape[-1]


def _new_scalar(device):
    return torch.empty((), dtype=torch.float32, device=device)


def _new_partials(device, want):
    return torch.empty(2, dtype=torch.float64, device=device) if want else None


# ------------------------------------------------------------------------------------------ SSI
@_on_tensor_device
def masked_shift_and_scale(depth_preds, depth_gt, mask_valid):
    """``:449-533`` - per (b, c) lower-median / mean-absolute-deviation alignment.
    Returns ``(depth_pred_aligned, depth_gt_aligned)``."""
    p, g = _f32(depth_preds, "depth_preds"), _f32(depth_gt, "depth_gt")
    rows, L = _rows_L(p)
    m = _mask_u8(mask_valid, p)
    pa, ga = torch.empty_like(p), torch.empty_like(g)
    ws = _workspace(p.device, rows, 1)
    lib = _lib.load()
    _lib.check(lib.dad_masked_shift_and_scale(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), rows, L, _lib.ptr(pa),
                                              _lib.ptr(ga), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
               "masked_shift_and_scale")
    return pa, ga


def masked_l1_loss(preds, target, mask_valid, dense=False):
    """``:535-542`` (elementwise; kept in torch ops on the device - it is a single fused pass when
    reached through :class:`SSILoss`)."""
    mv = mask_valid if mask_valid.dtype == torch.bool else (mask_valid != 0)
    e = (preds - target).abs()
    e = torch.where(mv, e, torch.zeros((), dtype=e.dtype, device=e.device))   # reference: loss[~mask] = 0 (NaN-safe)
    if dense:
        return e
    return e.sum() / (mask_valid.sum() + 1e-6)


@_on_tensor_device
def _ssi(depth_preds, depth_gt, mask_valid, dense, want_partials=False):
    p, g = _f32(depth_preds, "depth_preds"), _f32(depth_gt, "depth_gt")
    rows, L = _rows_L(p)
    m = _mask_u8(mask_valid, p)
    ws = _workspace(p.device, rows, 1)
    lib = _lib.load()
    dense_out = torch.empty_like(p) if dense else None
    out = None if dense else _new_scalar(p.device)
    part = _new_partials(p.device, want_partials and not dense)
    _lib.check(lib.dad_ssi_loss(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), rows, L, _lib.ptr(dense_out), _lib.ptr(out),
                                _lib.ptr(part), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "SSILoss")
    return (dense_out if dense else out), part


# ------------------------------------------------------------------------------------------ autograd (8f N1, first slice)
# The scalar SSI / HDN / gradient-preservation losses are differentiable w.r.t. the prediction, so they can sit in the
# reference's training loop on top of any autograd student; the target map is treated as detached (it is the no_grad
# teacher output there).  Backward kernels: csrc/losses.cu (bwd_reduce_kernel / bwd_apply_kernel / sobel_bwd_kernel).
def _gout(g, device):
    return g.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()


class _SSIFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, pred, gt, mask):
        out, _ = _ssi(pred.detach(), gt, mask, False)
        ctx.save_for_backward(pred.detach(), gt.detach(), mask)
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        pred, gt, mask = ctx.saved_tensors
        p, t = _f32(pred, "depth_preds"), _f32(gt, "depth_gt")
        rows, L = _rows_L(p)
        m = _mask_u8(mask, p)
        grad = torch.empty_like(p)
        ws = _workspace(p.device, rows, 1)
        go = _gout(g, p.device)
        _lib.check(_lib.load().dad_ssi_loss_bwd(_lib.ptr(p), _lib.ptr(t), _lib.ptr(m), rows, L, _lib.ptr(go), _lib.ptr(grad),
                                                _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "SSILoss.backward")
        return grad.reshape(pred.shape).to(pred.dtype), None, None


class _HDNFn(torch.autograd.Function):
    """mode 'dr': contexts from (level, gt, mask) on the fly; mode 'ctx': explicit bool [K,B,1,H,W] contexts."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, pred, gt, mode, level, mask_or_ctx):
        ctx.mode, ctx.level = mode, level
        if mode == "dr":
            out = hdn_loss_dr(pred.detach(), gt, mask_or_ctx, level)
        else:
            out, _ = _hdn(pred.detach(), gt, mask_or_ctx)
        ctx.save_for_backward(pred.detach(), gt.detach(), mask_or_ctx if mask_or_ctx is not None else torch.empty(0))
        ctx.has_aux = mask_or_ctx is not None
        return out

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        pred, gt, aux = ctx.saved_tensors
        aux = aux if ctx.has_aux else None
        p, t = _f32(pred, "depth_preds"), _f32(gt, "depth_gt")
        B, L = p.shape[0], p.shape[2] * p.shape[3]
        grad = torch.empty_like(p)
        go = _gout(g, p.device)
        lib = _lib.load()
        if ctx.mode == "dr":
            m = _mask_u8(aux, t)
            ws = _workspace(p.device, B, 2 ** ctx.level - 1)
            _lib.check(lib.dad_hdn_loss_dr_bwd(ctx.level, _lib.ptr(p), _lib.ptr(t), _lib.ptr(m), B, L, _lib.ptr(go),
                                               _lib.ptr(grad), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "hdn backward")
        else:
            K = aux.shape[0]
            c8 = _mask_u8(aux, aux)
            ws = _workspace(p.device, B, K)
            _lib.check(lib.dad_hdn_loss_bwd(_lib.ptr(p), _lib.ptr(t), _lib.ptr(c8), K, B, L, _lib.ptr(go), _lib.ptr(grad),
                                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "hdn backward")
        return grad.reshape(pred.shape).to(pred.dtype), None, None, None, None


class _GradFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, depth):
        ctx.save_for_backward(depth.detach())
        return _grad(depth.detach())[0]

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        (depth,) = ctx.saved_tensors
        d = _f32(depth, "depth")
        grad = torch.empty_like(d)
        go = _gout(g, d.device)
        _lib.check(_lib.load().dad_grad_loss_bwd(_lib.ptr(d), d.shape[0], d.shape[2], d.shape[3], _lib.ptr(go),
                                                 _lib.ptr(grad), _lib.stream_ptr()), "gradient_preservation_loss.backward")
        return grad.to(depth.dtype)


class _FeatFn(torch.autograd.Function):
    @staticmethod
    @_on_tensor_device
    def forward(ctx, student, teacher):
        ctx.save_for_backward(student.detach(), teacher.detach())
        return _feat(student.detach(), teacher.detach())[0]

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        student, teacher = ctx.saved_tensors
        s, t = _f32(student, "student_features"), _f32(teacher, "teacher_features")
        grad = torch.empty_like(s)
        go = _gout(g, s.device)
        _lib.check(_lib.load().dad_feat_cos_loss_bwd(_lib.ptr(s), _lib.ptr(t), s.shape[0], s.shape[1], s.shape[2], t.shape[2],
                                                     _lib.ptr(go), _lib.ptr(grad), _lib.stream_ptr()),
                   "feature_distillation_loss.backward")
        return grad.to(student.dtype), None


class _DistillFn(torch.autograd.Function):
    """distillation_loss(a, b, strategy): gradients w.r.t. whichever of the two maps require them (the local-global term
    of the training loop feeds two student outputs, tools/train_distillation.py:1524-1529)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, a, b, strategy, num_segments):
        ctx.strategy, ctx.nseg = strategy, num_segments
        ctx.save_for_backward(a.detach(), b.detach())
        return _distill(a.detach(), b.detach(), strategy, num_segments)[0]

    @staticmethod
    @_on_tensor_device
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        fa, fb = _f32(a, "student_depth"), _f32(b, "teacher_depth")
        B, L = fa.shape[0], fa.shape[1] * fa.shape[2] * fa.shape[3]
        go = _gout(g, fa.device)
        lib = _lib.load()
        out = []
        for need, x, y in ((ctx.needs_input_grad[0], fa, fb), (ctx.needs_input_grad[1], fb, fa)):
            if not need:
                out.append(None)
                continue
            grad = torch.empty_like(x)
            ws = _workspace(x.device, B, 1)
            _lib.check(lib.dad_distill_loss_bwd(_lib.ptr(x), _lib.ptr(y), _STRATEGY[ctx.strategy], int(ctx.nseg), B, L,
                                                _lib.ptr(go), _lib.ptr(grad), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       "distillation_loss.backward")
            out.append(grad)
        return out[0], out[1], None, None


def _wants_grad(t):
    return isinstance(t, torch.Tensor) and t.requires_grad and torch.is_grad_enabled()


class SSILoss(nn.Module):
    """Scale-shift-invariant MAE (``:675-684``); ``window_size`` is stored and unused, as upstream."""

    def __init__(self, window_size=11):
        super().__init__()
        self.window_size = window_size

    def forward(self, depth_preds, depth_gt, mask_valid, dense=False):
        if _wants_grad(depth_preds):
            if dense:
                raise NotImplementedError("SSILoss(dense=True) has no backward here; the scalar loss and the HDN loss do")
            return _SSIFn.apply(depth_preds, depth_gt, mask_valid)
        return _ssi(depth_preds, depth_gt, mask_valid, dense)[0]


# ------------------------------------------------------------------------------------------ HDN
@_on_tensor_device
def get_contexts_dr(level, depth_gt, mask_valid):
    """``:544-576`` -> bool ``[2**level - 1, B, 1, H, W]``.  The returned tensor remembers how it was
    made so :func:`compute_hdn_loss` can take the fused path that never reads it."""
    g = _f32(depth_gt, "depth_gt")
    if g.dim() != 4 or g.shape[1] != 1:
        raise ValueError("get_contexts_dr expects depth_gt of shape [B, 1, H, W]")
    B, L = g.shape[0], g.shape[2] * g.shape[3]
    m = _mask_u8(mask_valid, g)
    K = 2 ** level - 1
    out = torch.empty((K,) + tuple(g.shape), dtype=torch.uint8, device=g.device)
    ws = _workspace(g.device, B, K)
    _lib.check(_lib.load().dad_contexts_dr(level, _lib.ptr(g), _lib.ptr(m), B, L, _lib.ptr(out), _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr()), "get_contexts_dr")
    ctx = out.view(torch.bool)
    # tag for the fused path: valid only while gt / mask / the contexts themselves are unmodified (weak reference to gt:
    # a re-allocated tensor at the same address never matches)
    ctx._dad_dr = (level, weakref.ref(depth_gt), depth_gt._version, tuple(depth_gt.shape),
                   None if mask_valid is None else (mask_valid.data_ptr(), mask_valid._version), mask_valid, ctx._version)
    return ctx


def _dr_tag(mask_valid_list, depth_gt):
    """The (level, mask) of contexts made by :func:`get_contexts_dr` from exactly this ``depth_gt``, else None."""
    tag = getattr(mask_valid_list, "_dad_dr", None)
    if tag is None or tag[1]() is not depth_gt or tag[2] != depth_gt._version or tag[3] != tuple(depth_gt.shape):
        return None
    if tag[4] is not None and tag[4] != (tag[5].data_ptr(), tag[5]._version):
        return None
    if tag[6] != mask_valid_list._version:   # the contexts were edited in place: use them as given
        return None
    return tag[0], tag[5]


@_on_tensor_device
def _hdn(depth_preds, depth_gt, mask_valid_list, want_partials=False):
    p, g = _f32(depth_preds, "depth_preds"), _f32(depth_gt, "depth_gt")
    if p.dim() != 4 or p.shape[1] != 1:
        raise ValueError("compute_hdn_loss expects maps of shape [B, 1, H, W]")
    B, L = p.shape[0], p.shape[2] * p.shape[3]
    lib = _lib.load()
    out = _new_scalar(p.device)
    part = _new_partials(p.device, want_partials)
    tag = _dr_tag(mask_valid_list, depth_gt)
    if tag is not None:
        level, m = tag[0], _mask_u8(tag[1], g)
        ws = _workspace(p.device, B, 2 ** level - 1)
        _lib.check(lib.dad_hdn_loss_dr(level, _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), B, L, _lib.ptr(out),
                                       _lib.ptr(part), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "compute_hdn_loss")
        return out, part
    ctx = mask_valid_list
    K = ctx.shape[0]
    if tuple(ctx.shape[1:]) != tuple(p.shape):
        raise ValueError("mask_valid_list must be [K, B, 1, H, W]")
    c8 = _mask_u8(ctx, ctx)
    ws = _workspace(p.device, B, K)
    _lib.check(lib.dad_hdn_loss(_lib.ptr(p), _lib.ptr(g), _lib.ptr(c8), K, B, L, _lib.ptr(out), _lib.ptr(part),
                                _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "compute_hdn_loss")
    return out, part


def compute_hdn_loss(ssi_loss, depth_preds, depth_gt, mask_valid_list):
    """``:686-707``.  ``ssi_loss`` is accepted for signature parity (the kernel *is* SSI-MAE)."""
    if _wants_grad(depth_preds):
        tag = _dr_tag(mask_valid_list, depth_gt)
        if tag is not None:
            return _HDNFn.apply(depth_preds, depth_gt, "dr", tag[0], tag[1])
        return _HDNFn.apply(depth_preds, depth_gt, "ctx", 0, mask_valid_list)
    return _hdn(depth_preds, depth_gt, mask_valid_list)[0]


@_on_tensor_device
def hdn_loss_dr(depth_preds, depth_gt, mask_valid=None, level=3, want_partials=False):
    """Fused ``compute_hdn_loss(SSILoss(), p, g, get_contexts_dr(level, g, mask))`` (training call
    site ``:1547-1553``) without materialising the contexts."""
    if _wants_grad(depth_preds) and not want_partials:
        return _HDNFn.apply(depth_preds, depth_gt, "dr", level, mask_valid)
    p, g = _f32(depth_preds, "depth_preds"), _f32(depth_gt, "depth_gt")
    B, L = p.shape[0], p.shape[2] * p.shape[3]
    m = _mask_u8(mask_valid, g)
    out = _new_scalar(p.device)
    part = _new_partials(p.device, want_partials)
    ws = _workspace(p.device, B, 2 ** level - 1)
    _lib.check(_lib.load().dad_hdn_loss_dr(level, _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), B, L, _lib.ptr(out),
                                           _lib.ptr(part), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "hdn_loss_dr")
    return (out, part) if want_partials else out


@_on_tensor_device
def ssi_hdn_dr(depth_preds, depth_gt, mask_valid=None, level=3, want_partials=False):
    """``(SSILoss()(p, g, mask), compute_hdn_loss(SSILoss(), p, g, get_contexts_dr(level, g, mask)))`` from ONE shared
    sweep over the maps (``:449-542``, ``:544-576``, ``:686-707``): the step that reports both losses reads the maps five
    times instead of ten.  ``mask_valid=None`` = all pixels valid (for SSILoss that equals an all-ones mask).
    ``level`` 1..3.  With ``want_partials`` also returns the two float64 ``(numerator, denominator)`` pairs for the
    multi-GPU all-reduce.  Forward only: when a gradient is wanted the two autograd-aware losses are evaluated instead."""
    if _wants_grad(depth_preds) and not want_partials:
        full = torch.ones_like(depth_gt, dtype=torch.bool) if mask_valid is None else mask_valid
        return SSILoss()(depth_preds, depth_gt, full), hdn_loss_dr(depth_preds, depth_gt, mask_valid, level)
    p, g = _f32(depth_preds, "depth_preds"), _f32(depth_gt, "depth_gt")
    if p.dim() != 4 or p.shape[1] != 1 or g.shape != p.shape:
        raise ValueError("ssi_hdn_dr expects maps of shape [B, 1, H, W]")
    if not 1 <= int(level) <= 3:
        raise NotImplementedError("ssi_hdn_dr: the fused path covers HDN levels 1..3")
    B, L = p.shape[0], p.shape[2] * p.shape[3]
    m = _mask_u8(mask_valid, p)
    ssi, hdn = _new_scalar(p.device), _new_scalar(p.device)
    ps, ph = _new_partials(p.device, want_partials), _new_partials(p.device, want_partials)
    ws = _workspace(p.device, B, 8)
    _lib.check(_lib.load().dad_ssi_hdn_dr_loss(int(level), _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), B, L, _lib.ptr(ssi),
                                               _lib.ptr(hdn), _lib.ptr(ps), _lib.ptr(ph), _lib.ptr(ws), ws.numel(),
                                               _lib.stream_ptr()), "ssi_hdn_dr")
    return (ssi, hdn, ps, ph) if want_partials else (ssi, hdn)


@_on_tensor_device
def get_contexts_dp(level, depth_gt, mask_valid):
    """``:578-644`` -> bool ``[2**level - 1, B, 1, H, W]``: depth-percentile bins between the
    ``nanquantile`` values of the valid pixels (exact order statistics by radix select + ATen's lerp)."""
    g = _f32(depth_gt, "depth_gt")
    if g.dim() != 4 or g.shape[1] != 1:
        raise ValueError("get_contexts_dp expects depth_gt of shape [B, 1, H, W]")
    if mask_valid is None:  # the reference indexes with ~mask_valid (:590): None is a TypeError upstream
        raise TypeError("get_contexts_dp: mask_valid must be a bool tensor")
    B, L = g.shape[0], g.shape[2] * g.shape[3]
    m = _mask_u8(mask_valid, g)
    K = 2 ** level - 1
    out = torch.empty((K,) + tuple(g.shape), dtype=torch.uint8, device=g.device)
    ws = _workspace(g.device, B, 2 ** level + 2)
    _lib.check(_lib.load().dad_contexts_dp(level, _lib.ptr(g), _lib.ptr(m), B, L, _lib.ptr(out), _lib.ptr(ws),
                                           ws.numel(), _lib.stream_ptr()), "get_contexts_dp")
    return out.view(torch.bool)


@_on_tensor_device
def get_contexts_ds(level, mask_valid):
    """``:646-673`` -> bool ``[1 + 4 + ... + 4**(level-1), B, 1, H, W]``: valid mask AND an n x n spatial
    grid per level (template side = ``mask_valid.shape[-1]``; square maps, as upstream)."""
    if not isinstance(mask_valid, torch.Tensor) or not mask_valid.is_cuda:
        raise RuntimeError("mask_valid must be a CUDA tensor: the B200 loss path has no CPU fallback")
    if mask_valid.dim() != 4 or mask_valid.shape[1] != 1:
        raise ValueError("get_contexts_ds expects mask_valid of shape [B, 1, H, W]")
    B, _, H, W = mask_valid.shape
    if H != W:
        raise RuntimeError(f"get_contexts_ds: template masks are {W}x{W} but the map is {H}x{W} "
                           "(the reference's broadcast fails the same way)")
    m = _mask_u8(mask_valid, mask_valid)
    K = sum(4 ** i for i in range(level))
    out = torch.empty((K, B, 1, H, W), dtype=torch.uint8, device=mask_valid.device)
    _lib.check(_lib.load().dad_contexts_ds(level, _lib.ptr(m), B, H, W, _lib.ptr(out), _lib.stream_ptr()),
               "get_contexts_ds")
    return out.view(torch.bool)


# ------------------------------------------------------------------------------------------ Sobel
@_on_tensor_device
def _grad(depth, want_partials=False):
    d = _f32(depth, "depth")
    if d.dim() != 4 or d.shape[1] != 1:
        raise ValueError("gradient_preservation_loss expects [B, 1, H, W]")
    out, part = _new_scalar(d.device), _new_partials(d.device, want_partials)
    ws = _workspace(d.device, 1, 1)
    _lib.check(_lib.load().dad_grad_loss(_lib.ptr(d), d.shape[0], d.shape[2], d.shape[3], _lib.ptr(out),
                                         _lib.ptr(part), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
               "gradient_preservation_loss")
    return out, part


def gradient_preservation_loss(depth):
    """``:430-446``."""
    if _wants_grad(depth):
        return _GradFn.apply(depth)
    return _grad(depth)[0]


# ------------------------------------------------------------------------------------------ feature cosine
@_on_tensor_device
def _feat(student_features, teacher_features, want_partials=False):
    s, t = _f32(student_features, "student_features"), _f32(teacher_features, "teacher_features")
    if s.dim() != 3 or t.dim() != 3 or s.shape[0] != t.shape[0]:
        raise NotImplementedError("feature_distillation_loss: only [B,N,Ds] vs [B,N,Dt] tensors are on the hot path")
    if s.shape[1] != t.shape[1]:
        raise NotImplementedError("feature_distillation_loss: token counts differ; the reference would draw fresh "
                                  "random projections every call (tools/train_distillation.py:363-377)")
    out, part = _new_scalar(s.device), _new_partials(s.device, want_partials)
    ws = _workspace(s.device, 1, 1)
    _lib.check(_lib.load().dad_feat_cos_loss(_lib.ptr(s), _lib.ptr(t), s.shape[0], s.shape[1], s.shape[2], t.shape[2],
                                             _lib.ptr(out), _lib.ptr(part), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr()), "feature_distillation_loss")
    return out, part


def feature_distillation_loss(student_features, teacher_features, device=None):
    """``:284-428``: tensor branch, or lists averaged over non-None pairs (``:415-428``)."""
    if isinstance(student_features, (list, tuple)) or isinstance(teacher_features, (list, tuple)):
        tot, n = None, 0
        for s, t in zip(student_features, teacher_features):
            if s is None or t is None:
                continue
            v = _FeatFn.apply(s, t) if _wants_grad(s) else _feat(s, t)[0]
            tot = v if tot is None else tot + v
            n += 1
        if tot is None:
            return torch.tensor(0.0, device=device)
        return tot / max(n, 1)
    if _wants_grad(student_features):
        if student_features.dim() != 3 or teacher_features.dim() != 3 or student_features.shape[1] != teacher_features.shape[1]:
            raise NotImplementedError("feature_distillation_loss: only [B,N,Ds] vs [B,N,Dt] tensors are on the hot path")
        return _FeatFn.apply(student_features, teacher_features)
    return _feat(student_features, teacher_features)[0]


# ------------------------------------------------------------------------------------------ normalised L1
@_on_tensor_device
def _distill(student_depth, teacher_depth, norm_strategy, num_segments=4, want_partials=False, want_norm=False):
    if norm_strategy not in _STRATEGY:
        raise ValueError(f"Unknown normalization strategy: {norm_strategy}")
    s, t = _f32(student_depth, "student_depth"), _f32(teacher_depth, "teacher_depth")
    if s.shape != t.shape or s.dim() != 4:
        raise ValueError("distillation_loss expects two [B, C, H, W] maps of equal shape")
    B, L = s.shape[0], s.shape[1] * s.shape[2] * s.shape[3]
    out, part = _new_scalar(s.device), _new_partials(s.device, want_partials)
    ns = torch.empty_like(s) if want_norm else None
    nt = torch.empty_like(t) if want_norm else None
    ws = _workspace(s.device, B, 1)
    _lib.check(_lib.load().dad_distill_loss(_lib.ptr(s), _lib.ptr(t), _STRATEGY[norm_strategy], int(num_segments), B, L,
                                            _lib.ptr(out), _lib.ptr(part), _lib.ptr(ns), _lib.ptr(nt), _lib.ptr(ws),
                                            ws.numel(), _lib.stream_ptr()), "distillation_loss")
    return out, part, ns, nt


def distillation_loss(student_depth, teacher_depth, norm_strategy, num_segments=4):
    """``:271-282``."""
    if _wants_grad(student_depth) or _wants_grad(teacher_depth):
        if norm_strategy not in _STRATEGY:
            raise ValueError(f"Unknown normalization strategy: {norm_strategy}")
        return _DistillFn.apply(student_depth, teacher_depth, norm_strategy, num_segments)
    return _distill(student_depth, teacher_depth, norm_strategy, num_segments)[0]


def global_normalize(depth):
    """``:173-181``."""
    return _distill(depth, depth, "global", want_norm=True)[2]


def hybrid_normalize(depth, num_segments):
    """``:217-249``."""
    return _distill(depth, depth, "hybrid", num_segments, want_norm=True)[2]


def local_normalize(depth, num_segments):
    """``:251-254``."""
    return hybrid_normalize(depth, num_segments)


def normalize_depth(depth, strategy, num_segments=4):
    """``:256-267``."""
    if strategy == "global":
        return global_normalize(depth)
    if strategy in ("hybrid", "local"):
        return hybrid_normalize(depth, num_segments)
    if strategy == "none":
        return depth
    raise ValueError(f"Unknown normalization strategy: {strategy}")
