"""Oracle: fp32 CPU restatement of the reference's distillation losses.

All citations are into ``/root/reference/tools/train_distillation.py``.
Written vectorised (no per-image Python loops where the maths allows) but with
the reference's exact operation order wherever rounding decides a comparison
(HDN context thresholds, segment boundaries).

TEST INFRASTRUCTURE: see oracle/__init__.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- SSI
def _masked_lower_median(x, m):
    """Lower median of x over m per row ([R, L]); 0 when a row has no valid
    element.  == nanmedian of the NaN-filled clone followed by NaN->0 (:464-490)."""
    cnt = m.sum(-1)
    srt = torch.sort(x.masked_fill(~m, float("inf")), dim=-1).values
    idx = ((cnt - 1).clamp(min=0) // 2).unsqueeze(-1)
    med = srt.gather(-1, idx).squeeze(-1)
    return torch.where(cnt > 0, med, torch.zeros_like(med)), cnt


def _shift_scale(x, mask):
    """One side of masked_shift_and_scale (:472-498): returns aligned map, t, s."""
    B, C = x.shape[:2]
    xf, mf = x.reshape(B * C, -1), mask.reshape(B * C, -1)
    t, cnt = _masked_lower_median(xf, mf)
    t = t.unsqueeze(-1)
    s = ((xf - t).abs() * mf).sum(-1, keepdim=True) / (cnt.unsqueeze(-1) + 1)  # :470 count+1
    al = (xf - t) / (s + 1e-6)
    return al.reshape(x.shape), t.reshape(B, C, 1, 1), s.reshape(B, C, 1, 1)


def masked_shift_and_scale(depth_preds, depth_gt, mask_valid):
    """:449-533.  Median / mean-absolute-deviation alignment, per (b, c) row."""
    gt_al, _, _ = _shift_scale(depth_gt, mask_valid)
    pr_al, _, _ = _shift_scale(depth_preds, mask_valid)
    return pr_al, gt_al


def masked_l1_loss(preds, target, mask_valid, dense=False):
    """:535-542."""
    e = (preds - target).abs() * mask_valid
    if dense:
        return e
    return e.sum() / (mask_valid.sum() + 1e-6)


class SSILoss(nn.Module):
    """:675-684 (window_size stored, unused)."""

    def __init__(self, window_size=11):
        super().__init__()
        self.window_size = window_size

    def forward(self, depth_preds, depth_gt, mask_valid, dense=False):
        p, g = masked_shift_and_scale(depth_preds, depth_gt, mask_valid)
        return masked_l1_loss(p, g, mask_valid, dense)


# --------------------------------------------------------------------------- HDN
def get_contexts_dr(level, depth_gt, mask_valid):
    """:544-576.  bool [2**level - 1, B, 1, H, W].

    Thresholds follow the reference's fp32 evaluation order
    ``min + ((max - min) * i) * bin`` and ``... + 1e-30``."""
    if mask_valid is None:
        mask_valid = torch.ones_like(depth_gt, dtype=torch.bool)
    bins = [0.5 ** i for i in range(level)][::-1]
    per_image = []
    for b in range(depth_gt.shape[0]):
        d, v = depth_gt[b], mask_valid[b]
        if not bool(v.any()):
            per_image.append(torch.stack([v] * (2 ** level - 1)))
            continue
        vals = d[v]
        mx, mn = vals.max(), vals.min()
        ctx = []
        for bs in bins:
            for i in range(int(1 / bs)):
                lo = mn + (mx - mn) * i * bs
                hi = mn + (mx - mn) * (i + 1) * bs + 1e-30
                ctx.append((d >= lo) & (d < hi) & v)
        per_image.append(torch.stack(ctx))
    return torch.stack(per_image).swapdims(0, 1)


def get_contexts_dp(level, depth_gt, mask_valid):
    """:578-644 (nanquantile bins)."""
    d = depth_gt.clone()
    d[~mask_valid] = float("nan")
    flat = d.view(d.shape[0], d.shape[1], -1)
    bins = [0.5 ** i for i in range(level)][::-1]
    out = []
    for bs in bins:
        for i in range(int(1 / bs)):
            lo = flat.nanquantile(i * bs, dim=-1).unsqueeze(-1).unsqueeze(-1)
            hi = flat.nanquantile((i + 1) * bs, dim=-1).unsqueeze(-1).unsqueeze(-1)
            out.append(mask_valid & (depth_gt >= lo) & (depth_gt < hi))
    return torch.stack(out)


def get_contexts_ds(level, mask_valid):
    """:646-673 (spatial grid; square maps)."""
    size = mask_valid.shape[-1]
    bins = [0.5 ** i for i in range(level)][::-1]
    tm = []
    for bs in bins:
        n = int(1 / bs)
        for h in range(n):
            for w in range(n):
                t = torch.zeros(1, 1, size, size, dtype=torch.bool)
                t[:, :, int(h * bs * size):int((h + 1) * bs * size),
                  int(w * bs * size):int((w + 1) * bs * size)] = True
                tm.append(t)
    tm = torch.stack(tm).to(mask_valid.device)
    return mask_valid.unsqueeze(0) & tm


def compute_hdn_loss(ssi_loss, depth_preds, depth_gt, mask_valid_list):
    """:686-707."""
    K = mask_valid_list.shape[0]
    rep = lambda x: x.unsqueeze(0).expand(K, *x.shape).reshape(-1, *x.shape[-3:])
    dense = ssi_loss(rep(depth_preds), rep(depth_gt),
                     mask_valid_list.reshape(-1, *mask_valid_list.shape[-3:]), dense=True)
    per_ctx = dense.reshape(*mask_valid_list.shape).sum(0)
    times = mask_valid_list.sum(0)
    valid = times != 0
    per_px = torch.where(valid, per_ctx / times.clamp(min=1), per_ctx)
    return per_px.sum() / (valid.sum() + 1e-6)


# --------------------------------------------------------------------------- Sobel
def gradient_preservation_loss(depth):
    """:430-446.  mean(exp(-sqrt(gx^2 + gy^2 + 1e-6))), zero-padded cross-correlation."""
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3).to(depth)
    ky = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]]).view(1, 1, 3, 3).to(depth)
    gx = F.conv2d(depth, kx, padding=1)
    gy = F.conv2d(depth, ky, padding=1)
    return torch.exp(-torch.sqrt(gx ** 2 + gy ** 2 + 1e-6)).mean()


# --------------------------------------------------------------------------- feature cosine
def feature_distillation_loss(student_features, teacher_features, device=None):
    """:284-428, tensor branch with equal token count; list inputs
    average over non-None pairs (:415-428)."""
    if isinstance(student_features, (list, tuple)) or isinstance(teacher_features, (list, tuple)):
        tot, n = 0.0, 0
        for s, t in zip(student_features, teacher_features):
            if s is None or t is None:
                continue
            tot = tot + feature_distillation_loss(s, t, device)
            n += 1
        return tot / max(n, 1)
    s, t = student_features, teacher_features
    if s.dim() != 3 or t.dim() != 3 or s.shape[1] != t.shape[1]:
        raise NotImplementedError("only [B,N,Ds] vs [B,N,Dt] with equal N (reference draws "
                                  "fresh random projections otherwise, :363-377)")
    if s.shape[2] != t.shape[2]:
        small = min(s.shape[2], t.shape[2])
        if s.shape[2] != small:
            s = F.interpolate(s, size=(small,), mode="nearest")
        if t.shape[2] != small:
            t = F.interpolate(t, size=(small,), mode="nearest")
    sn = F.normalize(s, p=2, dim=1)
    tn = F.normalize(t, p=2, dim=1)
    return 1.0 - F.cosine_similarity(sn, tn, dim=1).mean()


# --------------------------------------------------------------------------- normalised L1
def global_normalize(depth):
    """:173-181 (torch.median == lower median over C*H*W)."""
    med = torch.median(depth.reshape(depth.shape[0], -1), dim=1, keepdim=True)[0][..., None, None]
    mad = (depth - med).abs().mean(dim=(1, 2, 3), keepdim=True)
    return (depth - med) / (mad + 1e-6)


def hybrid_normalize(depth, num_segments):
    """:183-249.  Equal depth-range segments, inclusive bounds, later segment wins."""
    b = depth.shape[0]
    flat = depth.reshape(b, -1)
    mn = flat.min(dim=1, keepdim=True)[0]
    mx = flat.max(dim=1, keepdim=True)[0]
    rng = mx - mn
    bounds = [(mn + (i / num_segments) * rng).reshape(b, 1, 1, 1) for i in range(num_segments + 1)]
    out = torch.zeros_like(depth)
    for i in range(num_segments):
        m = (depth >= bounds[i]) & (depth <= bounds[i + 1])
        if not bool(m.any()):
            continue
        mf = m.float()
        seg = torch.where(m, depth, torch.zeros_like(depth))
        cnt = mf.sum(dim=(1, 2, 3), keepdim=True) + 1e-6
        mean = seg.sum(dim=(1, 2, 3), keepdim=True) / cnt
        mad = ((seg - mean).abs() * mf).sum(dim=(1, 2, 3), keepdim=True) / cnt
        out = torch.where(m, (seg - mean) / (mad + 1e-6), out)
    return out


def normalize_depth(depth, strategy, num_segments=4):
    """:256-267."""
    if strategy == "global":
        return global_normalize(depth)
    if strategy in ("hybrid", "local"):
        return hybrid_normalize(depth, num_segments)
    if strategy == "none":
        return depth
    raise ValueError(f"Unknown normalization strategy: {strategy}")


def distillation_loss(student_depth, teacher_depth, norm_strategy, num_segments=4):
    """:271-282."""
    if norm_strategy != "none":
        return F.l1_loss(normalize_depth(student_depth, norm_strategy, num_segments),
                         normalize_depth(teacher_depth, norm_strategy, num_segments))
    return F.l1_loss(student_depth, teacher_depth)
