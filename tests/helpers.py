"""Shared helpers for the parity tests."""
import torch


def sub(x, n=24):
    """Same sub-sampling as oracle/make_golden.py."""
    h = torch.linspace(0, x.shape[-2] - 1, min(n, x.shape[-2])).round().long()
    w = torch.linspace(0, x.shape[-1] - 1, min(n, x.shape[-1])).round().long()
    return x[..., h, :][..., w].contiguous()


def rel_depth_err(d, ref, floor_frac=0.1):
    """Per-pixel relative depth error |d-ref| / max(|ref|, floor_frac*max|ref|).
    The floor keeps pixels that the final ReLU clips to ~0 from dividing by zero
    (SURVEY.md F7); north_star tolerances: 2e-2 (bf16 mode), 1e-4 (fp32 mode)."""
    ref = ref.float()
    den = ref.abs().clamp(min=floor_frac * ref.abs().max().item() + 1e-12)
    return ((d.float() - ref).abs() / den)


def rel_scalar(a, b):
    a, b = float(a), float(b)
    return abs(a - b) / max(abs(b), 1e-6)
