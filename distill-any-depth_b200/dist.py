"""Data-parallel sharding of the hot path: images shard across ranks, nothing in the
forward crosses GPUs, and each loss is finished by ONE all-reduce (sum) of its (numerator,
denominator) partials so the result equals the single-process full-batch value."""
import torch
import torch.distributed as dist

LOSS_EPS = {"ssi": 1e-6, "hdn": 1e-6, "grad": 0.0, "feat": 0.0, "distill": 0.0}


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n items owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world=None):
    if rank is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def finish_losses(partials, group=None):
    """partials: {name: (kind, float64[2] tensor)} of per-rank (numerator, denominator).
    One all-reduce over the stacked vector; returns {name: fp32 scalar tensor} of full-batch losses."""
    names = sorted(partials)
    vec = torch.stack([partials[n][1].to(torch.float64) for n in names])  # [n, 2]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    out = {}
    for i, n in enumerate(names):
        kind = partials[n][0]
        v = vec[i, 0] / (vec[i, 1] + LOSS_EPS[kind])
        if kind == "feat":
            v = 1.0 - v
        out[n] = v.to(torch.float32)
    return out
