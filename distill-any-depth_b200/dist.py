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


def stack_partials(partials):
    """{name: (kind, float64[2])} -> (names, kinds, float64 [n, 2] tensor).  Pure device ops, no collective: this half can
    sit inside a captured CUDA graph while the all-reduce of :func:`finish_stacked` stays outside it (a graph that holds
    NCCL kernels keeps the communicator busy at teardown: ``destroy_process_group`` was observed to hang on it)."""
    names = sorted(partials)
    vec = torch.stack([partials[n][1].to(torch.float64) for n in names])  # [n, 2]
    return names, [partials[n][0] for n in names], vec


def finish_stacked(names, kinds, vec, group=None):
    """ONE all-reduce (sum) of the stacked per-rank (numerator, denominator) partials, in place, then the ratios."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    out = {}
    for i, (n, kind) in enumerate(zip(names, kinds)):
        v = vec[i, 0] / (vec[i, 1] + LOSS_EPS[kind])
        if kind == "feat":
            v = 1.0 - v
        out[n] = v.to(torch.float32)
    return out


def finish_losses(partials, group=None):
    """partials: {name: (kind, float64[2] tensor)} of per-rank (numerator, denominator).
    One all-reduce over the stacked vector; returns {name: fp32 scalar tensor} of full-batch losses."""
    return finish_stacked(*stack_partials(partials), group=group)


# ---------------------------------------------------------------------------------------------- data-parallel training
# The reference trains in a single process, so its gradient is that of the FULL-batch losses.  Every loss is
# numerator / (denominator + eps) with a denominator that does not depend on the parameters (valid-pixel / element counts),
# so the full-batch gradient is sum_r (den_r + eps) / (den_total + eps) * grad(loss_r): scale each local loss by its shard
# weight before backward(), then SUM-all-reduce the gradients - exactly the single-process gradient, with two small
# collectives per step (the denominators, then one flat gradient buffer).
def shard_loss_weights(partials, group=None):
    """partials as for finish_losses (per-rank (numerator, denominator)); -> {name: fp32 scalar} shard weights
    (den_local + eps) / (den_total + eps) (1.0 when not distributed)."""
    names = sorted(partials)
    den = torch.stack([partials[n][1].to(torch.float64)[1] for n in names])
    tot = den.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
    out = {}
    for i, n in enumerate(names):
        eps = LOSS_EPS[partials[n][0]]
        out[n] = ((den[i] + eps) / (tot[i] + eps)).to(torch.float32)
    return out


def allreduce_gradients(parameters, group=None, average=False, bucket_bytes=256 << 20):
    """SUM (or mean) all-reduce of the .grad of `parameters` through flat fp32 buckets (one collective per bucket;
    NVSwitch reduces in the fabric, so buckets are sized for launch latency, not link count).  Parameters whose .grad is
    None on this rank but not on another are treated as zeros; parameters without a gradient on every rank are left at
    None.  Returns the number of collectives (including the has-grad bitmap exchange)."""
    params = [p for p in parameters if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or not params:
        return 0
    world = dist.get_world_size(group)
    # parameters without a gradient on EVERY rank stay at .grad = None, as in single-process training (the optimiser
    # then skips them: no weight decay / moment update on mask_token, refinenet4.resConfUnit1): one tiny MAX all-reduce
    # of a has-grad bitmap decides it (counted in the returned number of collectives)
    has = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], dtype=torch.float32, device=params[0].device)
    dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
    keep = has.cpu().tolist()
    params = [p for p, k in zip(params, keep) if k > 0]
    n_coll, bucket, size = 1, [], 0

    def flush():
        nonlocal n_coll, bucket, size
        if not bucket:
            return
        for p in bucket:
            if p.grad is None:
                p.grad = torch.zeros_like(p, dtype=torch.float32)
        flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        n_coll += 1
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll
