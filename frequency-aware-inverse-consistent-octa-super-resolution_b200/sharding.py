"""Multi-GPU use of the hot path: plain batch sharding, one process per GPU.

Every (n, c) image plane is independent in the DWT / IDWT (depthwise filtering,
``groups=C`` at ``pw/dwt/lowlevel.py:143,164,253``) and in SSIM up to the final
mean, so ranks take contiguous slices of the batch and the data path needs NO
collective.  The only exchange is the scalar SSIM mean (and, in a training step,
DDP's gradient all-reduce, which is outside this package).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of ``n`` items for ``rank`` (earlier ranks get the remainder)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size: %r/%r" % (rank, world_size))
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world_size=None):
    """This rank's slice of a batch-first tensor."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(x.shape[0], rank, world_size)
    return x[lo:hi]


def global_mean(local_mean, local_count, group=None, reducer="mean", total_count=None):
    """Mean over all ranks of per-rank means weighted by their element counts (ragged shards allowed).  One
    all-reduce of 2 floats (1 when ``total_count`` is given, which also avoids the host read-back of the count).

    The returned value is the global mean on every rank; its gradient flows through the local term only, scaled for
    the gradient reducer that follows in the training step, so that the reduced parameter gradients equal those of the
    single-process reference (``ssim_map.mean()`` over the whole batch, ``ssim.py:33-37``):

    * ``reducer="mean"`` (default; torch DDP averages gradients over ranks):
      d value / d local_mean = world_size * local_count / total_count   (1.0 for equal shards)
    * ``reducer="sum"`` (gradients are summed over ranks): d value / d local_mean = local_count / total_count
    """
    if reducer not in ("mean", "sum"):
        raise ValueError("reducer must be 'mean' or 'sum', not %r" % (reducer,))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_mean
    world = dist.get_world_size(group)
    with torch.no_grad():
        if total_count is None:
            buf = torch.stack([local_mean.detach().double() * float(local_count),
                               torch.tensor(float(local_count), dtype=torch.float64, device=local_mean.device)])
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            total = float(buf[1].item())
            total_sum = buf[0]
        else:
            total = float(total_count)
            total_sum = local_mean.detach().double() * float(local_count)
            dist.all_reduce(total_sum, op=dist.ReduceOp.SUM, group=group)
        value = (total_sum / total).to(local_mean.dtype)
    weight = float(local_count) / total * (world if reducer == "mean" else 1.0)
    return value.detach() + (local_mean - local_mean.detach()) * weight


def sharded_ssim(ssim_module, img1, img2, group=None, reducer="mean", total_count=None):
    """``size_average=True`` SSIM over a batch that is sharded across ranks: each rank evaluates its shard with the
    fused kernel, then the scalar means are combined (see ``global_mean`` for the gradient convention)."""
    local = ssim_module(img1, img2)
    return global_mean(local, img1.numel(), group=group, reducer=reducer, total_count=total_count)


def average_gradients(params, group=None):
    """Average the gradients of ``params`` over the ranks with ONE flat all-reduce (what DDP's reducer does for a
    bucket), for training steps that DDP's per-forward bookkeeping does not fit -- the reference's step calls each
    generator three times before one backward and never uses ``NetworkA2B.unet`` (train.py:170-240, model.py:249).
    Parameters without a gradient are skipped (they must be the same on every rank).  No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(f)
