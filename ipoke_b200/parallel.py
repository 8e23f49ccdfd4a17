"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed).

Samples are independent: every rank holds a full replica of the weights, the global batch is
split into contiguous per-rank slices, and the only collective on the data path is ONE gather of the frames at the end
(NCCL over NVLink on the box, gloo in the CPU tests).  Noise is drawn once for the GLOBAL batch from the CPU generator
(second_stage_video.py:300) and sliced per rank, so results do not depend on the number of ranks.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous split of n items over `world` ranks; the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_noise(batch, c0, seed):
    g = torch.Generator().manual_seed(int(seed))
    return torch.randn((batch, c0, 8, 8), generator=g)


def sharded_sample(compute, z, cond, x0, length, gather=True, group=None):
    """compute(z_local, cond_local, x0_local, length) -> frames_local [b,T,3,S,S] on the local device.
    z / cond / x0 are GLOBAL-batch tensors (host or device); every rank slices its own part.
    Returns the gathered [B,T,3,S,S] tensor on rank 0 (None elsewhere) or the local slice when gather=False."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = z.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    local = compute(z[lo:hi], cond[lo:hi], x0[lo:hi], length)
    if not gather or world == 1:
        return local
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    if all((b - a) == (sizes[0][1] - sizes[0][0]) for a, b in sizes):
        out = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
        dist.gather(local.contiguous(), out, dst=0, group=group)          # the single collective on the data path
        return torch.cat(out, dim=0) if rank == 0 else None
    # ragged split: pad to the largest shard, gather, trim
    mx = max(b - a for a, b in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)
