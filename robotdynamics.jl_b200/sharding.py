"""Multi-GPU partitioning of the knot-point batch: one process per GPU, no data-path collective.

Every knot-point evaluation depends only on its own (x, u, t, dt) (reference: the caller's `for k in 1:N` loop,
src/discretized_dynamics.jl:129-136), so ranks own disjoint contiguous knot ranges (or contiguous trajectory blocks,
so rollouts and per-trajectory consumers stay on one GPU) and never exchange data while computing.  torch.distributed
(NCCL on the B200 box, gloo in CPU tests) is used only for the timing barrier, max-over-ranks reductions and the
OPTIONAL all-gather of Jacobian shards for consumers that want the whole trajectory on every GPU.
"""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def partition(N, world_size, rank, align=1):
    """Contiguous [lo, hi) of `N` units for `rank`; block boundaries are multiples of `align` (e.g. the kernel tile or a
    trajectory length) so no tile straddles two ranks; the remainder goes to the leading ranks one block at a time."""
    if world_size < 1 or not (0 <= rank < world_size) or N < 0 or align < 1:
        raise ValueError("bad partition arguments")
    blocks = (N + align - 1) // align
    base, rem = divmod(blocks, world_size)
    lo_b = rank * base + min(rank, rem)
    hi_b = lo_b + base + (1 if rank < rem else 0)
    return min(lo_b * align, N), min(hi_b * align, N)


def partition_segments(segments, world_size, rank, align=1):
    """Mixed sweeps (BASELINE config 5): `segments` = {name: units}; every rank takes its contiguous share of EVERY
    segment, so ranks are balanced by construction whatever the per-unit cost of each model."""
    return {name: partition(n, world_size, rank, align) for name, n in segments.items()}


def barrier_max_ms(local_ms, device=None):
    """max over ranks of a device-timed duration (the only number a multi-GPU measurement may report)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_ms)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_gather_shards(local, counts, dim=0):
    """Optional epilogue: concatenate the per-rank shards (sizes `counts` along `dim`) on every rank.  Uneven shards are
    padded to the largest one for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    big = max(counts)
    pad_shape = list(local.shape)
    pad_shape[dim] = big
    buf = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    parts = [torch.empty_like(buf) for _ in counts]
    dist.all_gather(parts, buf)
    return torch.cat([p.narrow(dim, 0, c) for p, c in zip(parts, counts)], dim=dim)
