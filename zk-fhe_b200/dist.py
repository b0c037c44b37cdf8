"""Multi-GPU plumbing (one process per GPU, torch.distributed: NCCL on GPUs, gloo in CPU tests).

The prove path shards in two ways, neither needs a data-path reduction:
  * proofs are independent units  -> every rank proves its own witnesses (what bench.py does);
    the only collectives are the barrier and the max-over-ranks of the device time;
  * inside one proof, the columns of a commit phase are independent -> `commit_columns_sharded`
    gives each rank a contiguous block of columns and all-gathers the 64-byte commitments in column
    order (EC points cannot be all-reduced; 64 B x columns is latency-, not bandwidth-bound).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` for `rank`; blocks differ in size by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_commitments(local, n_cols, device="cpu"):
    """local: uint8 tensor [my_cols, 64] (this rank's block).  Returns [n_cols, 64] in column order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank()
    sizes = [shard_range(n_cols, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width, 64), dtype=torch.uint8, device=device)
    lo, hi = sizes[rank]
    pad[:hi - lo] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([out[r][:sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def commit_columns_sharded(ctx, d_scalars, n_cols, basis=1):
    """Column-sharded commit phase: d_scalars is a CUDA uint8/int64 tensor holding n_cols x 2^k Fr
    (replicated on every rank); returns a CUDA uint8 tensor [n_cols, 64] of affine commitments."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_range(n_cols, rank, world)
    col_bytes = 32 << ctx.srs_k
    local = torch.empty((hi - lo, 64), dtype=torch.uint8, device=d_scalars.device)
    if hi > lo:
        ctx.msm_g1_dev(d_scalars.data_ptr() + lo * col_bytes, hi - lo, basis, local.data_ptr())
        ctx.sync()
    return gather_commitments(local, n_cols, device=d_scalars.device)
