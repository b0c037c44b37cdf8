"""Multi-GPU plumbing (one process per GPU; torch.distributed carries the rendezvous: NCCL on GPUs, gloo in CPU tests).

The prove path shards in two ways; neither needs a data-path reduction:
  * proofs are independent units -> every rank proves its own witnesses (bench.py's throughput number); the only
    collectives are the barrier and the max-over-ranks of the device time;
  * ONE proof over several GPUs -> the library itself shards the commitment phases by column and the quotient by
    coset behind zkfhe_prove_* (csrc/comm.cu, csrc/prover.cu) once a communicator is bound to the context.  This
    module only does the rendezvous: rank 0 makes the NCCL unique id, torch.distributed broadcasts its 128 bytes,
    every rank calls Context.comm_init.
"""
import torch
import torch.distributed as dist

from . import capi


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` that shard `rank` of `world` owns -- the library's own rule
    (zkfhe_shard_range: ceil(n_items / world) items per shard, the last shards may be short or empty)."""
    return capi.shard_range(n_items, world, rank)


def max_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(payload, n_bytes, src=0, device="cpu"):
    """`payload` (bytes, rank `src` only) to every rank of the default process group."""
    buf = torch.zeros(n_bytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        assert len(payload) == n_bytes
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def make_unique_id(device="cpu", factory=None):
    """The 128-byte NCCL unique id of a new communicator, created on rank 0 and broadcast."""
    make = factory or capi.comm_unique_id
    uid = make() if dist.get_rank() == 0 else None
    return broadcast_bytes(uid, 128, 0, device)


def bind_sharded_prover(ctx, device="cpu"):
    """Collective over the default process group: give `ctx` an NCCL communicator spanning all ranks, so that every
    zkfhe_prove_* call on it is one shard of a single proof.  Every rank must then make the same calls on the same
    input and seed; every rank gets the same proof bytes."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return
    ctx.comm_init(dist.get_rank(), world, make_unique_id(device))
