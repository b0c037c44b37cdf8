#!/usr/bin/env python3
"""One commit phase sharded by columns over N GPUs (BASELINE.json config 4).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/bench_sharded_commit.py --k 16 --cols 116 [--iters 10]

Every rank holds the same `cols` full-size columns (the grand-product round of a proof: scalars are
replicated because every rank ran the same witness kernels), commits its contiguous block with the
batched MSM and all-gathers the 64-byte points over NCCL.  Time = CUDA events around the whole phase,
max over ranks.  Rank 0 also commits all columns alone and checks that the gathered points are
byte-identical to that (the sharding is exact: points are not reduced across GPUs, only gathered)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--cols", type=int, default=116)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(v, d)) for v, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)                       # NCCL's banner goes to stdout: keep stdout for the JSON line
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    import zk_fhe_b200
    from zk_fhe_b200 import dist as zd

    ctx = zk_fhe_b200.Context(local_rank)
    ctx.srs_setup(args.k, 0x5EED5EED5EED)
    n = 1 << args.k
    rng = np.random.default_rng(1)          # same seed on every rank: replicated columns
    a = rng.integers(0, 1 << 63, size=(args.cols * n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    d = torch.from_numpy(a.view(np.int64).reshape(-1)).to(dev)
    ctx.fr_convert_dev(d.data_ptr(), args.cols * n, True)
    ctx.sync()
    got = zd.commit_columns_sharded(ctx, d, args.cols)          # warm-up: workspaces, NCCL channels
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.iters):
        got = zd.commit_columns_sharded(ctx, d, args.cols)
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = zd.max_over_ranks(e0.elapsed_time(e1) / args.iters, device=dev)
    identical = None
    if rank == 0:
        alone = torch.empty((args.cols, 64), dtype=torch.uint8, device=dev)
        ctx.msm_g1_dev(d.data_ptr(), args.cols, 1, alone.data_ptr())
        ctx.sync()
        identical = bool(torch.equal(alone, got))
        print(json.dumps({"what": "column-sharded commit phase", "k": args.k, "columns": args.cols, "n_gpus": world,
                          "ms_per_phase": ms, "columns_per_s": 1e3 * args.cols / ms,
                          "identical_to_single_gpu": identical}))
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0 if identical in (None, True) else 1


if __name__ == "__main__":
    sys.exit(main())
