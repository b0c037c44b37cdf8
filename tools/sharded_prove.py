#!/usr/bin/env python3
"""ONE proof over N GPUs: launched under torchrun, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/sharded_prove.py [--k 13|16] [--proofs 5] [--transcript 1] [--out result.json]

Every rank loads the same SRS and proving key, proves the same input with the same seed once on its own (the
single-GPU proof) and then, after binding an NCCL communicator to its context, as one shard of the sharded proof.
Checks that every rank's sharded proof equals the single-GPU bytes, and reports the single-proof latency of both
(host wall clock around the whole prove, max over ranks) with the device time spent in the collectives.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=13)
    ap.add_argument("--proofs", type=int, default=5)
    ap.add_argument("--transcript", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from zk_fhe_b200 import sharded
    res = sharded.run(args.k, args.proofs, args.transcript, dist, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            json.dump(res, open(args.out, "w"))
    if world > 1:
        dist.destroy_process_group()
    return 0 if res["identical_to_single_gpu"] else 1


if __name__ == "__main__":
    sys.exit(main())
