"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding / gather helpers the
GPU path uses under NCCL (zk-fhe_b200/dist.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cols, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from zk_fhe_b200 import dist as zd

    lo, hi = zd.shard_range(n_cols, rank, world)
    # stand-in for the per-column MSM: a deterministic 64-byte "commitment" per global column index
    local = torch.stack([torch.full((64,), (c * 7 + 3) % 251, dtype=torch.uint8) for c in range(lo, hi)]) if hi > lo \
        else torch.zeros((0, 64), dtype=torch.uint8)
    full = zd.gather_commitments(local, n_cols)
    want = torch.stack([torch.full((64,), (c * 7 + 3) % 251, dtype=torch.uint8) for c in range(n_cols)])
    ok = torch.equal(full, want)
    slowest = zd.max_over_ranks(10.0 + rank)
    torch.save({"ok": ok, "range": (lo, hi), "max": slowest}, os.path.join(result_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [197, 3, 1])
def test_column_sharding_and_gather_world2(tmp_path, n_cols):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_cols, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert all(r["ok"] for r in res)
    assert res[0]["range"][0] == 0 and res[0]["range"][1] == res[1]["range"][0] and res[1]["range"][1] == n_cols
    assert all(r["max"] == 11.0 for r in res)            # max over ranks of the per-rank time


def test_shard_range_partitions_exactly():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from zk_fhe_b200.dist import shard_range
    for n in (0, 1, 7, 197, 411):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
