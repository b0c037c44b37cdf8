"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the rendezvous helpers the sharded prover uses under
NCCL (zk-fhe_b200/dist.py), and the library's own partition rule (zkfhe_shard_range) is checked exhaustively."""
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

    # the id rendezvous (rank 0 creates, everybody receives the same 128 bytes); the factory stands in for
    # ncclGetUniqueId, which needs no GPU but would tie the CPU suite to libnccl
    uid = zd.make_unique_id(factory=lambda: bytes((7 * i + 1) % 256 for i in range(128)))
    # a sharded commit phase as the library lays it out: ceil(n / world) slots per rank, gathered in rank order
    lo, hi = zd.shard_range(n_cols, rank, world)
    per = -(-n_cols // world)
    mine = torch.zeros((per, 64), dtype=torch.uint8)
    for c in range(lo, hi):
        mine[c - lo] = (c * 7 + 3) % 251
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    full = torch.cat(out)[:n_cols]
    want = torch.stack([torch.full((64,), (c * 7 + 3) % 251, dtype=torch.uint8) for c in range(n_cols)])
    slowest = zd.max_over_ranks(10.0 + rank)
    torch.save({"ok": torch.equal(full, want), "range": (lo, hi), "max": slowest, "uid": uid}, os.path.join(result_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cols", [197, 3, 1])
def test_rendezvous_and_column_layout_world2(tmp_path, n_cols):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_cols, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert all(r["ok"] for r in res)
    assert res[0]["range"][0] == 0 and res[0]["range"][1] == res[1]["range"][0] and res[1]["range"][1] == n_cols
    assert all(r["max"] == 11.0 for r in res)            # max over ranks of the per-rank time
    assert res[0]["uid"] == res[1]["uid"] == bytes((7 * i + 1) % 256 for i in range(128))


def test_shard_range_partitions_exactly():
    """zkfhe_shard_range (the C ABI's rule, used for the columns of a commit phase and for the expression list of a
    coset): contiguous, in order, covers [0, n), ceil(n / world) per shard."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from zk_fhe_b200.dist import shard_range
    for n in (0, 1, 3, 7, 100, 197, 411):
        for world in (1, 2, 3, 4, 5, 8, 64):
            blocks = [shard_range(n, r, world) for r in range(world)]
            per = -(-n // world)
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert all(hi - lo <= per for lo, hi in blocks)
            assert all(lo == min(r * per, n) for r, (lo, hi) in enumerate(blocks))


def test_comm_defaults_without_a_gpu():
    import zk_fhe_b200
    from zk_fhe_b200 import capi
    lib = zk_fhe_b200.load_library()
    import ctypes
    lo, hi = ctypes.c_uint32(), ctypes.c_uint32()
    assert lib.zkfhe_shard_range(10, 0, 0, ctypes.byref(lo), ctypes.byref(hi)) == capi.ERR_ARG
    assert lib.zkfhe_shard_range(10, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) == capi.ERR_ARG
