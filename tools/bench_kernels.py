#!/usr/bin/env python3
"""Stage (2) / (3) kernels alone, on the column shapes of one config-1 proof (no profiler inside).

    python tools/bench_kernels.py [--k 13] [--iters 5] [--what msm,ntt]

MSM: the 137 full-size columns of the grand-product round and the 194 witness-like columns of the
phase-1 commit, through zkfhe_msm_g1_dev.  NTT: 406 iNTT(2^k) and 406 coeff_to_extended(2^k -> 2^(k+2)).
Times are CUDA events recorded inside the library around every launch (ctx.timing).  This is the
command the `ncu --set full` captures under profiles/ are taken on.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CATS = {0: "accumulate", 1: "ntt", 2: "sort", 3: "fold", 4: "final"}


def witness_like(rng, cols, n):
    a = np.zeros((cols, n, 4), np.uint64)
    small = rng.integers(0, 1 << 29, size=(cols, n), dtype=np.uint64)
    small[rng.random((cols, n)) < 0.6] &= np.uint64(0xFF)
    a[:, :, 0] = small
    return a


def full_size(rng, cols, n):
    a = rng.integers(0, 1 << 63, size=(cols, n, 4), dtype=np.uint64)
    a[:, :, 3] &= np.uint64((1 << 60) - 1)
    return a


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=13)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--what", default="micro,msm,ntt")
    ap.add_argument("--full-cols", type=int, default=137)
    ap.add_argument("--small-cols", type=int, default=194)
    ap.add_argument("--ntt-cols", type=int, default=406)
    args = ap.parse_args()
    import torch

    import zk_fhe_b200

    ctx = zk_fhe_b200.Context(0)
    k, n = args.k, 1 << args.k
    rng = np.random.default_rng(7)
    dev = torch.device("cuda", 0)

    def upload(a):
        return torch.from_numpy(a.view(np.int64).reshape(-1)).to(dev)

    def report(tag, pairs=None, elems=None):
        ctx.sync()
        t = {name: ctx.timing(c) for c, name in CATS.items()}
        line = f"{tag:28s}" + "  ".join(f"{name} {v[0] / args.iters:8.3f} ms" for name, v in t.items() if v[1])
        if pairs:
            acc = t["accumulate"][0] / args.iters
            tot = sum(v[0] for v in t.values()) / args.iters
            line += f"  | accumulate {96 * pairs / acc / 1e6:7.1f} GB/s  whole MSM {96 * pairs / tot / 1e6:7.1f} GB/s"
        if elems:
            ms = t["ntt"][0] / args.iters
            line += f"  | {64 * elems / ms / 1e6:7.1f} GB/s"
        print(line, flush=True)

    if "micro" in args.what:
        ms, ops = ctx.microbench(0, 2000)
        print(f"montgomery products, full GPU   {ops / ms / 1e6:8.2f} G/s  ({ms:.3f} ms)", flush=True)
        for kind, name, it in ((1, "xyzz add", 200), (2, "mixed add", 200), (3, "field product", 2000),
                               (4, "inversion (binary Euclid)", 20), (5, "inversion (Fermat)", 20)):
            ms, ops = ctx.microbench(kind, it)
            print(f"one-warp chain: {name:28s} {1e3 * ms / ops:9.3f} us per op", flush=True)
    if "msm" in args.what:
        ctx.srs_setup(k, 0x5EED5EED)
        out = torch.zeros(64 * 512, dtype=torch.uint8, device=dev)
        for tag, cols, gen, small in (("msm full-size", args.full_cols, full_size, False),
                                      ("msm witness-like", args.small_cols, witness_like, False),
                                      ("msm witness-like +hint", args.small_cols, witness_like, True),
                                      ("msm full-size x3", 3, full_size, False), ("msm full-size x1", 1, full_size, False)):
            d = upload(gen(rng, cols, n))
            ctx.fr_convert_dev(d.data_ptr(), cols * n, True)
            ctx.msm_g1_dev(d.data_ptr(), cols, 1, out.data_ptr(), small_values=small)      # warm-up (workspace allocation)
            ctx.sync()
            ctx.timing_reset()
            for _ in range(args.iters):
                ctx.msm_g1_dev(d.data_ptr(), cols, 1, out.data_ptr(), small_values=small)
            report(f"{tag} [{cols} x 2^{k}]", pairs=cols * n)
    if "ntt" in args.what:
        cols = args.ntt_cols
        d = upload(full_size(rng, cols, n))
        ext = torch.zeros(cols * n * 4 * 4, dtype=torch.int64, device=dev)
        ctx.ntt_fr_dev(d.data_ptr(), k, cols, True, False)
        ctx.coeff_to_extended_dev(d.data_ptr(), k, ext.data_ptr(), k + 2, cols)
        ctx.sync()
        ctx.timing_reset()
        for _ in range(args.iters):
            ctx.ntt_fr_dev(d.data_ptr(), k, cols, True, False)
        report(f"intt [{cols} x 2^{k}]", elems=cols * n)
        ctx.timing_reset()
        for _ in range(args.iters):
            ctx.coeff_to_extended_dev(d.data_ptr(), k, ext.data_ptr(), k + 2, cols)
        report(f"coeff_to_extended [{cols} x 2^{k + 2}]", elems=cols * n * 4)
    ctx.close()


if __name__ == "__main__":
    main()
