#!/usr/bin/env python3
"""keygen + prove + verify at an arbitrary (N, Q, T, B, k): the other BASELINE.json configurations.

    python tools/run_config.py --n 4096 --q 2305843009213693951 --t 65537 --k 16 [--proofs 3]
    python tools/run_config.py --n 16384 --q 36028797018963913 --t 65537 --k 19      # one RNS limb of config 5

Prints the column shape keygen chose, the proving time of each proof (host strings -> proof bytes),
the verification time of the product verifier and the per-category kernel times.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TAU = 0x5EED5EED5EED5EED5EED5EED
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R_INV = pow(1 << 256, -1, R_MOD)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--q", type=int, default=(1 << 61) - 1)
    ap.add_argument("--t", type=int, default=65537)
    ap.add_argument("--b", type=int, default=19)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--unusable-rows", type=int, default=109)
    ap.add_argument("--proofs", type=int, default=3)
    args = ap.parse_args()
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, bfv_py, prover

    ctx = zk_fhe_b200.Context(0)
    t0 = time.perf_counter()
    ctx.srs_setup(args.k, TAU)
    ctx.sync()
    params = bfv.BfvParams(N=args.n, Q=args.q, T=args.t, B=args.b)
    zeros = {key: ["0"] * (args.n + 1 if key == "cyclo" else args.n) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(3)
    pk = prover.keygen(kg.wit, args.k, args.unusable_rows)
    del kg
    ctx.sync()
    info = pk.info
    print(f"setup + keygen {time.perf_counter() - t0:.2f} s: k={info['k']} advice columns gate {info['n_gate0']}+{info['n_gate1']} "
          f"rlc {info['n_rlc']} lookup {info['n_lookup']}, {info['n_chunks']} permutation products, {info['instances']} instances", flush=True)
    inp = bfv_py.keygen_and_encrypt(ctx, params, np.random.default_rng(args.n))
    vkb, s_g2 = pk.vk_bytes(), ctx.srs_g2(TAU)
    circ = bfv.BfvCircuit(ctx, params)
    pr = prover.Prover(pk, bytes(32))
    names = {0: "accumulate", 1: "ntt", 2: "sort", 3: "fold", 4: "final"}
    for it in range(args.proofs):
        circ.wit.reset()
        ctx.sync()
        ctx.timing_reset()
        t0 = time.perf_counter()
        circ.phase0(inp)
        pr.reset(it.to_bytes(32, "little"))
        gamma = pr.phase0(circ.wit)
        circ.phase1(gamma)
        proof = pr.finish(circ.wit)
        ctx.sync()
        t1 = time.perf_counter()
        raw = circ.wit.download(4)                     # (instances, 4) uint64, Montgomery
        inst = [int.from_bytes(row.tobytes(), "little") * R_INV % R_MOD for row in raw]
        t2 = time.perf_counter()
        ok = prover.verify(ctx, vkb, inst, proof, s_g2)
        t3 = time.perf_counter()
        cats = {names[c]: round(ctx.timing(c)[0], 2) for c in names}
        print(f"proof {it}: prove {1e3 * (t1 - t0):.1f} ms ({len(proof)} bytes)  verify {1e3 * (t3 - t2):.1f} ms -> {ok}  kernels(ms) {cats}",
              flush=True)
        assert ok
    ctx.close()


if __name__ == "__main__":
    main()
