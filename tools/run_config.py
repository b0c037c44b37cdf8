#!/usr/bin/env python3
"""keygen + prove + verify at an arbitrary (N, Q, T, B, k): the other BASELINE.json configurations.

    python tools/run_config.py --n 4096 --q 2305843009213693951 --t 65537 --k 16 [--proofs 3]
    python tools/run_config.py --n 4096 --k 16 --rns-bits 109 --limbs 2               # config 3: 109-bit Q as two limb circuits
    python tools/run_config.py --n 16384 --k 19 --rns-bits 438 --limbs 8 [--limb i]   # config 5: 438-bit Q, eight limb circuits
    torchrun --nproc-per-node 8 tools/run_config.py --ring-degree 16384 --k 19 --rns-bits 438 --limbs 8
        (under torchrun with 8 ranks every rank proves limb RANK on its own GPU: limbs are independent proofs)

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

TAU = None          # set below: the trapdoor of the reference's own fallback SRS (zkfhe_reference_test_tau)
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
R_INV = pow(1 << 256, -1, R_MOD)


def run_rns(args):
    """One limb circuit per prime: keygen, prove (x --proofs), verify for each limb this process owns."""
    import json
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover, rns
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = int(os.environ.get("LOCAL_RANK", 0))
    primes = rns.limb_primes(args.rns_bits, args.limbs, args.n)
    par = rns.RnsParams(N=args.n, primes=tuple(primes), T=args.t, B=args.b)
    mine = [args.limb] if args.limb >= 0 else [i for i in range(args.limbs) if i % world == rank]
    ctx = zk_fhe_b200.Context(dev)
    ctx.srs_setup(args.k, TAU)
    enc = rns.sample_encryption(par, np.random.default_rng(args.n))          # same on every rank
    out = {"N": args.n, "k": args.k, "Q_bits": par.Q.bit_length(), "limb_bits": [q.bit_length() for q in primes], "T": args.t,
           "rank": rank, "world": world, "limbs": []}
    if rank == 0:
        print(f"RNS: Q = product of {args.limbs} primes = {par.Q.bit_length()} bits, limbs {out['limb_bits']} bits, N = {args.n}, k = {args.k}", flush=True)
    for i in mine:
        p = par.limb(i)
        t0 = time.perf_counter()
        zeros = {key: ["0"] * (args.n + 1 if key == "cyclo" else args.n) for key in bfv.INPUT_KEYS}
        kg = bfv.BfvCircuit(ctx, p, record=True)
        kg.phase0(zeros).phase1(3)
        pk = prover.keygen(kg.wit, args.k, args.unusable_rows)
        del kg
        ctx.sync()
        t_keygen = time.perf_counter() - t0
        inp = rns.limb_input(ctx, par, enc, i)
        circ = bfv.BfvCircuit(ctx, p)
        pr = prover.Prover(pk, bytes(32))
        vkb, s_g2 = pk.vk_bytes(), ctx.srs_g2(TAU)
        times = []
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
            times.append(1e3 * (time.perf_counter() - t0))
        raw = circ.wit.download(4)
        inst = [int.from_bytes(row.tobytes(), "little") * R_INV % R_MOD for row in raw]
        t0 = time.perf_counter()
        ok = prover.verify(ctx, vkb, inst, proof, s_g2)
        t_verify = 1e3 * (time.perf_counter() - t0)
        names = {0: "accumulate", 1: "ntt", 2: "sort", 3: "fold", 4: "final"}
        cats = {names[c]: round(ctx.timing(c)[0], 2) for c in names}
        rec = {"limb": i, "q_bits": p.Q.bit_length(), "keygen_s": round(t_keygen, 2), "prove_ms": [round(x, 1) for x in times],
               "verify_ms": round(t_verify, 1), "verified": bool(ok), "proof_bytes": len(proof), "advice_columns": pk.info["n_advice"],
               "kernels_ms_last_proof": cats}
        out["limbs"].append(rec)
        print(f"rank {rank} limb {i} ({p.Q.bit_length()}-bit prime): keygen {t_keygen:.2f} s, prove {rec['prove_ms']} ms, verify {t_verify:.1f} ms -> {ok}, "
              f"kernels(ms) {cats}", flush=True)
        assert ok
        del pr, circ, pk
    if args.json:
        path = args.json if world == 1 else args.json.replace(".json", f".rank{rank}.json")
        json.dump(out, open(path, "w"), indent=1)
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--ring-degree", dest="n", type=int, default=4096, help="N (under torchrun write --ring-degree: torchrun reads a bare --n as one of its own options)")
    ap.add_argument("--q", type=int, default=(1 << 61) - 1)
    ap.add_argument("--t", type=int, default=65537)
    ap.add_argument("--b", type=int, default=19)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--unusable-rows", type=int, default=109)
    ap.add_argument("--proofs", type=int, default=3)
    ap.add_argument("--rns-bits", type=int, default=0, help="RNS mode: total bits of Q = product of --limbs primes (zk-fhe_b200/rns.py)")
    ap.add_argument("--limbs", type=int, default=0)
    ap.add_argument("--limb", type=int, default=-1, help="RNS mode: prove only this limb (default: all, or limb RANK under torchrun)")
    ap.add_argument("--json", default="", help="write a machine-readable summary here")
    args = ap.parse_args()
    global TAU
    import zk_fhe_b200 as _z
    TAU = _z.reference_test_tau()
    if args.rns_bits:
        return run_rns(args)
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
    sys.exit(main())
