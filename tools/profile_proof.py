#!/usr/bin/env python3
"""Per-kernel-category device time of one bfv.in proof (CUDA events inside the library, no profiler).

    python tools/profile_proof.py [--iters 5]

Categories: accumulate (MSM bucket accumulation), ntt, sort / fold / final (MSM counting sort and the
two bucket-reduction levels).  `other` is everything else in the proof's wall time: witness kernels,
grand products, quotient, evaluations, SHPLONK, host transcript and synchronisation.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--input", default=os.path.join(ROOT, "tests", "golden", "bfv.in"))
    args = ap.parse_args()
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover

    ctx = zk_fhe_b200.Context(0)
    ctx.srs_setup(13, 777)
    inp = bfv.load_input(args.input)
    zeros = {key: ["0"] * len(v) for key, v in inp.items()}
    kg = bfv.BfvCircuit(ctx, record=True)
    kg.phase0(zeros).phase1(1)
    pk = prover.keygen(kg.wit, 13, 109)
    del kg
    circ = bfv.BfvCircuit(ctx)
    pr = prover.Prover(pk, bytes(32))
    names = {0: "accumulate", 1: "ntt", 2: "sort", 3: "fold", 4: "final"}
    for it in range(args.iters):
        circ.wit.reset()
        ctx.sync()
        ctx.timing_reset()
        t0 = time.perf_counter()
        circ.phase0(inp)
        t1 = time.perf_counter()
        pr.reset(bytes(32))
        g = pr.phase0(circ.wit)
        t2 = time.perf_counter()
        circ.phase1(g)
        t3 = time.perf_counter()
        proof = pr.finish(circ.wit)
        ctx.sync()
        t4 = time.perf_counter()
        cats = {names[c]: round(ctx.timing(c)[0], 3) for c in names}
        wall = (t4 - t0) * 1e3
        print(f"iter {it}: wall {wall:.2f} ms  [phase0 issue {1e3 * (t1 - t0):.2f} | commit0 {1e3 * (t2 - t1):.2f} | "
              f"phase1 issue {1e3 * (t3 - t2):.2f} | finish {1e3 * (t4 - t3):.2f}]  kernels(ms) {cats}  "
              f"other {wall - sum(cats.values()):.2f}  proof {len(proof)} B")
        print("        rounds(ms)", pr.round_ms())


if __name__ == "__main__":
    main()
