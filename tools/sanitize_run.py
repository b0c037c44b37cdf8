#!/usr/bin/env python3
"""One pass over every kernel family of the prove path, small enough to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Field self test, NTT (forward / inverse / coset / coeff_to_extended), MSM (both tables, cluster and single-CTA sort,
skewed columns and runs of equal scalars), stage (1) + keygen + prove + verify of a small BFV circuit (N = 16, k = 10)
with both transcripts, the same proof as three virtual shards,
and -- with --full -- one config-1 proof (N = 1024, k = 13).  Results are checked by the product verifier, so a run
that sanitises clean has also produced valid proofs.  SURVEY.md section 5 (race detection / sanitizers)."""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAU = 0x1234567890ABCDEF1234567890ABCDEF


def synthetic_input(rng, N, Q, T, B):
    pk0, pk1 = [rng.randrange(Q) for _ in range(N)], [rng.randrange(Q) for _ in range(N)]
    u = [rng.choice([0, 1, Q - 1]) for _ in range(N)]
    e0, e1 = [rng.randrange(-B, B + 1) % Q for _ in range(N)], [rng.randrange(-B, B + 1) % Q for _ in range(N)]
    m = [rng.randrange(-(T // 2), T // 2 + 1) % Q for _ in range(N)]

    def ring_mul(a, b):
        p = [0] * (2 * N - 1)
        for i in range(N):
            for j in range(N):
                p[i + j] += a[i] * b[j]
        return [(p[N - 1 + i] - (p[i - 1] if i else 0)) % Q for i in range(N)]

    c0 = [(x + (Q // T) * mm + e) % Q for x, mm, e in zip(ring_mul(pk0, u), m, e0)]
    c1 = [(x + e) % Q for x, e in zip(ring_mul(pk1, u), e1)]
    d = {"pk0": pk0, "pk1": pk1, "m": m, "u": u, "e0": e0, "e1": e1, "c0": c0, "c1": c1, "cyclo": [1] + [0] * (N - 1) + [1]}
    return {k: [str(x) for x in v] for k, v in d.items()}


def main():
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover
    ctx = zk_fhe_b200.Context(0)
    assert ctx.selftest(1 << 10, seed=5) == 0
    rng = np.random.default_rng(0)
    for k, batch in ((6, 3), (10, 4), (13, 2)):
        data = rng.integers(0, 1 << 60, size=(batch << k, 4), dtype=np.uint64)
        want = data.copy()
        for inverse, coset in ((False, False), (True, False), (False, True), (True, True)):
            ctx.ntt_fr(data, k, batch, inverse=inverse, coset=coset)
        # forward then inverse, coset forward then coset inverse: the identity
        assert np.array_equal(data, want), f"NTT round trip k={k}"
    for k in (10, 13):
        ctx.srs_setup(k, TAU)
        n = 1 << k
        sc = rng.integers(0, 1 << 60, size=(3 * n, 4), dtype=np.uint64)
        sc[n:2 * n] = 0
        sc[n:2 * n, 0] = rng.integers(0, 3, size=n).astype(np.uint64)            # skewed: three values only
        sc[2 * n + 100:2 * n + 900] = sc[2 * n + 99]                              # a run of equal scalars (prefix-table path)
        a = ctx.msm_g1(sc, 3, basis=1)
        b = ctx.msm_g1(sc, 3, basis=0)
        assert len(a) == 192 and len(b) == 192
    k, unusable = 10, 20
    ctx.srs_setup(k, TAU)
    params = bfv.BfvParams(N=16, Q=536870909, T=7, B=19)
    zeros = {key: ["0"] * (17 if key == "cyclo" else 16) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(3)
    pk = prover.keygen(kg.wit, k, unusable)
    inp = synthetic_input(random.Random(44), 16, params.Q, params.T, params.B)
    chk = bfv.BfvCircuit(ctx, params, record=True)
    chk.phase0(inp).phase1(12345)
    assert chk.wit.mock() == 0
    R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    rinv = pow(1 << 256, -1, R_MOD)
    for transcript in (1, 0):
        proof, circ = prover.prove(pk, lambda: bfv.BfvCircuit(ctx, params), inp, bytes(32), transcript)
        inst = [int.from_bytes(row.tobytes(), "little") * rinv % R_MOD for row in circ.wit.download(4)]
        assert prover.verify(ctx, pk.vk_bytes(), inst, proof, ctx.srs_g2(TAU), transcript=transcript)
    pk2 = prover.import_key(ctx, pk.export_bytes())
    assert pk2.vk_bytes() == pk.vk_bytes()
    # one proof as three shards (virtual ranks): the sharded commit / grand-product / quotient / opening paths
    want, _ = prover.prove(pk, lambda: bfv.BfvCircuit(ctx, params), inp, bytes(32), 0)
    ctx.set_virtual_ranks(3)
    got, _ = prover.prove(pk, lambda: bfv.BfvCircuit(ctx, params), inp, bytes(32), 0)
    ctx.set_virtual_ranks(0)
    assert got == want
    print("small circuit (N=16, k=10): keygen, mock, prove x2, verify x2, pk export/import, 3-shard proof ok", flush=True)
    if "--full" in sys.argv:
        import json
        ctx.srs_setup(13, TAU)
        gold = os.path.join(ROOT, "tests", "golden")
        full = bfv.BfvCircuit(ctx, record=True)
        full.phase0(json.load(open(os.path.join(gold, "bfv_empty.in")))).phase1(7)
        pk13 = prover.keygen(full.wit, 13, 109)
        del full
        inp13 = json.load(open(os.path.join(gold, "bfv.in")))
        proof, circ = prover.prove(pk13, lambda: bfv.BfvCircuit(ctx), inp13, bytes(32), 0)
        inst = [int.from_bytes(row.tobytes(), "little") * rinv % R_MOD for row in circ.wit.download(4)]
        assert prover.verify(ctx, pk13.vk_bytes(), inst, proof, ctx.srs_g2(TAU), transcript=0)
        print("config 1 (N=1024, k=13): keygen, prove, verify ok", flush=True)
    print(f"sanitize_run ok: {ctx.launch_count()} kernel launches")
    ctx.close()


if __name__ == "__main__":
    main()
