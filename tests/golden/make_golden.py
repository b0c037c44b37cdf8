#!/usr/bin/env python3
"""Regenerates tests/golden/* from /root/reference (run in the build container
only; the GPU box has no /root/reference).

  bfv.in, bfv_empty.in   minified copies of the reference's witness fixtures
                         (data/bfv/*.in) -- the known-answer vectors (c0, c1)
  bfv_pinning.json       minified copy of configs/bfv.json (layout KAT)
  oracle_digests.json    sha256 digests of the oracle's advice tables for
                         bfv.in under a fixed gamma, plus small NTT / MSM /
                         field vectors computed by the Python oracle
"""
import hashlib
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import bfv, curve, field, ntt  # noqa: E402

GAMMA = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF % field.R_MOD


def digest_ints(vals):
    h = hashlib.sha256()
    for v in vals:
        h.update(int(v).to_bytes(32, "little"))
    return h.hexdigest()


def main():
    for name in ("bfv.in", "bfv_empty.in"):
        d = json.load(open(f"{REF}/data/bfv/{name}"))
        json.dump(d, open(f"{HERE}/{name}", "w"), separators=(",", ":"))
    pin = json.load(open(f"{REF}/configs/bfv.json"))
    json.dump(pin, open(f"{HERE}/bfv_pinning.json", "w"), separators=(",", ":"))

    inp = bfv.load_input(f"{HERE}/bfv.in")
    tab = bfv.build_tables(inp, GAMMA)
    out = {
        "gamma": hex(GAMMA),
        "phase0_advice_sha256": digest_ints(tab["phase0"].ctx.advice),
        "phase1_gate_advice_sha256": digest_ints(tab["ctx_gate"].advice),
        "phase1_rlc_advice_sha256": digest_ints(tab["ctx_rlc"].advice),
        "lookup_cells_sha256": digest_ints(v for col in tab["lookup"] for v in col),
        "instances_sha256": digest_ints(tab["instances"]),
        "counts": {
            "phase0": len(tab["phase0"].ctx.advice),
            "phase1_gate": len(tab["ctx_gate"].advice),
            "phase1_rlc": len(tab["ctx_rlc"].advice),
            "lookups": sum(len(c) for c in tab["lookup"]),
            "instances": len(tab["instances"]),
        },
    }
    rng = random.Random(20261017)
    # small NTT vectors (k = 4, 8) checked against the O(n^2) definition
    vecs = {}
    for k in (4, 8):
        a = [rng.randrange(field.R_MOD) for _ in range(1 << k)]
        f = ntt.ntt(a, k)
        assert f == ntt.dft_naive(a, k)
        vecs[f"k{k}"] = {"in": [hex(x) for x in a], "out": [hex(x) for x in f]}
    out["ntt"] = vecs
    # small MSM vector: 16 points = multiples of the generator
    pts = [curve.g1_mul(curve.G1_GEN, rng.randrange(1, field.R_MOD)) for _ in range(16)]
    sc = [rng.randrange(field.R_MOD) for _ in range(16)]
    res = curve.msm_naive(sc, pts)
    assert res == curve.msm_pippenger(sc, pts, c=4)
    out["msm"] = {"points": [[hex(p[0]), hex(p[1])] for p in pts],
                  "scalars": [hex(s) for s in sc], "result": [hex(res[0]), hex(res[1])]}
    json.dump(out, open(f"{HERE}/oracle_digests.json", "w"), indent=1)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
