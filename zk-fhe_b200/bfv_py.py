"""Witness front-end: make a `bfv.in` for arbitrary (N, Q, T, B).

The reference's README.md:25 points at an external `bfv-py` script for this; SURVEY.md §8(f) rank 3.
A BFV public key and one encryption are sampled on the host (numpy, seeded), the ring arithmetic
c0 = pk0*u + delta*m + e0, c1 = pk1*u + e1 in Z_Q[x]/(x^N + 1) runs on the GPU with the library's own
`Poly::mul` / `reduce_by_modulus` / `divide_by_cyclo` (so the file is consistent with what the circuit
recomputes), and the result has the schema of examples/bfv.rs:50-61 (nine arrays of decimal strings,
highest-degree coefficient first).

    python -m zk_fhe_b200.bfv_py --n 1024 --q 536870909 --t 7 --b 19 --seed 0 --out data/bfv/synth.in

TEST / BENCHMARK INPUT GENERATOR ONLY: secrets, encryption randomness and errors come from a SEEDED numpy
generator (default seed 0), so every value in the file is publicly reproducible.  Do not use it to encrypt
anything real; a real front-end samples s, u, e0, e1 from a CSPRNG.
"""
import argparse
import json

import numpy as np

from .bfv import INPUT_KEYS, BfvParams
from .poly import Poly


def encrypt_with(ctx, params, values):
    """c0, c1 for GIVEN pk0, pk1, u, e0, e1, m (ints in [0, Q)) with the library's Poly arithmetic; returns the bfv.in dict.
    `params.delta` is the scaling factor (Q // T, or an RNS limb's (Q_total // T) mod q_i)."""
    N, Q = params.N, params.Q
    cyclo = [1] + [0] * (N - 1) + [1]
    pc = Poly.from_string(ctx, [str(x) for x in cyclo], Q)

    def ring_mul(a, b):
        pa, pb = Poly.from_string(ctx, [str(x) for x in a], Q), Poly.from_string(ctx, [str(x) for x in b], Q)
        _, rem = pa.mul(pb).reduce_by_modulus(Q).divide_by_cyclo(pc, Q)
        return rem.coefficients[-N:]

    delta = params.delta
    c0 = [(r + delta * mi + ei) % Q for r, mi, ei in zip(ring_mul(values["pk0"], values["u"]), values["m"], values["e0"])]
    c1 = [(r + ei) % Q for r, ei in zip(ring_mul(values["pk1"], values["u"]), values["e1"])]
    d = dict(values, c0=c0, c1=c1, cyclo=cyclo)
    return {k: [str(x) for x in d[k]] for k in INPUT_KEYS}


def keygen_and_encrypt(ctx, params=BfvParams(), rng=None, with_secret_key=True):
    """One synthetic (public key, message, randomness, ciphertext) tuple as a bfv.in dict.

    with_secret_key: pk = (-(a*s) + e, a) for a ternary secret s, as BFV key generation does; otherwise pk0, pk1
    are uniform (what SURVEY.md §8(d) specifies for throughput runs -- the circuit does not look at s)."""
    rng = rng or np.random.default_rng(0)
    N, Q, T, B = params.N, params.Q, params.T, params.B
    if Q >= 1 << 63:
        raise ValueError("Q must fit the reference's `modulus: u64`")
    cyclo = [1] + [0] * (N - 1) + [1]

    def uniform():
        return [int(x) for x in rng.integers(0, Q, N, dtype=np.uint64)]

    def ternary():
        return [Q - 1 if x == 2 else int(x) for x in rng.integers(0, 3, N)]

    def error():
        return [int(x) % Q for x in np.clip(np.rint(rng.normal(0, 3.2, N)), -B, B).astype(np.int64)]

    def poly(v):
        return Poly.from_string(ctx, [str(x) for x in v], Q)

    pc = poly(cyclo)

    def ring_mul(a, b):                       # a*b mod (x^N + 1, Q), N coefficients, highest degree first
        red = poly(a).mul(poly(b)).reduce_by_modulus(Q)
        _, rem = red.divide_by_cyclo(pc, Q)
        return rem.coefficients[-N:]

    if with_secret_key:
        a, s, e = uniform(), ternary(), error()
        pk0 = [(-x + y) % Q for x, y in zip(ring_mul(a, s), e)]
        pk1 = a
    else:
        pk0, pk1 = uniform(), uniform()
    u, e0, e1 = ternary(), error(), error()
    m = [int(x) % Q for x in rng.integers(-(T // 2), T // 2 + 1, N)]
    delta = params.delta                       # Q // T, or the RNS limb's (Q_total // T) mod q_i
    c0 = [(r + delta * mi + ei) % Q for r, mi, ei in zip(ring_mul(pk0, u), m, e0)]
    c1 = [(r + ei) % Q for r, ei in zip(ring_mul(pk1, u), e1)]
    d = dict(pk0=pk0, pk1=pk1, m=m, u=u, e0=e0, e1=e1, c0=c0, c1=c1, cyclo=cyclo)
    return {k: [str(x) for x in d[k]] for k in INPUT_KEYS}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--q", type=int, default=536870909)
    ap.add_argument("--t", type=int, default=7)
    ap.add_argument("--b", type=int, default=19)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    from . import Context
    ctx = Context(args.device)
    inp = keygen_and_encrypt(ctx, BfvParams(N=args.n, Q=args.q, T=args.t, B=args.b), np.random.default_rng(args.seed))
    with open(args.out, "w") as f:
        json.dump(inp, f)
    print(f"wrote {args.out}: N={args.n} Q={args.q} T={args.t} B={args.b}")
    ctx.close()


if __name__ == "__main__":
    main()
