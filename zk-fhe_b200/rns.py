"""RNS front-end: BFV with a ciphertext modulus Q = q_0 * ... * q_{L-1} wider than the reference's `modulus: u64`.

BASELINE.json's configurations 3-5 name moduli the reference cannot express (109 bit at N = 4096, 438 bit at
N = 16384; `const Q: u64` at examples/bfv.rs:28, `modulus: u64` at src/poly.rs:21,113,180 and src/poly_chip.rs:193,230,361).
HE libraries hold such a Q in residue form, and so does this module: ONE instance of the reference's circuit per limb
prime q_i (each q_i < 2^63, so every limb is a circuit the reference's own API could state), proving

    c0 = pk0 * u + delta_i * m + e0   (mod q_i, mod x^N + 1)          delta_i = (Q // T) mod q_i
    c1 = pk1 * u + e1                 (mod q_i, mod x^N + 1)

for the residues of the same (pk0, pk1, c0, c1) and the same small (u, m, e0, e1).  By the CRT the L limb statements
together are the encryption equation mod Q.  What the limb proofs do NOT prove is that u, m, e0, e1 are the same in
every limb: they are private inputs of independent proofs (a cross-limb consistency argument -- committing to them once
and opening in every limb -- is the natural next step and is stated here, not built).
Limbs are independent proofs: they shard over GPUs with no data-path collective (config 5: 8 limbs on 8 GPUs).
"""
from dataclasses import dataclass

import numpy as np

from .bfv import INPUT_KEYS, BfvParams


def is_prime(n):
    """Deterministic Miller-Rabin for n < 2^64."""
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def limb_primes(total_bits, limbs, N):
    """`limbs` distinct primes q = 1 (mod 2N) (NTT-friendly, as BFV libraries choose them) whose product has exactly
    `total_bits` bits: limb widths differ by at most one bit, each prime is the largest candidate below its width."""
    assert limbs >= 1 and total_bits >= limbs * 20
    base, extra = divmod(total_bits, limbs)
    widths = [base + (1 if i < extra else 0) for i in range(limbs)]
    assert max(widths) <= 62, "a limb must fit the reference's u64 modulus with room for the circuit's range checks"
    for _ in range(8):
        primes, step = [], 2 * N
        for w in widths:
            q = ((1 << w) - 1) // step * step + 1
            while q > (1 << (w - 1)) and (not is_prime(q) or q in primes):
                q -= step
            assert q > (1 << (w - 1)), f"no {w}-bit prime = 1 mod {step}"
            primes.append(q)
        prod = 1
        for q in primes:
            prod *= q
        if prod.bit_length() == total_bits:
            return primes
        widths[-1] += 1          # the largest primes of each width fell one bit short in the product: widen the last limb
    raise ValueError("no limb set found")


@dataclass
class RnsParams:
    N: int
    primes: tuple
    T: int
    B: int = 19

    @property
    def Q(self):
        q = 1
        for p in self.primes:
            q *= p
        return q

    def limb(self, i):
        """The reference circuit's parameters for limb i."""
        return BfvParams(N=self.N, Q=self.primes[i], T=self.T, B=self.B, delta_override=(self.Q // self.T) % self.primes[i])


def sample_encryption(params, rng):
    """One BFV encryption over Z_Q: uniform public key, ternary u, rounded-Gaussian errors clipped to +-B, message in
    [-T/2, T/2] -- as big integers, centred representatives for the small values.  TEST / BENCHMARK inputs (seeded)."""
    N, Q, T, B = params.N, params.Q, params.T, params.B
    nbytes = (Q.bit_length() + 7) // 8 + 8

    def uniform():
        return [int.from_bytes(rng.bytes(nbytes), "little") % Q for _ in range(N)]

    small = {"u": [int(x) - 1 for x in rng.integers(0, 3, N)],
             "e0": [int(x) for x in np.clip(np.rint(rng.normal(0, 3.2, N)), -B, B).astype(np.int64)],
             "e1": [int(x) for x in np.clip(np.rint(rng.normal(0, 3.2, N)), -B, B).astype(np.int64)],
             "m": [int(x) for x in rng.integers(-(T // 2), T // 2 + 1, N)]}
    return {"pk0": uniform(), "pk1": uniform(), **small}


def limb_input(ctx, params, enc, i):
    """The bfv.in dict of limb i: residues of the public key, the small values mapped into [0, q_i), c0 / c1 computed by
    the library's own Poly arithmetic (device) for that limb."""
    from . import bfv_py
    p = params.limb(i)
    q = p.Q
    fixed = {k: [v % q for v in enc[k]] for k in ("pk0", "pk1", "u", "e0", "e1", "m")}
    return bfv_py.encrypt_with(ctx, p, fixed)


def crt(residues, primes):
    """The integer in [0, prod primes) with the given residues."""
    Q = 1
    for p in primes:
        Q *= p
    x = 0
    for r, p in zip(residues, primes):
        m = Q // p
        x += r * m * pow(m, -1, p)
    return x % Q


def reference_ciphertext(params, enc):
    """c0, c1 mod Q by plain big-integer negacyclic convolution (O(N^2): small N only) -- the checker for the limbs."""
    N, Q, T = params.N, params.Q, params.T

    def ring_mul(a, b):
        out = [0] * N
        for i in range(N):                  # big-endian: index 0 = x^(N-1)
            for j in range(N):
                d = (N - 1 - i) + (N - 1 - j)
                if d >= N:
                    out[N - 1 - (d - N)] -= a[i] * b[j]
                else:
                    out[N - 1 - d] += a[i] * b[j]
        return [v % Q for v in out]

    delta = Q // T
    c0 = [(x + delta * m + e) % Q for x, m, e in zip(ring_mul(enc["pk0"], enc["u"]), enc["m"], enc["e0"])]
    c1 = [(x + e) % Q for x, e in zip(ring_mul(enc["pk1"], enc["u"]), enc["e1"])]
    return c0, c1


__all__ = ["RnsParams", "limb_primes", "sample_encryption", "limb_input", "crt", "reference_ciphertext", "INPUT_KEYS"]
