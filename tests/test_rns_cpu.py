"""RNS front-end (zk-fhe_b200/rns.py), the parts that need no GPU: limb primes, the CRT identity between the per-limb
encryption equations and the encryption mod Q, and -- through the Python oracle -- that a limb circuit with the
`delta_i = (Q // T) mod q_i` override is satisfied by the limb residues."""
import numpy as np
import pytest

from oracle import bfv as obfv
from oracle.poly import Poly as OPoly


def _rns():
    from zk_fhe_b200 import rns
    return rns


@pytest.mark.parametrize("bits,limbs,N", [(109, 2, 4096), (438, 8, 16384), (60, 2, 16)])
def test_limb_primes(bits, limbs, N):
    rns = _rns()
    primes = rns.limb_primes(bits, limbs, N)
    prod = 1
    for q in primes:
        prod *= q
    assert len(set(primes)) == limbs and prod.bit_length() == bits
    assert all(rns.is_prime(q) and q % (2 * N) == 1 and q < 1 << 62 for q in primes)
    assert max(q.bit_length() for q in primes) - min(q.bit_length() for q in primes) <= 1


def _oracle_limb(par, enc, i):
    """Limb i's bfv.in by the oracle's own big-int Poly arithmetic (src/poly.rs restated)."""
    p = par.limb(i)
    q, N = p.Q, p.N
    vals = {k: [v % q for v in enc[k]] for k in ("pk0", "pk1", "u", "e0", "e1", "m")}
    cyclo = OPoly([1] + [0] * (N - 1) + [1], q.bit_length())

    def ring_mul(a, b):
        prod = OPoly(a, q.bit_length()).mul(OPoly(b, q.bit_length())).reduce_by_modulus(q)
        return prod.divide_by_cyclo(cyclo, q)[1].coefficients[-N:]

    c0 = [(r + p.delta * m + e) % q for r, m, e in zip(ring_mul(vals["pk0"], vals["u"]), vals["m"], vals["e0"])]
    c1 = [(r + e) % q for r, e in zip(ring_mul(vals["pk1"], vals["u"]), vals["e1"])]
    d = dict(vals, c0=c0, c1=c1, cyclo=[1] + [0] * (N - 1) + [1])
    return {k: [str(x) for x in d[k]] for k in rns_keys()}


def rns_keys():
    return _rns().INPUT_KEYS


def test_limb_equations_recombine_to_the_encryption_mod_Q():
    rns = _rns()
    N = 16
    par = rns.RnsParams(N=N, primes=tuple(rns.limb_primes(90, 3, N)), T=257, B=9)
    enc = rns.sample_encryption(par, np.random.default_rng(5))
    want_c0, want_c1 = rns.reference_ciphertext(par, enc)
    limbs = [_oracle_limb(par, enc, i) for i in range(3)]
    for j in range(N):
        assert rns.crt([int(l["c0"][j]) for l in limbs], par.primes) == want_c0[j]
        assert rns.crt([int(l["c1"][j]) for l in limbs], par.primes) == want_c1[j]
    # delta_i is NOT q_i // T: the limb circuit needs the override
    assert any(par.limb(i).delta != par.primes[i] // par.T for i in range(3))


def test_limb_circuit_is_satisfied_in_the_oracle():
    """oracle/bfv.py (examples/bfv.rs restated) with the delta override: every is_equal of the two final
    constrain_equality calls is 1 on the limb residues, and a limb computed with the plain q_i // T is not."""
    rns = _rns()
    N = 16
    par = rns.RnsParams(N=N, primes=tuple(rns.limb_primes(60, 2, N)), T=257, B=9)
    enc = rns.sample_encryption(par, np.random.default_rng(6))
    for i in range(2):
        p = par.limb(i)
        op = obfv.BfvParams(N=N, Q=p.Q, T=p.T, B=p.B, delta_override=p.delta)
        st = obfv.phase0(_oracle_limb(par, enc, i), op)
        ctx_gate, _ = obfv.phase1(st, 12345, op)
        tail = ctx_gate.advice[-12 * N:]                      # c1.constrain_equality: is_zero flag is cell 4 of 12
        assert all(tail[12 * j + 4] == 1 for j in range(N))
        same = lambda st_: [c.value for c in st_.computed["c0"].assigned_coefficients] == \
            [c.value for c in st_.polys["expected_c0"].assigned_coefficients]
        assert same(st)
        wrong = obfv.BfvParams(N=N, Q=p.Q, T=p.T, B=p.B)      # delta = q_i // T: the circuit's c0 differs from the limb's c0
        assert wrong.delta != op.delta
        st_w = obfv.phase0(_oracle_limb(par, enc, i), wrong)
        obfv.phase1(st_w, 12345, wrong)
        assert not same(st_w)
