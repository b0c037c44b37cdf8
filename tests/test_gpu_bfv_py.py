"""The witness front-end (zk_fhe_b200.bfv_py, the reference README's `bfv-py` step): the inputs it writes
satisfy the circuit and agree with the oracle's restatement of src/poly.rs on the same polynomials."""
import random

import numpy as np
import pytest

from oracle.poly import Poly as OPoly

pytestmark = pytest.mark.gpu


def _oracle_ring_mul(a, b, N, Q):
    cyclo = OPoly([1] + [0] * (N - 1) + [1], Q.bit_length())
    prod = OPoly([int(x) for x in a], Q.bit_length()).mul(OPoly([int(x) for x in b], Q.bit_length())).reduce_by_modulus(Q)
    _, rem = prod.divide_by_cyclo(cyclo, Q)
    return rem.coefficients[-N:]


@pytest.mark.parametrize("N,Q,T,B,with_sk", [(16, 536870909, 7, 19, True), (64, 1099511627689, 17, 19, True),
                                             (1024, 536870909, 7, 19, False)])
def test_generated_input_is_a_valid_encryption_and_satisfies_the_circuit(N, Q, T, B, with_sk):
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, bfv_py
    ctx = zk_fhe_b200.Context(0)
    params = bfv.BfvParams(N=N, Q=Q, T=T, B=B)
    inp = bfv_py.keygen_and_encrypt(ctx, params, np.random.default_rng(N + T), with_secret_key=with_sk)
    assert set(inp) == set(bfv.INPUT_KEYS) and len(inp["cyclo"]) == N + 1 and all(len(inp[k]) == N for k in bfv.INPUT_KEYS[:-1])
    iv = {k: [int(x) for x in v] for k, v in inp.items()}
    assert all(0 <= x < Q for k in bfv.INPUT_KEYS for x in iv[k])
    assert set(iv["u"]) <= {0, 1, Q - 1}
    assert all(min(x, Q - x) <= B for x in iv["e0"] + iv["e1"]) and all(min(x, Q - x) <= T // 2 for x in iv["m"])
    # the ciphertext is the BFV encryption of m under pk with randomness (u, e0, e1): oracle recomputation
    delta = Q // T
    r0 = _oracle_ring_mul(iv["pk0"], iv["u"], N, Q)
    r1 = _oracle_ring_mul(iv["pk1"], iv["u"], N, Q)
    assert iv["c0"] == [(r + delta * m + e) % Q for r, m, e in zip(r0, iv["m"], iv["e0"])]
    assert iv["c1"] == [(r + e) % Q for r, e in zip(r1, iv["e1"])]
    # and the circuit accepts it (mock prover: every gate, copy and lookup constraint)
    circ = bfv.BfvCircuit(ctx, params, record=True)
    circ.phase0(inp).phase1(random.Random(1).randrange(1 << 200))
    assert circ.wit.mock() == 0
    ctx.close()
