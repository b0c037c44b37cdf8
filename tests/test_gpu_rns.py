"""RNS front-end on the GPU: one reference circuit per limb prime (zk-fhe_b200/rns.py), each keygen'd, proved and
verified; the limb ciphertexts recombine (CRT) to the big-integer encryption mod Q; a limb proved with the plain
q_i // T scaling instead of (Q // T) mod q_i is unsatisfiable."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TAU = 0x1234567890ABCDEF1234567890ABCDEF
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_two_limb_encryption_proves_limb_by_limb_and_recombines():
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover, rns
    ctx = zk_fhe_b200.Context(0)
    k, unusable, N = 10, 20, 16
    ctx.srs_setup(k, TAU)
    par = rns.RnsParams(N=N, primes=tuple(rns.limb_primes(60, 2, N)), T=257, B=9)
    enc = rns.sample_encryption(par, np.random.default_rng(11))
    want_c0, want_c1 = rns.reference_ciphertext(par, enc)
    limbs = [rns.limb_input(ctx, par, enc, i) for i in range(2)]
    for j in range(N):
        assert rns.crt([int(l["c0"][j]) for l in limbs], par.primes) == want_c0[j]
        assert rns.crt([int(l["c1"][j]) for l in limbs], par.primes) == want_c1[j]
    rinv = pow(1 << 256, -1, R_MOD)
    for i in range(2):
        p = par.limb(i)
        assert p.delta == (par.Q // par.T) % par.primes[i]
        zeros = {key: ["0"] * (N + 1 if key == "cyclo" else N) for key in bfv.INPUT_KEYS}
        kg = bfv.BfvCircuit(ctx, p, record=True)
        kg.phase0(zeros).phase1(3)
        pk = prover.keygen(kg.wit, k, unusable)
        chk = bfv.BfvCircuit(ctx, p, record=True)
        chk.phase0(limbs[i]).phase1(12345)
        assert chk.wit.mock() == 0
        proof, circ = prover.prove(pk, lambda: bfv.BfvCircuit(ctx, p), limbs[i], bytes(32))
        inst = [int.from_bytes(row.tobytes(), "little") * rinv % R_MOD for row in circ.wit.download(4)]
        assert inst[2 * N:3 * N] == [int(x) for x in limbs[i]["c0"]]          # the limb's ciphertext is what the proof exposes
        assert prover.verify(ctx, pk.vk_bytes(), inst, proof, ctx.srs_g2(TAU))
        # the same limb under the single-modulus scaling q_i // T: the c0 equality cannot hold
        plain = bfv.BfvParams(N=N, Q=p.Q, T=p.T, B=p.B)
        bad = bfv.BfvCircuit(ctx, plain, record=True)
        bad.phase0(limbs[i]).phase1(12345)
        with pytest.raises(zk_fhe_b200.ZkfheError):
            bad.wit.mock()
    ctx.close()
