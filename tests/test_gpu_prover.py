"""keygen / prove on the GPU vs the oracle.

Pinned by reference fixtures: the pinning keygen writes must equal the reference's
configs/bfv.json (column counts and all 158 break points).  Everything after that
(commitments, proof) has no reference artefact to compare with -- parity unpinned -- so it
is checked against definitions (MSM / NTT of the same columns in the oracle) and by the
oracle's independent verifier accepting the proof and rejecting mutated ones.
"""
import json
import os
import random

import numpy as np
import pytest

from oracle import bfv as obfv
from oracle import cbind, curve, field, verifier
from oracle.poly import Poly as OPoly
from tests.util import fr_to_mont_array, mont_array_to_fr

pytestmark = pytest.mark.gpu
TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF


@pytest.fixture(scope="module")
def ctx13():
    import zk_fhe_b200
    c = zk_fhe_b200.Context(0)
    c.srs_setup(13, TAU)
    yield c
    c.close()


@pytest.fixture(scope="module")
def pk13(ctx13, bfv_empty_input):
    from zk_fhe_b200 import bfv, prover
    circ = bfv.BfvCircuit(ctx13, record=True)
    circ.phase0(bfv_empty_input).phase1(7)
    return prover.keygen(circ.wit, 13, 109)


def _vk(pk, unusable_rows):
    cms = [curve.g1_from_mont_bytes(row.tobytes()) for row in pk.fixed_commitments()]
    return verifier.Vk(pk.info, unusable_rows, cms)


def test_keygen_pinning_equals_reference_bfv_json(pk13, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "bfv_pinning.json")))
    assert pk13.pinning() == ref
    info = pk13.info
    assert (info["n_gate0"], info["n_gate1"], info["n_rlc"], info["n_lookup"]) == (3, 153, 5, 36)
    assert info["n_advice"] == 197 and info["n_perm"] == 199 and info["n_chunks"] == 100 and info["instances"] == 5121


def test_keygen_fixed_columns_match_oracle_layout(pk13, ctx13, bfv_empty_input):
    tab = obfv.build_tables(bfv_empty_input, 7)
    info = pk13.info
    n = 1 << 13
    # gate selectors of a few columns (first, a middle one, the last) and the RLC selectors
    sel_cols = tab["gate0"][1] + tab["gate1"][1]
    for c in (0, 2, 3, 77, 155):
        want = [1 if q else 0 for q in sel_cols[c]] + [0] * (n - len(sel_cols[c]))
        assert mont_array_to_fr(pk13.fixed(c)) == want, f"gate selector {c}"
    for j, s in enumerate(tab["rlc"][1]):
        want = [1 if q else 0 for q in s] + [0] * (n - len(s))
        assert mont_array_to_fr(pk13.fixed(156 + j)) == want
    table = mont_array_to_fr(pk13.fixed(info["fx_table"]))
    assert table == list(range(256)) + [0] * (n - 256)
    # sigma columns: a permutation of the positions {delta^c w^r}; unconstrained positions map to themselves
    w = field.omega(13)
    seen = set()
    for c in (0, 5, 100, 196, 197, 198):
        sig = mont_array_to_fr(pk13.fixed(info["fx_sigma"] + c))
        assert len(set(sig)) == n and not (seen & set(sig))
        seen.update(sig)
        # the last row of every column is never used by the circuit: fixed point of the permutation
        assert sig[n - 1] == pow(field.FR_DELTA, c, field.R_MOD) * pow(w, n - 1, field.R_MOD) % field.R_MOD
    # commitment / coefficient / extended forms of one sigma column against the oracle's MSM and NTT
    _, gl = cbind.srs(13, TAU, want_g=False)
    col = pk13.fixed(info["fx_sigma"] + 5)
    assert np.array_equal(pk13.fixed_commitments()[info["fx_sigma"] + 5], cbind.msm(col, gl, n, 1)[0])
    coef = col.copy()
    cbind.ntt(coef, 13, 1, inverse=True)
    assert np.array_equal(pk13.fixed(info["fx_sigma"] + 5, form=1), coef)
    ext = np.zeros((4 * n, 4), np.uint64)
    ext[:n] = coef
    cbind.ntt(ext, 15, 1, coset=True)
    assert np.array_equal(pk13.fixed(info["fx_sigma"] + 5, form=2), ext)


def _prove(ctx, pk, inp, seed, params=None, transcript=0):
    from zk_fhe_b200 import bfv, prover
    proof, circ = prover.prove(pk, lambda: bfv.BfvCircuit(ctx, params or bfv.BfvParams()), inp, seed, transcript)
    inst = mont_array_to_fr(circ.wit.download(4))
    return proof, inst


def _pverify(ctx, *args, **kw):
    """The product verifier on the BLAKE2b transcript (the mutation sweeps below replay the proof dozens of times; the
    oracle's pure-Python Poseidon costs ~2.5 s per replay, so they run on the cheap hash -- the Poseidon default has
    its own test)."""
    from zk_fhe_b200 import prover
    kw.setdefault("transcript", 0)
    return prover.verify(ctx, *args, **kw)


def test_prove_with_the_default_poseidon_transcript(ctx13, pk13, bfv_input):
    """The configuration the reference's `prove` runs (snark-verifier PoseidonTranscript) is the default of
    Prover / prove / verify: both verifiers accept, both reject a flipped bit and the other hash."""
    from zk_fhe_b200 import bfv, prover
    proof, circ = prover.prove(pk13, lambda: bfv.BfvCircuit(ctx13), bfv_input, bytes(range(32)))
    inst = mont_array_to_fr(circ.wit.download(4))
    vk, vkb, s_g2 = _vk(pk13, 109), pk13.vk_bytes(), ctx13.srs_g2(TAU)
    assert verifier.verify(vk, inst, proof, TAU, transcript_kind=1)
    assert prover.verify(ctx13, vkb, inst, proof, s_g2)
    assert not prover.verify(ctx13, vkb, inst, proof, s_g2, transcript=0)
    bad = bytearray(proof)
    bad[32 * 200 + 7] ^= 4                     # inside the x coordinate of an advice commitment (32 bytes per point)
    assert not prover.verify(ctx13, vkb, inst, bytes(bad), s_g2)
    with pytest.raises(verifier.VerifyError):
        verifier.verify(vk, inst, bytes(bad), TAU, transcript_kind=1)
    # a second proof with OS-entropy blinding (the default seed) differs and verifies
    proof2, _ = prover.prove(pk13, lambda: bfv.BfvCircuit(ctx13), bfv_input)
    assert proof2 != proof and prover.verify(ctx13, vkb, inst, proof2, s_g2)


def test_proving_key_file_round_trip(ctx13, pk13, bfv_input):
    """data/<name>.pk (README.md:38): export, import on the same SRS, same proof bytes for the same seed."""
    from zk_fhe_b200 import prover
    blob = pk13.export_bytes()
    assert len(blob) > pk13.info["n_fixed"] * 8192 * 32
    pk2 = prover.import_key(ctx13, blob)
    assert pk2.info == pk13.info and pk2.pinning() == pk13.pinning() and pk2.vk_bytes() == pk13.vk_bytes()
    a, _ = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    b, _ = _prove(ctx13, pk2, bfv_input, bytes(range(32)))
    assert a == b
    from zk_fhe_b200.capi import ZkfheError
    for bad in (blob[:-1], b"XX" + blob[2:], blob[:64]):
        with pytest.raises(ZkfheError):
            prover.import_key(ctx13, bad)


def test_prove_bfv_in_and_oracle_verifier_accepts(ctx13, pk13, bfv_input):
    proof, inst = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    assert inst[:1024] == [int(x) for x in bfv_input["pk0"]]
    vk = _vk(pk13, 109)
    assert verifier.verify(vk, inst, proof, TAU)
    # the same proof under the real verifier equation: pairing against [tau]_2, no trapdoor
    from oracle import pairing
    s_g2 = pairing.g2_mul(pairing.G2_GEN, TAU)
    assert verifier.verify(vk, inst, proof, None, s_g2=s_g2)
    with pytest.raises(verifier.VerifyError):
        verifier.verify(vk, inst, proof, None, s_g2=pairing.g2_mul(pairing.G2_GEN, TAU + 1))
    # deterministic for a fixed seed, different for another seed (blinding)
    proof2, _ = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    assert proof2 == proof
    proof3, _ = _prove(ctx13, pk13, bfv_input, bytes(32))
    assert proof3 != proof and verifier.verify(vk, inst, proof3, TAU)
    # soundness smoke: any mutation is rejected
    rng = random.Random(5)
    for _ in range(4):
        bad = bytearray(proof)
        bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
        with pytest.raises(verifier.VerifyError):
            verifier.verify(vk, inst, bytes(bad), TAU)
    bad_inst = list(inst)
    bad_inst[2048 + 17] = (bad_inst[2048 + 17] + 1) % 536870909       # a different ciphertext coefficient
    with pytest.raises(verifier.VerifyError):
        verifier.verify(vk, bad_inst, proof, TAU)
    with pytest.raises(verifier.VerifyError):
        verifier.verify(vk, inst, proof, TAU + 1)                      # wrong SRS


def test_product_verifier_agrees_with_oracle_verifier(ctx13, pk13, bfv_input):
    """zkfhe_verify (the reference's `verify` subcommand: transcript replay + identities on the host, one MSM on
    the GPU, pairing on the host) accepts what the oracle verifier accepts and rejects what it rejects."""
    from zk_fhe_b200 import prover
    proof, inst = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    vkb, s_g2 = pk13.vk_bytes(), ctx13.srs_g2(TAU)
    vk = _vk(pk13, 109)
    assert len(vkb) == 72 + 64 * pk13.info["n_fixed"]
    assert _pverify(ctx13, vkb, inst, proof, s_g2)
    assert verifier.verify(vk, inst, proof, TAU)
    # verifying borrows the commitment-key slot for its MSM and must give it back intact
    proof2, _ = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    assert proof2 == proof
    rng = random.Random(11)
    # points travel compressed, as halo2 writes them: 411 x 32 bytes of commitments + 1517 x 32 bytes of evaluations
    assert len(proof) == 32 * 411 + 32 * 1517
    offsets = [(5, None), (32 * 3 + 8, None), (32 * 197 + 3, None), (len(proof) // 2, None), (len(proof) - 100, None),
               (len(proof) - 1, None),
               (32 * 5 + 31, 6),          # only the y-parity bit of a commitment: the other root, a different point
               (32 * 7 + 31, 7)]          # the identity flag on a finite point: not an encoding
    offsets += [(rng.randrange(len(proof)), None) for _ in range(6)]
    for off, bit in offsets:
        bad = bytearray(proof)
        bad[off] ^= 1 << (rng.randrange(8) if bit is None else bit)
        assert not _pverify(ctx13, vkb, inst, bytes(bad), s_g2), off
        assert "rejected" in ctx13.last_rejection
        with pytest.raises(verifier.VerifyError):
            verifier.verify(vk, inst, bytes(bad), TAU)
    assert not _pverify(ctx13, vkb, inst, proof[:-1], s_g2)
    assert not _pverify(ctx13, vkb, inst, proof + b"\0", s_g2)
    bad_inst = list(inst)
    bad_inst[2048 + 17] = (bad_inst[2048 + 17] + 1) % 536870909
    assert not _pverify(ctx13, vkb, bad_inst, proof, s_g2)
    assert not _pverify(ctx13, vkb, inst[:-1], proof, s_g2)
    assert not _pverify(ctx13, vkb, inst, proof, ctx13.srs_g2(TAU + 1))          # another SRS
    bad_vk = bytearray(vkb)
    bad_vk[72 + 64 * 200 + 7] ^= 4                                                      # a fixed commitment
    assert not _pverify(ctx13, bytes(bad_vk), inst, proof, s_g2)
    import zk_fhe_b200
    with pytest.raises(zk_fhe_b200.ZkfheError):
        _pverify(ctx13, vkb[:-3], inst, proof, s_g2)                               # malformed key: an error, not a verdict
    # `usable` (header word 9) and `n_chunks` (word 8) are not in the digest: a key that carries anything but the
    # values the layout implies is refused outright
    for word, delta in ((9, -1), (9, 1), (8, 1)):
        hdr = bytearray(vkb)
        off = 8 + 4 * word
        hdr[off:off + 4] = (int.from_bytes(hdr[off:off + 4], "little") + delta).to_bytes(4, "little")
        with pytest.raises(zk_fhe_b200.ZkfheError):
            _pverify(ctx13, bytes(hdr), inst, proof, s_g2)
    assert _pverify(ctx13, vkb, inst, proof, s_g2)


def test_proof_of_a_wrong_ciphertext_is_rejected(ctx13, pk13, bfv_input):
    import zk_fhe_b200
    inp = dict(bfv_input)
    inp["c0"] = list(inp["c0"])
    inp["c0"][3] = str((int(inp["c0"][3]) + 1) % 536870909)
    try:
        proof, inst = _prove(ctx13, pk13, inp, bytes(32))
    except zk_fhe_b200.ZkfheError as e:
        assert e.code == -6
        return
    with pytest.raises(verifier.VerifyError):
        verifier.verify(_vk(pk13, 109), inst, proof, TAU)


def _synthetic_input(rng, N, Q, T, B):
    pk0 = [rng.randrange(Q) for _ in range(N)]
    pk1 = [rng.randrange(Q) for _ in range(N)]
    u = [rng.choice([0, 1, Q - 1]) for _ in range(N)]
    e0 = [max(-B, min(B, round(rng.gauss(0, 3.2)))) % Q for _ in range(N)]
    e1 = [max(-B, min(B, round(rng.gauss(0, 3.2)))) % Q for _ in range(N)]
    m = [rng.randint(-(T // 2), T // 2) % Q for _ in range(N)]
    cyclo = [1] + [0] * (N - 1) + [1]

    def enc(pk, extra):
        P = OPoly(pk, Q.bit_length()).mul(OPoly(u, Q.bit_length())).reduce_by_modulus(Q)
        _, r = P.divide_by_cyclo(OPoly(cyclo, Q.bit_length()), Q)
        return [(a + b) % Q for a, b in zip(r.coefficients[-N:], extra)]
    c0 = enc(pk0, [((Q // T) * mi + ei) % Q for mi, ei in zip(m, e0)])
    c1 = enc(pk1, e1)
    d = dict(pk0=pk0, pk1=pk1, m=m, u=u, e0=e0, e1=e1, c0=c0, c1=c1, cyclo=cyclo)
    return {k: [str(x) for x in v] for k, v in d.items()}


def test_n4096_k16_keygen_prove_verify():
    """BASELINE.json config 3/4 shape (N = 4096, k = 16) with the widest modulus the reference can
    express (`modulus: u64`): Q = 2^61 - 1, T = 65537.  keygen on zeros, prove a synthetic
    encryption, independent verifier accepts; mock accepts the witness and rejects a tampered one."""
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover
    ctx = zk_fhe_b200.Context(0)
    k, unusable = 16, 109
    ctx.srs_setup(k, TAU)
    params = bfv.BfvParams(N=4096, Q=(1 << 61) - 1, T=65537, B=19)
    zeros = {key: ["0"] * (4097 if key == "cyclo" else 4096) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(3)
    pk = prover.keygen(kg.wit, k, unusable)
    del kg
    assert pk.info["k"] == 16 and pk.info["instances"] == 5 * 4096 + 1
    inp = _synthetic_input(random.Random(4096), 4096, params.Q, params.T, params.B)
    chk = bfv.BfvCircuit(ctx, params, record=True)
    chk.phase0(inp).phase1(12345)
    assert chk.wit.mock() == 0
    del chk
    proof, inst = _prove(ctx, pk, inp, bytes(32), params)
    vk = _vk(pk, unusable)
    assert verifier.verify(vk, inst, proof, TAU)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 4
    with pytest.raises(verifier.VerifyError):
        verifier.verify(vk, inst, bytes(bad), TAU)
    ctx.close()


@pytest.mark.parametrize("transcript", [0, 1])
def test_small_circuit_end_to_end_both_transcripts(transcript):
    """N = 16 at k = 10: keygen on zeros, prove a synthetic encryption, verify."""
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover
    ctx = zk_fhe_b200.Context(0)
    k, unusable = 10, 20
    ctx.srs_setup(k, TAU)
    params = bfv.BfvParams(N=16, Q=536870909, T=7, B=19)
    zeros = {key: ["0"] * (17 if key == "cyclo" else 16) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(3)
    pk = prover.keygen(kg.wit, k, unusable)
    inp = _synthetic_input(random.Random(44), 16, params.Q, params.T, params.B)
    proof, inst = _prove(ctx, pk, inp, bytes(32), params, transcript)
    vk = _vk(pk, unusable)
    assert verifier.verify(vk, inst, proof, TAU, transcript_kind=transcript)
    assert prover.verify(ctx, pk.vk_bytes(), inst, proof, ctx.srs_g2(TAU), transcript=transcript)
    assert not prover.verify(ctx, pk.vk_bytes(), inst, proof, ctx.srs_g2(TAU), transcript=1 - transcript)
    with pytest.raises(verifier.VerifyError):
        verifier.verify(vk, inst, proof, TAU, transcript_kind=1 - transcript)
    ctx.close()


# ---- one proof sharded over several (here: virtual) ranks ------------------------------------------------------
@pytest.mark.parametrize("G", [2, 3, 4, 5, 8, 13])
def test_sharded_proof_is_byte_identical_small_circuit(G):
    """Column-sharded commitment phases, the expression-sharded quotient and the column-sharded openings
    (zkfhe_set_virtual_ranks: all G shards computed on this GPU, no collective) must give the single-GPU proof byte
    for byte -- every rank count, incl. ones that do not divide the column / chunk / lookup counts."""
    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover
    ctx = zk_fhe_b200.Context(0)
    k, unusable = 10, 20
    ctx.srs_setup(k, TAU)
    params = bfv.BfvParams(N=16, Q=536870909, T=7, B=19)
    zeros = {key: ["0"] * (17 if key == "cyclo" else 16) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(3)
    pk = prover.keygen(kg.wit, k, unusable)
    inp = _synthetic_input(random.Random(45), 16, params.Q, params.T, params.B)
    want, inst = _prove(ctx, pk, inp, bytes(range(32)), params, 0)
    ctx.set_virtual_ranks(G)
    assert ctx.comm_info() == (0, 1, G)
    got, _ = _prove(ctx, pk, inp, bytes(range(32)), params, 0)
    assert got == want
    ctx.set_virtual_ranks(0)
    again, _ = _prove(ctx, pk, inp, bytes(range(32)), params, 0)
    assert again == want
    assert verifier.verify(_vk(pk, unusable), inst, got, TAU)
    ctx.close()


@pytest.mark.parametrize("G", [2, 8])
def test_sharded_proof_is_byte_identical_config1(ctx13, pk13, bfv_input, G):
    want, _ = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    ctx13.set_virtual_ranks(G)
    try:
        got, _ = _prove(ctx13, pk13, bfv_input, bytes(range(32)))
    finally:
        ctx13.set_virtual_ranks(0)
    assert got == want


def test_sharded_proof_over_nccl_two_gpus(tmp_path):
    """Two processes, two GPUs, one proof (tools/sharded_prove.py under torchrun): both ranks must return the bytes of
    the single-GPU proof.  Skipped on a one-GPU box (the virtual-rank tests above cover the arithmetic there)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "sharded.json"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(root, "tools", "sharded_prove.py"), "--k", "13", "--proofs", "2", "--out", str(out)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    assert res["identical_to_single_gpu"] is True and res["n_ranks"] == 2
