"""GPU parity for stage (1): witness generation through the C ABI vs the oracle's restatement
of src/poly.rs, src/poly_chip.rs and examples/bfv.rs -- bit-exact, cell by cell, on the
reference's own fixture (bfv.in, a known-answer vector for c0/c1) and on synthetic inputs.
"""
import hashlib
import random

import numpy as np
import pytest

from oracle import bfv as obfv
from oracle import field
from oracle.poly import Poly as OPoly
from tests.util import fr_to_mont_array

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import zk_fhe_b200
    c = zk_fhe_b200.Context(0)
    yield c
    c.close()


def _canon(arr):
    """(n,4) uint64 Montgomery limbs -> list of canonical ints."""
    b = np.ascontiguousarray(arr).tobytes()
    return [field.from_mont_fast(int.from_bytes(b[i:i + 32], "little")) for i in range(0, len(b), 32)]


def _digest(vals):
    h = hashlib.sha256()
    for v in vals:
        h.update(int(v).to_bytes(32, "little"))
    return h.hexdigest()


def _run_gpu(ctx, inp, gamma, params=None):
    from zk_fhe_b200 import bfv
    circ = bfv.BfvCircuit(ctx, params or bfv.BfvParams())
    circ.phase0(inp)
    circ.phase1(gamma)
    circ.wit.status()
    return circ


def test_poly_ops_match_poly_rs_on_reference_fixture(ctx, bfv_input):
    from zk_fhe_b200.poly import Poly
    Q = 536870909
    pk0 = Poly.from_string(ctx, bfv_input["pk0"], Q)
    u = Poly.from_string(ctx, bfv_input["u"], Q)
    cyclo = Poly.from_string(ctx, bfv_input["cyclo"], Q)
    o_pk0 = OPoly.from_string(bfv_input["pk0"], Q)
    o_u = OPoly.from_string(bfv_input["u"], Q)
    o_cyclo = OPoly.from_string(bfv_input["cyclo"], Q)
    prod, o_prod = pk0.mul(u), o_pk0.mul(o_u)
    assert prod.coefficients == o_prod.coefficients and prod.max_bits == o_prod.max_bits == 68
    red, o_red = prod.reduce_by_modulus(Q), o_prod.reduce_by_modulus(Q)
    assert red.coefficients == o_red.coefficients and red.max_bits == 29
    (q, r), (oq, orr) = red.divide_by_cyclo(cyclo, Q), o_red.divide_by_cyclo(o_cyclo, Q)
    assert q.coefficients == oq.coefficients and r.coefficients == orr.coefficients
    assert len(q) == 1025 and len(r) == 2049
    qc, oqc = q.mul(cyclo), oq.mul(o_cyclo)
    assert qc.coefficients == oqc.coefficients and qc.max_bits == oqc.max_bits == 69


def test_poly_error_behaviour_matches_reference_asserts(ctx):
    import zk_fhe_b200
    from zk_fhe_b200.poly import Poly
    E = zk_fhe_b200.ZkfheError
    Poly.from_string(ctx, ["5", "8"], 7)                    # coeff > modulus -> sticky assert (poly.rs:28)
    with pytest.raises(E) as e:
        ctx.status()
    assert e.value.code == -4 and "poly.rs:28" in str(e.value)
    ctx.status()                                            # cleared after being reported
    assert Poly.from_string(ctx, ["7", "0"], 7).coefficients == [7, 0]     # `<=`, not `<`
    ctx.status()
    # the decimal parser (poly.rs:25 `parse().unwrap()`): malformed numbers, signs, blanks and > u64 are errors
    for bad in (["12", "x3"], ["-1", "2"], ["1", ""], ["1 ", "2"], ["18446744073709551616", "1"], ["1,2", "3"]):
        with pytest.raises(E) as e:
            Poly.from_string(ctx, bad, 7)
        assert e.value.code == -2, bad
    big = Poly.from_string(ctx, ["18446744073709551615", "0", "00042"], (1 << 64) - 1)
    assert big.coefficients == [18446744073709551615, 0, 42]
    ctx.status()
    with pytest.raises(E) as e:                             # equal degrees required (poly.rs:78)
        Poly.from_string(ctx, ["1", "2"], 7).mul(Poly.from_string(ctx, ["1", "2", "3"], 7))
    assert e.value.code == -4
    a = Poly.from_string(ctx, ["3", "2", "1"], 7)
    sq = a.mul(a)
    assert sq.coefficients == [9, 12, 10, 4, 1] and sq.max_bits == 3 + 3 + 2
    # quotient strips to empty -> usize underflow at poly.rs:158
    cyc = Poly.from_string(ctx, ["1", "0", "0", "0", "1"], 536870909)
    d = Poly.from_string(ctx, ["0", "0", "0", "0", "5", "1", "2"], 536870909)
    with pytest.raises(E) as e:
        d.divide_by_cyclo(cyc, 536870909)
    assert "poly.rs:158" in str(e.value)
    # all-zero dividend: shortcut, whatever the divisor holds (keygen path, poly.rs:118-123)
    z = Poly.from_string(ctx, ["0"] * 7, 536870909)
    zc = Poly.from_string(ctx, ["0"] * 5, 536870909)
    q, r = z.divide_by_cyclo(zc, 536870909)
    assert q.coefficients == [0] * 5 and r.coefficients == [0] * 9
    with pytest.raises(E) as e:                             # product would not be exact over Fr
        Poly.from_big_int(ctx, [1 << 130, 1], 131).mul(Poly.from_big_int(ctx, [1 << 130, 1], 131))
    assert e.value.code == -5
    with pytest.raises(E):                                  # from_big_int bits assert (poly.rs:51)
        Poly.from_big_int(ctx, [255, 256], 8)
        ctx.status()


def test_bfv_in_advice_tables_match_oracle_cell_by_cell(ctx, bfv_input, oracle_tables, golden_gamma, digests):
    circ = _run_gpu(ctx, bfv_input, golden_gamma)
    counts = circ.wit.counts()
    assert counts["advice"] == [23558, 1231992, 32764]
    assert counts["lookups"] == 286756 and counts["instances"] == 5121
    ph0 = _canon(circ.wit.download(0))
    assert ph0 == oracle_tables["phase0"].ctx.advice
    gate = circ.wit.download(1)
    want = fr_to_mont_array(oracle_tables["ctx_gate"].advice)
    if not np.array_equal(gate, want):
        bad = np.nonzero((gate != want).any(axis=1))[0]
        pytest.fail(f"phase-1 gate advice differs at {len(bad)} cells, first at flat offset {bad[0]}")
    assert _canon(circ.wit.download(2)) == oracle_tables["ctx_rlc"].advice
    lk = _canon(circ.wit.download(3))
    assert lk == [v for col in oracle_tables["lookup"] for v in col]
    assert _canon(circ.wit.download(4)) == oracle_tables["instances"]
    # and against the committed golden digests (no oracle run needed to check these)
    assert _digest(ph0) == digests["phase0_advice_sha256"]
    assert _digest(lk) == digests["lookup_cells_sha256"]


def test_bfv_in_ciphertext_known_answer(ctx, bfv_input, golden_gamma):
    """c0, c1 in bfv.in are the reference's known answers for the whole stage-(1) algebra."""
    circ = _run_gpu(ctx, bfv_input, golden_gamma)
    gate = circ.wit.download(1)
    for name in ("c0", "c1"):
        ap = circ.computed[name].ap
        idx = ap.base + np.arange(ap.len) * ap.stride
        assert _canon(gate[idx]) == [int(x) for x in bfv_input[name]]


def test_bfv_empty_in_keeps_shape_and_is_all_consistent(ctx, bfv_empty_input):
    circ = _run_gpu(ctx, bfv_empty_input, 5)
    tab = obfv.build_tables(bfv_empty_input, 5)
    assert circ.wit.counts()["advice"] == [23558, 1231992, 32764]
    assert _canon(circ.wit.download(0)) == tab["phase0"].ctx.advice
    assert np.array_equal(circ.wit.download(1), fr_to_mont_array(tab["ctx_gate"].advice))
    assert _canon(circ.wit.download(2)) == tab["ctx_rlc"].advice


def _synthetic_input(rng, N, Q, T, B):
    """SURVEY.md §8(d) recipe, scaled down."""
    def neg(x):
        return x % Q
    pk0 = [rng.randrange(Q) for _ in range(N)]
    pk1 = [rng.randrange(Q) for _ in range(N)]
    u = [rng.choice([0, 1, Q - 1]) for _ in range(N)]
    e0 = [neg(max(-B, min(B, round(rng.gauss(0, 3.2))))) for _ in range(N)]
    e1 = [neg(max(-B, min(B, round(rng.gauss(0, 3.2))))) for _ in range(N)]
    m = [neg(rng.randint(-(T // 2), T // 2)) for _ in range(N)]
    cyclo = [1] + [0] * (N - 1) + [1]
    delta = Q // T

    def enc(pk, extra):
        P = OPoly(pk, Q.bit_length()).mul(OPoly(u, Q.bit_length())).reduce_by_modulus(Q)
        _, r = P.divide_by_cyclo(OPoly(cyclo, Q.bit_length()), Q)
        rem = r.coefficients[-N:]
        return [(a + b) % Q for a, b in zip(rem, extra)]
    c0 = enc(pk0, [(delta * mi + ei) % Q for mi, ei in zip(m, e0)])
    c1 = enc(pk1, e1)
    d = dict(pk0=pk0, pk1=pk1, m=m, u=u, e0=e0, e1=e1, c0=c0, c1=c1, cyclo=cyclo)
    return {k: [str(x) for x in v] for k, v in d.items()}


@pytest.mark.parametrize("N,Q,T,B", [(16, 536870909, 7, 19), (64, 1032193, 5, 6), (256, (1 << 61) - 1, 65537, 19),
                                      (4096, 536870909, 7, 19)])
def test_synthetic_witness_matches_oracle(ctx, N, Q, T, B):
    from zk_fhe_b200 import bfv
    rng = random.Random(N * 31 + T)
    inp = _synthetic_input(rng, N, Q, T, B)
    gamma = rng.randrange(field.R_MOD)
    circ = _run_gpu(ctx, inp, gamma, bfv.BfvParams(N=N, Q=Q, T=T, B=B))
    op = obfv.BfvParams(N=N, Q=Q, T=T, B=B)
    st = obfv.phase0(inp, op)
    g, r = obfv.phase1(st, gamma, op)
    assert _canon(circ.wit.download(0)) == st.ctx.advice
    assert np.array_equal(circ.wit.download(1), fr_to_mont_array(g.advice))
    assert _canon(circ.wit.download(2)) == r.advice
    assert _canon(circ.wit.download(3)) == [c.value for c in g.cells_to_lookup]
    assert _canon(circ.wit.download(4)) == [c.value for c in st.make_public]


def test_invalid_witness_is_still_assigned_like_the_reference(ctx, bfv_input, golden_gamma):
    """A wrong ciphertext does not change witness generation (constraints fail later, in
    mock/prove): is_equal then takes the non-zero branch and assigns a real field inverse."""
    inp = dict(bfv_input)
    inp["c0"] = list(inp["c0"])
    inp["c0"][5] = str((int(inp["c0"][5]) + 1) % 536870909)
    circ = _run_gpu(ctx, inp, golden_gamma)
    st = obfv.phase0(inp)
    g, _ = obfv.phase1(st, golden_gamma)
    assert np.array_equal(circ.wit.download(1), fr_to_mont_array(g.advice))


def test_chip_overflow_guards(ctx):
    import zk_fhe_b200
    from zk_fhe_b200.poly import Poly
    from zk_fhe_b200.poly_chip import PolyChip, Witness
    w = Witness(ctx)
    a = PolyChip.from_poly(Poly.from_big_int(ctx, [1, 2, 3], 252), w)
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        a.add(a, ctx_gate=0).add(a, ctx_gate=0)             # 252 -> 253 -> 254 bits: add's guard
    assert e.value.code == -5 and "add" in str(e.value)
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        a.constrain_coefficients_in_range(9, 9)             # z < y
    assert e.value.code == -4
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        a.constrain_mul(a, a)                               # challenge not set yet
    assert e.value.code == -3


# ---------------------------------------------------------------- structure / mock -----
def _cell_index(ctx_sizes):
    """global index of (ctx, off) in the concatenation of the three contexts"""
    base = np.concatenate([[0], np.cumsum(ctx_sizes)])
    return base


def _partition_labels(n_total, pairs):
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    if len(pairs):
        p = np.asarray(pairs, dtype=np.int64)
        g = coo_matrix((np.ones(len(p), np.int8), (p[:, 0], p[:, 1])), shape=(n_total, n_total))
    else:
        g = coo_matrix((n_total, n_total), dtype=np.int8)
    return connected_components(g, directed=False)[1]


def test_recorded_structure_matches_oracle_and_mock_accepts(ctx, bfv_input, oracle_tables, golden_gamma):
    """Selectors, constant constraints, the copy-constraint partition, lookup wiring and public
    cells recorded by the kernels == what the oracle's halo2-base restatement records."""
    from zk_fhe_b200 import bfv
    circ = bfv.BfvCircuit(ctx, record=True)
    circ.phase0(bfv_input)
    circ.phase1(golden_gamma)
    circ.wit.status()
    octx = [oracle_tables["phase0"].ctx, oracle_tables["ctx_gate"], oracle_tables["ctx_rlc"]]
    sizes = [len(c.advice) for c in octx]
    base = _cell_index(sizes)
    n_total = int(base[-1])
    gpu_pairs, gpu_consts = [], set()
    for cid in range(3):
        flags, copy = circ.wit.structure(cid)
        assert np.array_equal((flags & 1).astype(bool), np.array(octx[cid].selector, dtype=bool)), f"selectors ctx {cid}"
        has = np.nonzero(copy)[0]
        src_ctx = (copy[has] >> np.uint64(60)).astype(np.int64) - 1
        src_off = (copy[has] & np.uint64((1 << 60) - 1)).astype(np.int64)
        gpu_pairs.append(np.stack([base[cid] + has, base[src_ctx] + src_off], axis=1))
        vals = None
        for off in np.nonzero(flags & 2)[0]:
            if vals is None:
                vals = _canon(circ.wit.download(cid))
            gpu_consts.add((cid, int(off), vals[off]))
        gpu_consts |= {(cid, int(o), 0) for o in np.nonzero(flags & 4)[0]}
        gpu_consts |= {(cid, int(o), 1) for o in np.nonzero(flags & 8)[0]}
    o_pairs, o_consts = [], set()
    for c in octx:
        for (c1, o1), (c2, o2) in c.advice_equality:
            o_pairs.append((base[c1] + o1, base[c2] + o2))
        for const, (c1, o1) in c.constant_equality:
            o_consts.add((c1, o1, const))
    assert gpu_consts == o_consts
    la = _partition_labels(n_total, np.concatenate(gpu_pairs))
    lb = _partition_labels(n_total, o_pairs)
    # same partition <=> the label pairs form a bijection
    pairs = np.unique(np.stack([la, lb], axis=1), axis=0)
    assert len(pairs) == la.max() + 1 == lb.max() + 1
    # lookup wiring and instances
    src = circ.wit.lookup_sources()
    want = [((c.ctx + 1) << 60) | c.offset for c in octx[1].cells_to_lookup]
    assert [int(x) for x in src] == want
    pub = circ.wit.public_cells()
    assert [int(x) for x in pub] == [((c.ctx + 1) << 60) | c.offset for c in oracle_tables["phase0"].make_public]
    assert circ.wit.mock() == 0


def test_mock_rejects_a_wrong_ciphertext_and_a_bad_range(ctx, bfv_input, golden_gamma):
    import zk_fhe_b200
    from zk_fhe_b200 import bfv
    inp = dict(bfv_input)
    inp["c1"] = list(inp["c1"])
    inp["c1"][100] = str((int(inp["c1"][100]) + 1) % 536870909)
    circ = bfv.BfvCircuit(ctx, record=True)
    circ.phase0(inp).phase1(golden_gamma)
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        circ.wit.mock()
    assert e.value.code == -6 and "1 constraint violations" in str(e.value)
    inp = dict(bfv_input)
    inp["e0"] = list(inp["e0"])
    inp["e0"][7] = "20"                                     # outside [0, 19] u [Q-19, Q-1]
    circ = bfv.BfvCircuit(ctx, record=True)
    circ.phase0(inp).phase1(golden_gamma)
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        circ.wit.mock()
    assert e.value.code == -6
    w = bfv.BfvCircuit(ctx).wit                             # not recording -> mock is a state error
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        w.mock()
    assert e.value.code == -3
