"""Oracle vs the reference's own known-answer material:
  * data/bfv/bfv.in  : c0, c1 are known answers for the whole stage-(1) algebra
  * configs/bfv.json : column counts and every break point pin the layout
  * data/bfv/bfv_empty.in + poly.rs:118-123 : the keygen / zero path
"""
import hashlib
import json
import os
import random

import pytest

from oracle import bfv, field
from oracle.poly import OracleError, Poly, divide_by_cyclo_closed_form, log2_ceil


def _digest(vals):
    h = hashlib.sha256()
    for v in vals:
        h.update(int(v).to_bytes(32, "little"))
    return h.hexdigest()


def test_bfv_in_ciphertext_known_answer(oracle_tables, bfv_input):
    comp = oracle_tables["phase0"].computed
    assert [c.value for c in comp["c0"].assigned_coefficients] == [int(x) for x in bfv_input["c0"]]
    assert [c.value for c in comp["c1"].assigned_coefficients] == [int(x) for x in bfv_input["c1"]]


def test_layout_matches_reference_pinning(oracle_tables, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "bfv_pinning.json")))
    assert oracle_tables["pinning"]["params"] == ref["params"]
    assert oracle_tables["pinning"]["break_points"] == ref["break_points"]
    assert len(ref["break_points"]["gate"][1]) == 152


def test_cell_counts_match_survey_appendix_e(oracle_tables):
    assert len(oracle_tables["phase0"].ctx.advice) == 23558
    assert len(oracle_tables["ctx_gate"].advice) == 1231992
    assert len(oracle_tables["ctx_rlc"].advice) == 32764
    assert sum(len(c) for c in oracle_tables["lookup"]) == 286756
    assert len(oracle_tables["instances"]) == 5121
    assert all(0 <= v < 256 for col in oracle_tables["lookup"] for v in col)


def test_table_digests_are_stable(oracle_tables, digests):
    assert _digest(oracle_tables["phase0"].ctx.advice) == digests["phase0_advice_sha256"]
    assert _digest(oracle_tables["ctx_gate"].advice) == digests["phase1_gate_advice_sha256"]
    assert _digest(oracle_tables["ctx_rlc"].advice) == digests["phase1_rlc_advice_sha256"]
    assert _digest(v for c in oracle_tables["lookup"] for v in c) == digests["lookup_cells_sha256"]


def test_every_gate_is_satisfied(oracle_tables, golden_gamma):
    """MockProver-style check of the vertical gate and the RLC gate on the flat
    contexts (the reference's only 'test' is its mock subcommand)."""
    r = field.R_MOD
    for ctx in (oracle_tables["phase0"].ctx, oracle_tables["ctx_gate"]):
        a = ctx.advice
        for i, q in enumerate(ctx.selector):
            if q:
                assert (a[i] + a[i + 1] * a[i + 2] - a[i + 3]) % r == 0, i
    ctx = oracle_tables["ctx_rlc"]
    a = ctx.advice
    n = 0
    for i, q in enumerate(ctx.selector):
        if q:
            n += 1
            assert (a[i] * golden_gamma + a[i + 1] - a[i + 2]) % r == 0, i
    assert n == 2 * (1024 + 1024 + 2047 + 1025 + 1025 + 2049) - 12   # one RLC gate per Horner step
    # copy constraints hold on values
    ctxs = {0: oracle_tables["phase0"].ctx, 1: oracle_tables["ctx_gate"], 2: oracle_tables["ctx_rlc"]}
    for c in ctxs.values():
        for (c1, o1), (c2, o2) in c.advice_equality:
            assert ctxs[c1].advice[o1] == ctxs[c2].advice[o2]
        for const, (c1, o1) in c.constant_equality:
            assert ctxs[c1].advice[o1] == const


def test_empty_input_takes_zero_shortcut_and_keeps_shape(bfv_empty_input, oracle_tables):
    tab = bfv.build_tables(bfv_empty_input, gamma=5)
    assert tab["pinning"] == oracle_tables["pinning"]   # keygen on bfv_empty.in pins the same layout


def test_divide_by_cyclo_closed_form_equals_long_division():
    rng = random.Random(7)
    for N in (4, 8, 32):
        Q = 536870909
        cyclo = Poly([1] + [0] * (N - 1) + [1], 29)
        for _ in range(5):
            D = [rng.randrange(Q) for _ in range(2 * N - 1)]
            D[0] = rng.randrange(1, Q)
            q, r = Poly(D, 29).divide_by_cyclo(cyclo, Q)
            q2, r2 = divide_by_cyclo_closed_form(D, N, Q)
            assert q.coefficients == q2 and r.coefficients == r2
            assert len(q2) == N + 1 and len(r2) == 2 * N + 1


def test_poly_error_behaviour():
    with pytest.raises(OracleError):
        Poly.from_string(["5", "8"], 7)                      # poly.rs:28
    assert Poly.from_string(["7", "0"], 7).coefficients == [7, 0]   # `<=`, not `<`
    with pytest.raises(OracleError):
        Poly.from_string(["1", "2"], 7).mul(Poly.from_string(["1", "2", "3"], 7))  # poly.rs:78
    a = Poly.from_string(["3", "2", "1"], 7)
    assert a.mul(a).coefficients == [9, 12, 10, 4, 1]
    assert a.mul(a).max_bits == 3 + 3 + log2_ceil(3)
    assert log2_ceil(1025) == 11 and log2_ceil(1024) == 10 and log2_ceil(1) == 0
    with pytest.raises(OracleError):                          # quotient strips to empty (poly.rs:158)
        Poly([0, 0, 0, 0, 5, 1, 2], 29).divide_by_cyclo(Poly([1, 0, 0, 0, 1], 29), 536870909)
