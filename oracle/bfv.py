"""BFV encryption circuit (oracle; test-only).

Restates /root/reference/examples/bfv.rs:63-304 (call order is the contract:
it fixes the order of cells in the advice table) and the column layout done by
halo2-base's thread builder [UPSTREAM-RECALL, SURVEY App. B], pinned by
/root/reference/configs/bfv.json.
"""
import json
from dataclasses import dataclass, field

from .halo2_base import Context, RangeChip, RlcChip
from .poly import Poly, _check
from .poly_chip import PolyChip


@dataclass
class BfvParams:
    """examples/bfv.rs:27-30 (compile-time consts there, runtime here)."""
    N: int = 1024
    Q: int = 536870909
    T: int = 7
    B: int = 19
    # RNS: one circuit per limb prime q_i proves c0 = pk0*u + delta_i*m + e0 (mod q_i) with delta_i = (Q_total // T) mod q_i,
    # Q_total the product of the limb primes (the reference has a single modulus and delta = Q // T, bfv.rs:112)
    delta_override: int = None

    @property
    def delta(self):
        if self.delta_override is not None:
            return self.delta_override
        return self.Q // self.T  # bfv.rs:112


INPUT_KEYS = ("pk0", "pk1", "m", "u", "e0", "e1", "c0", "c1", "cyclo")  # bfv.rs:50-61


def load_input(path):
    with open(path) as f:
        d = json.load(f)
    for k in INPUT_KEYS:
        _check(k in d, f"missing field {k}")
    return d


@dataclass
class Phase0:
    ctx: Context
    make_public: list
    polys: dict = field(default_factory=dict)   # name -> PolyChip
    delta_cell: object = None


def phase0(inp, params=BfvParams(), ctx=None):
    """bfv.rs:70-165."""
    N, Q = params.N, params.Q
    ctx = ctx or Context(0, 0)
    un = {k: Poly.from_string(inp[k], Q) for k in INPUT_KEYS}          # :71-79
    for k in INPUT_KEYS[:-1]:
        _check(un[k].deg() == N - 1, f"deg({k}) != N-1 (bfv.rs:82-89)")
    _check(un["cyclo"].deg() == N, "deg(cyclo) != N (bfv.rs:90)")

    st = Phase0(ctx, [])
    P = st.polys
    # :101-109 -- assignment order is the layout contract
    for name, key in (("pk0", "pk0"), ("pk1", "pk1"), ("m", "m"), ("u", "u"), ("e0", "e0"),
                      ("e1", "e1"), ("expected_c0", "c0"), ("expected_c1", "c1"), ("cyclo", "cyclo")):
        P[name] = PolyChip.from_poly(un[key], ctx)
    st.delta_cell = ctx.load_constant(params.delta)                      # :115
    for name in ("pk0", "pk1", "expected_c0", "expected_c1", "cyclo"):   # :118-122
        P[name].to_public(st.make_public)

    pk0_u = un["pk0"].mul(un["u"])                                       # :131-132
    pk1_u = un["pk1"].mul(un["u"])
    P["pk0_u"] = PolyChip.from_poly(pk0_u, ctx)                          # :135-136
    P["pk1_u"] = PolyChip.from_poly(pk1_u, ctx)
    pk0_u_red = pk0_u.reduce_by_modulus(Q)                               # :139-140
    pk1_u_red = pk1_u.reduce_by_modulus(Q)
    q0, r0 = pk0_u_red.divide_by_cyclo(un["cyclo"], Q)                   # :143-146
    q1, r1 = pk1_u_red.divide_by_cyclo(un["cyclo"], Q)
    q0c = q0.mul(un["cyclo"])                                            # :149-150
    q1c = q1.mul(un["cyclo"])
    P["quotient_0"] = PolyChip.from_poly(q0, ctx)                        # :156-157
    P["quotient_1"] = PolyChip.from_poly(q1, ctx)
    P["quotient_0_times_cyclo"] = PolyChip.from_poly(q0c, ctx)           # :160-161
    P["quotient_1_times_cyclo"] = PolyChip.from_poly(q1c, ctx)
    P["remainder_0"] = PolyChip.from_poly(r0, ctx)                       # :164-165
    P["remainder_1"] = PolyChip.from_poly(r1, ctx)
    return st


def phase1(st, gamma, params=BfvParams(), lookup_bits=8, ctx_gate=None, ctx_rlc=None):
    """The callback, bfv.rs:172-301.  Returns (ctx_gate, ctx_rlc)."""
    Q, T, B = params.Q, params.T, params.B
    P = st.polys
    ctx_gate = ctx_gate or Context(1, 1)
    ctx_rlc = ctx_rlc or Context(2, 1)
    range_ = RangeChip(lookup_bits)
    gate = range_.gate
    rlc = RlcChip(gamma)

    P["e0"].constrain_coefficients_in_range(ctx_gate, range_, B, Q)      # :189
    P["e1"].constrain_coefficients_in_range(ctx_gate, range_, B, Q)      # :190
    P["u"].constrain_from_distribution_chi_key(ctx_gate, gate, Q - 1)    # :201
    P["m"].constrain_coefficients_in_range(ctx_gate, range_, T // 2, Q)  # :210

    def half(pk, pk_u, quotient, qtc, remainder):
        P[pk].constrain_mul(P["u"].clone(), P[pk_u].clone(), ctx_gate, ctx_rlc, rlc)   # :215 / :264
        red = P[pk_u].reduce_by_modulo(ctx_gate, range_, Q)                              # :219 / :268
        P[quotient].constrain_coefficients_in_modulus_field(ctx_gate, range_, Q)         # :225 / :274
        P[remainder].constrain_coefficients_in_modulus_field(ctx_gate, range_, Q)        # :226 / :275
        return red.reduce_by_cyclo(P["cyclo"].clone(), P[quotient], P[qtc], P[remainder],
                                   range_, ctx_gate, ctx_rlc, rlc, Q)                     # :228 / :277

    pk0_u = half("pk0", "pk0_u", "quotient_0", "quotient_0_times_cyclo", "remainder_0")
    m_delta = P["m"].scalar_mul(ctx_gate, st.delta_cell, gate)           # :243
    t = pk0_u.add(ctx_gate, m_delta, gate)                               # :247
    c0 = t.add(ctx_gate, P["e0"], gate)                                  # :251
    c0 = c0.reduce_by_modulo(ctx_gate, range_, Q)                        # :255
    c0.constrain_equality(ctx_gate, P["expected_c0"], gate)              # :259

    pk1_u = half("pk1", "pk1_u", "quotient_1", "quotient_1_times_cyclo", "remainder_1")
    c1 = pk1_u.add(ctx_gate, P["e1"], gate)                              # :292
    c1 = c1.reduce_by_modulo(ctx_gate, range_, Q)                        # :296
    c1.constrain_equality(ctx_gate, P["expected_c1"], gate)              # :300
    st.computed = {"c0": c0, "c1": c1}
    return ctx_gate, ctx_rlc


# --- column layout (halo2-base GateThreadBuilder::assign_all) ----------------
def max_rows(k, unusable_rows):
    return (1 << k) - unusable_rows


def layout_gate_columns(advice, selector, mrows):
    """Cut one phase's flat advice vector into columns.  Returns
    (columns, selectors, break_points); each break duplicates the break cell at
    row 0 of the next column (copy-constrained upstream)."""
    cols, sels, bps = [[]], [[]], []
    row = 0
    for v, q in zip(advice, selector):
        cols[-1].append(v)
        sels[-1].append(False)
        if (q and row + 4 > mrows) or row >= mrows - 1:
            bps.append(row)
            row = 0
            cols.append([v])
            sels.append([False])
        if q:
            sels[-1][row] = True
        row += 1
    return cols, sels, bps


def layout_rlc_columns(advice, selector, mrows):
    """axiom-eth RlcCircuitBuilder::assign_rlc [UPSTREAM-RECALL]: as the gate
    layout but the RLC gate spans 3 rows."""
    cols, sels, bps = [[]], [[]], []
    row = 0
    for v, q in zip(advice, selector):
        cols[-1].append(v)
        sels[-1].append(False)
        if (q and row + 3 > mrows) or row >= mrows - 1:
            bps.append(row)
            row = 0
            cols.append([v])
            sels.append([False])
        if q:
            sels[-1][row] = True
        row += 1
    return cols, sels, bps


def layout_lookup_columns(cells, mrows):
    vals = [c.value for c in cells]
    return [vals[i:i + mrows] for i in range(0, len(vals), mrows)]


def build_tables(inp, gamma, params=BfvParams(), k=13, unusable_rows=109, lookup_bits=8):
    """Full advice table in reference layout.  Returns a dict with per-context
    flat vectors, the column cut and the pinning that `keygen` would write."""
    st = phase0(inp, params)
    ctx_gate, ctx_rlc = phase1(st, gamma, params, lookup_bits)
    mr = max_rows(k, unusable_rows)
    g0 = layout_gate_columns(st.ctx.advice, st.ctx.selector, mr)
    g1 = layout_gate_columns(ctx_gate.advice, ctx_gate.selector, mr)
    rl = layout_rlc_columns(ctx_rlc.advice, ctx_rlc.selector, mr)
    lookups = layout_lookup_columns(st.ctx.cells_to_lookup + ctx_gate.cells_to_lookup + ctx_rlc.cells_to_lookup, mr)
    pinning = {
        "params": {
            "degree": k,
            "num_rlc_columns": len(rl[0]),
            "num_range_advice": [len(g0[0]), len(g1[0]), 0],
            "num_lookup_advice": [0, len(lookups), 0],
            "num_fixed": 1,
            "unusable_rows": unusable_rows,
            "keccak_rows_per_round": 50,
            "lookup_bits": lookup_bits,
        },
        "break_points": {"gate": [g0[2], g1[2], []], "rlc": rl[2]},
    }
    return {
        "phase0": st, "ctx_gate": ctx_gate, "ctx_rlc": ctx_rlc,
        "gate0": g0, "gate1": g1, "rlc": rl, "lookup": lookups,
        "instances": [c.value for c in st.make_public],
        "pinning": pinning,
    }
