"""The BFV encryption circuit, host side (reference examples/bfv.rs:63-304), written
against the Poly / PolyChip mirrors so that it reads like the reference.  All arithmetic
runs on the GPU through the C ABI; the call ORDER below is the layout contract (it fixes
the order of cells in the advice table).
"""
import json
from dataclasses import dataclass

from .capi import ZkfheError
from .poly import Poly
from .poly_chip import CTX_GATE, CTX_PHASE0, CTX_RLC, PolyChip, Witness

INPUT_KEYS = ("pk0", "pk1", "m", "u", "e0", "e1", "c0", "c1", "cyclo")    # bfv.rs:50-61


@dataclass
class BfvParams:
    """bfv.rs:27-30 (compile-time consts in the reference, runtime parameters here)."""
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
        return self.Q // self.T        # bfv.rs:112


def load_input(path):
    """`CircuitInput` (bfv.rs:50-61): nine arrays of decimal strings."""
    with open(path) as f:
        d = json.load(f)
    for k in INPUT_KEYS:
        if k not in d:
            raise ZkfheError(-2, f"input file misses field `{k}`")
    return d


def _assert_eq(a, b, what):
    if a != b:
        raise ZkfheError(-4, f"assertion failed: `(left == right)` {what}: left: {a}, right: {b}")


class BfvCircuit:
    """Two-phase witness generation: `phase0` (bfv.rs:70-165), then -- once the phase-0
    commitment has produced gamma -- `phase1`, the callback (bfv.rs:172-301)."""

    def __init__(self, ctx, params=BfvParams(), lookup_bits=8, record=False):
        self.ctx = ctx
        self.params = params
        self.wit = Witness(ctx, lookup_bits, record=record)
        self.P = {}
        self.delta = None

    def upload(self, inp):
        """The nine `Poly::from_string` calls (bfv.rs:71-79) on their own, for callers that keep
        inputs resident in HBM between proofs."""
        return {k: Poly.from_string(self.ctx, inp[k], self.params.Q) for k in INPUT_KEYS}

    def phase0(self, inp, resident=None):
        N, Q = self.params.N, self.params.Q
        ctx, w = self.ctx, self.wit
        un = resident if resident is not None else self.upload(inp)                   # :71-79
        for k in INPUT_KEYS[:-1]:
            _assert_eq(un[k].deg(), N - 1, f"deg({k}) (examples/bfv.rs:82-89)")
        _assert_eq(un["cyclo"].deg(), N, "deg(cyclo) (examples/bfv.rs:90)")
        P = self.P
        for name, key in (("pk0", "pk0"), ("pk1", "pk1"), ("m", "m"), ("u", "u"), ("e0", "e0"), ("e1", "e1"),
                          ("expected_c0", "c0"), ("expected_c1", "c1"), ("cyclo", "cyclo")):   # :101-109
            P[name] = PolyChip.from_poly(un[key], w, CTX_PHASE0)
        self.delta = w.load_constant(CTX_PHASE0, self.params.delta)                      # :115
        for name in ("pk0", "pk1", "expected_c0", "expected_c1", "cyclo"):               # :118-122
            P[name].to_public()
        pk0_u = un["pk0"].mul(un["u"])                                                  # :131-132
        pk1_u = un["pk1"].mul(un["u"])
        P["pk0_u"] = PolyChip.from_poly(pk0_u, w, CTX_PHASE0)                            # :135-136
        P["pk1_u"] = PolyChip.from_poly(pk1_u, w, CTX_PHASE0)
        pk0_u_red = pk0_u.reduce_by_modulus(Q)                                          # :139-140
        pk1_u_red = pk1_u.reduce_by_modulus(Q)
        q0, r0 = pk0_u_red.divide_by_cyclo(un["cyclo"], Q, check=False)                 # :143-146
        q1, r1 = pk1_u_red.divide_by_cyclo(un["cyclo"], Q, check=False)
        q0c = q0.mul(un["cyclo"])                                                       # :149-150
        q1c = q1.mul(un["cyclo"])
        P["quotient_0"] = PolyChip.from_poly(q0, w, CTX_PHASE0)                          # :156-157
        P["quotient_1"] = PolyChip.from_poly(q1, w, CTX_PHASE0)
        P["quotient_0_times_cyclo"] = PolyChip.from_poly(q0c, w, CTX_PHASE0)             # :160-161
        P["quotient_1_times_cyclo"] = PolyChip.from_poly(q1c, w, CTX_PHASE0)
        P["remainder_0"] = PolyChip.from_poly(r0, w, CTX_PHASE0)                         # :164-165
        P["remainder_1"] = PolyChip.from_poly(r1, w, CTX_PHASE0)
        ctx.status()        # surface the data-dependent reference asserts of phase 0 (one sync)
        return self

    def phase1(self, gamma):
        Q, T, B = self.params.Q, self.params.T, self.params.B
        P, w = self.P, self.wit
        w.set_challenge(gamma)
        P["e0"].constrain_coefficients_in_range(B, Q)                                   # :189
        P["e1"].constrain_coefficients_in_range(B, Q)                                   # :190
        P["u"].constrain_from_distribution_chi_key(Q - 1)                               # :201
        P["m"].constrain_coefficients_in_range(T // 2, Q)                               # :210

        def half(pk, pk_u, quotient, qtc, remainder):
            P[pk].constrain_mul(P["u"].clone(), P[pk_u].clone())                        # :215 / :264
            red = P[pk_u].reduce_by_modulo(Q)                                           # :219 / :268
            P[quotient].constrain_coefficients_in_modulus_field(Q)                      # :225 / :274
            P[remainder].constrain_coefficients_in_modulus_field(Q)                     # :226 / :275
            return red.reduce_by_cyclo(P["cyclo"].clone(), P[quotient], P[qtc], P[remainder], Q)   # :228 / :277

        pk0_u = half("pk0", "pk0_u", "quotient_0", "quotient_0_times_cyclo", "remainder_0")
        m_delta = P["m"].scalar_mul(self.delta)                                         # :243
        c0 = pk0_u.add(m_delta).add(P["e0"])                                            # :247, :251
        c0 = c0.reduce_by_modulo(Q)                                                     # :255
        c0.constrain_equality(P["expected_c0"])                                         # :259
        pk1_u = half("pk1", "pk1_u", "quotient_1", "quotient_1_times_cyclo", "remainder_1")
        c1 = pk1_u.add(P["e1"])                                                         # :292
        c1 = c1.reduce_by_modulo(Q)                                                     # :296
        c1.constrain_equality(P["expected_c1"])                                         # :300
        self.computed = {"c0": c0, "c1": c1}
        return self


__all__ = ["BfvParams", "BfvCircuit", "load_input", "CTX_PHASE0", "CTX_GATE", "CTX_RLC"]
