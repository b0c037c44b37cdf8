"""Host-side mirror of `zk_fhe::poly_chip::PolyChip<F>` (reference src/poly_chip.rs) and of
the halo2-base `Context`s it writes to, for the test harness.  A `Witness` owns the flat
advice vectors in HBM (context 0: phase-0 gate, 1: phase-1 gate, 2: phase-1 RLC); every
PolyChip method is one call through the C ABI (include/zkfhe_b200.h, stage 1b) that assigns
exactly the cells the CPU builder assigns, in the same order.
"""
import ctypes

import numpy as np

from .capi import AssignedPoly, Cell, _addr

CTX_PHASE0, CTX_GATE, CTX_RLC = 0, 1, 2


class Witness:
    def __init__(self, ctx, lookup_bits=8, record=False):
        """record=True: also record selectors / copy constraints / constants (keygen, mock)."""
        self.ctx = ctx
        self.lookup_bits = lookup_bits
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.zkfhe_witness_new(ctx.h, lookup_bits, ctypes.byref(h)))
        self.h = h
        self.record = record
        if record:
            ctx._check(ctx.lib.zkfhe_witness_set_recording(h, 1))

    def structure(self, ctx_id):
        """(flags uint8[n], copy uint64[n]) of one context (recording mode)."""
        n = self.counts()["advice"][ctx_id]
        flags = np.zeros(n, dtype=np.uint8)
        copy = np.zeros(n, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.zkfhe_witness_download_structure(self.h, ctx_id, _addr(flags), _addr(copy)))
        return flags, copy

    def lookup_sources(self):
        src = np.zeros(self.counts()["lookups"], dtype=np.uint64)
        if len(src):
            self.ctx._check(self.ctx.lib.zkfhe_witness_download_lookup_sources(self.h, _addr(src)))
        return src

    def public_cells(self):
        ids = np.zeros(self.counts()["instances"], dtype=np.uint64)
        if len(ids):
            self.ctx._check(self.ctx.lib.zkfhe_witness_public_cells(self.h, _addr(ids)))
        return ids

    def mock(self):
        """The `mock` subcommand: raises ZkfheError(ZKFHE_ERR_UNSATISFIED) on any violated constraint."""
        n, first = ctypes.c_uint64(), ctypes.c_uint64()
        self.ctx._check(self.ctx.lib.zkfhe_witness_mock(self.h, ctypes.byref(n), ctypes.byref(first)))
        return int(n.value)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.zkfhe_witness_free(self.h)
        except Exception:
            pass
        self.h = None

    def reset(self):
        self.ctx._check(self.ctx.lib.zkfhe_witness_reset(self.h))

    def load_constant(self, ctx_id, value):
        c = Cell()
        self.ctx._check(self.ctx.lib.zkfhe_chip_load_constant(self.h, ctx_id, value, ctypes.byref(c)))
        c.value = value
        return c

    def set_challenge(self, gamma):
        """gamma: canonical int; stored as Fr Montgomery bytes."""
        from .capi import fr_mont_bytes
        buf = fr_mont_bytes(gamma)      # keep the buffer alive across the call
        self.ctx._check(self.ctx.lib.zkfhe_chip_set_challenge(self.h, _addr(buf)))

    def counts(self):
        adv = (ctypes.c_uint64 * 3)()
        lk, inst = ctypes.c_uint64(), ctypes.c_uint64()
        self.ctx._check(self.ctx.lib.zkfhe_witness_counts(self.h, adv, ctypes.byref(lk), ctypes.byref(inst)))
        return {"advice": [int(x) for x in adv], "lookups": int(lk.value), "instances": int(inst.value)}

    def download(self, which):
        """(cells, 4) uint64 Fr Montgomery limbs; which: 0..2 advice, 3 lookup cells, 4 instances."""
        c = self.counts()
        n = c["advice"][which] if which <= 2 else (c["lookups"] if which == 3 else c["instances"])
        out = np.zeros((n, 4), dtype=np.uint64)
        if n:
            self.ctx._check(self.ctx.lib.zkfhe_witness_download(self.h, which, _addr(out)))
        return out

    def status(self):
        self.ctx.status()


class PolyChip:
    """poly_chip.rs:19-23: { assigned_coefficients, max_num_bits, degree }."""

    def __init__(self, wit, ap):
        self.w = wit
        self.ap = ap

    @property
    def max_num_bits(self):
        return int(self.ap.max_num_bits)

    @property
    def degree(self):
        return int(self.ap.len) - 1

    def clone(self):
        return PolyChip(self.w, self.ap)

    def _call(self, fn, *args):
        self.w.ctx._check(fn(self.w.h, *args))

    @classmethod
    def from_poly(cls, poly, wit, ctx_id=CTX_PHASE0):
        ap = AssignedPoly()
        wit.ctx._check(wit.ctx.lib.zkfhe_chip_from_poly(wit.h, ctx_id, poly.h, ctypes.byref(ap)))
        return cls(wit, ap)

    def to_public(self):
        self._call(self.w.ctx.lib.zkfhe_chip_to_public, ctypes.byref(self.ap))

    def constrain_mul(self, b, c, ctx_gate=CTX_GATE, ctx_rlc=CTX_RLC):
        self._call(self.w.ctx.lib.zkfhe_chip_constrain_mul, ctx_gate, ctx_rlc, ctypes.byref(self.ap),
                   ctypes.byref(b.ap), ctypes.byref(c.ap))

    def add(self, other, ctx_gate=CTX_GATE):
        out = AssignedPoly()
        self._call(self.w.ctx.lib.zkfhe_chip_add, ctx_gate, ctypes.byref(self.ap), ctypes.byref(other.ap), ctypes.byref(out))
        return PolyChip(self.w, out)

    def scalar_mul(self, scalar, ctx_gate=CTX_GATE):
        out = AssignedPoly()
        self._call(self.w.ctx.lib.zkfhe_chip_scalar_mul, ctx_gate, ctypes.byref(self.ap), ctypes.byref(scalar),
                   scalar.value, ctypes.byref(out))
        return PolyChip(self.w, out)

    def reduce_by_cyclo(self, cyclo, quotient, quotient_times_cyclo, remainder, modulus,
                        ctx_gate=CTX_GATE, ctx_rlc=CTX_RLC):
        out = AssignedPoly()
        self._call(self.w.ctx.lib.zkfhe_chip_reduce_by_cyclo, ctx_gate, ctx_rlc, ctypes.byref(self.ap),
                   ctypes.byref(cyclo.ap), ctypes.byref(quotient.ap), ctypes.byref(quotient_times_cyclo.ap),
                   ctypes.byref(remainder.ap), modulus, ctypes.byref(out))
        return PolyChip(self.w, out)

    def reduce_by_modulo(self, modulus, ctx_gate=CTX_GATE):
        out = AssignedPoly()
        self._call(self.w.ctx.lib.zkfhe_chip_reduce_by_modulo, ctx_gate, ctypes.byref(self.ap), modulus, ctypes.byref(out))
        return PolyChip(self.w, out)

    def constrain_equality(self, other, ctx_gate=CTX_GATE):
        self._call(self.w.ctx.lib.zkfhe_chip_constrain_equality, ctx_gate, ctypes.byref(self.ap), ctypes.byref(other.ap))

    def constrain_coefficients_in_range(self, z, y, ctx_gate=CTX_GATE):
        self._call(self.w.ctx.lib.zkfhe_chip_constrain_coefficients_in_range, ctx_gate, ctypes.byref(self.ap), z, y)

    def constrain_from_distribution_chi_key(self, z, ctx_gate=CTX_GATE):
        self._call(self.w.ctx.lib.zkfhe_chip_constrain_from_distribution_chi_key, ctx_gate, ctypes.byref(self.ap), z)

    def constrain_coefficients_in_modulus_field(self, modulus, ctx_gate=CTX_GATE):
        self._call(self.w.ctx.lib.zkfhe_chip_constrain_coefficients_in_modulus_field, ctx_gate, ctypes.byref(self.ap), modulus)

    def safe_trim_leading_zeroes(self, degree):
        out = AssignedPoly()
        self._call(self.w.ctx.lib.zkfhe_chip_safe_trim_leading_zeroes, ctypes.byref(self.ap), degree, ctypes.byref(out))
        return PolyChip(self.w, out)
