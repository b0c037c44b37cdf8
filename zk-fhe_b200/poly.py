"""Host-side mirror of `zk_fhe::poly::Poly` (reference src/poly.rs:9-13) for the test
harness: same method names, argument meaning and error behaviour, but the
coefficients live in HBM and every method is one call through the C ABI
(include/zkfhe_b200.h, stage 1a).  Reference `assert!`s surface as ZkfheError.
"""
import ctypes

import numpy as np

from .capi import ZkfheError, _addr


class Poly:
    """Device-resident polynomial; big-endian coefficients (index 0 = highest degree)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = handle

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.zkfhe_poly_free(self.h)
        except Exception:
            pass
        self.h = None

    # -- constructors --------------------------------------------------------------------
    @classmethod
    def from_string(cls, ctx, coefficients, modulus):
        """Poly::from_string (poly.rs:21-40): decimal strings, each <= modulus."""
        try:
            text = ",".join(coefficients).encode("ascii")      # parsed in C: one pass instead of len(coefficients) int() calls
        except (TypeError, UnicodeEncodeError):
            raise ZkfheError(-2, "coefficients must be decimal strings")
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.zkfhe_poly_from_decimal(ctx.h, text, len(text), len(coefficients), modulus, ctypes.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_big_int(cls, ctx, coefficients, max_bits):
        """Poly::from_big_int (poly.rs:47-59)."""
        buf = np.frombuffer(b"".join(int(c).to_bytes(32, "little") for c in coefficients), dtype=np.uint64).copy()
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.zkfhe_poly_from_u256(ctx.h, _addr(buf), len(coefficients), max_bits, ctypes.byref(h)))
        return cls(ctx, h)

    # -- accessors -----------------------------------------------------------------------
    def __len__(self):
        return int(self.ctx.lib.zkfhe_poly_len(self.h))

    def deg(self):
        return len(self) - 1

    @property
    def degree(self):
        return self.deg()

    @property
    def max_bits(self):
        return int(self.ctx.lib.zkfhe_poly_max_bits(self.h))

    @property
    def coefficients(self):
        """Download as Python ints (test / debugging path)."""
        n = len(self)
        out = np.zeros(4 * n, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.zkfhe_poly_download(self.ctx.h, self.h, _addr(out)))
        b = out.tobytes()
        return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]

    # -- arithmetic ------------------------------------------------------------------------
    def mul(self, other):
        """Poly::mul (poly.rs:75-103)."""
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx.lib.zkfhe_poly_mul(self.ctx.h, self.h, other.h, ctypes.byref(h)))
        return Poly(self.ctx, h)

    def reduce_by_modulus(self, modulus):
        """Poly::reduce_by_modulus (poly.rs:180-191)."""
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx.lib.zkfhe_poly_reduce_by_modulus(self.ctx.h, self.h, modulus, ctypes.byref(h)))
        return Poly(self.ctx, h)

    def divide_by_cyclo(self, cyclo, modulus, check=True):
        """Poly::divide_by_cyclo (poly.rs:113-177).  With check=True the data-dependent
        reference panics are raised here (one stream synchronisation)."""
        q, r = ctypes.c_void_p(), ctypes.c_void_p()
        self.ctx._check(self.ctx.lib.zkfhe_poly_divide_by_cyclo(self.ctx.h, self.h, cyclo.h, modulus,
                                                                ctypes.byref(q), ctypes.byref(r)))
        out = Poly(self.ctx, q), Poly(self.ctx, r)
        if check:
            self.ctx.status()
        return out
