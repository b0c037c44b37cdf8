"""In-circuit polynomial chip (oracle; test-only).

Restates /root/reference/src/poly_chip.rs method by method on top of the
cell-level semantics in oracle/halo2_base.py.
"""
from .field import R_MOD
from .halo2_base import Constant, Existing
from .poly import OracleError, _check

P_BITS = R_MOD.bit_length()  # 254  (poly_chip.rs:90-91)


def bigint_to_fe(x):
    """halo2_base::utils::bigint_to_fe: negative -> -(|x|) mod r."""
    return x % R_MOD


class PolyChip:
    """poly_chip.rs:19-23."""

    def __init__(self, assigned_coefficients, max_num_bits):
        _check(len(assigned_coefficients) >= 1, "empty PolyChip (poly_chip.rs:49)")
        self.assigned_coefficients = list(assigned_coefficients)
        self.max_num_bits = max_num_bits
        self.degree = len(assigned_coefficients) - 1

    @classmethod
    def from_poly(cls, poly, ctx):
        # poly_chip.rs:27-42
        cells = [ctx.load_witness(bigint_to_fe(c)) for c in poly.coefficients[:poly.deg() + 1]]
        return cls(cells, poly.max_bits)

    def clone(self):
        return PolyChip(self.assigned_coefficients, self.max_num_bits)

    def to_public(self, make_public):
        # poly_chip.rs:58-62
        make_public.extend(self.assigned_coefficients)

    def constrain_mul(self, b, c, ctx_gate, ctx_rlc, rlc):
        # poly_chip.rs:81-116
        _check(c.max_num_bits < P_BITS, "overflow risk in constrain_mul (poly_chip.rs:94)")
        a_eval = rlc.compute_rlc_fixed_len(ctx_rlc, self.assigned_coefficients)
        b_eval = rlc.compute_rlc_fixed_len(ctx_rlc, b.assigned_coefficients)
        c_eval = rlc.compute_rlc_fixed_len(ctx_rlc, c.assigned_coefficients)
        ctx_gate.assign_region([Constant(0), Existing(a_eval), Existing(b_eval), Existing(c_eval)], [0])

    def add(self, ctx, other, gate):
        # poly_chip.rs:122-144
        out = [gate.add(ctx, self.assigned_coefficients[i], other.assigned_coefficients[i])
               for i in range(self.degree + 1)]
        mb = max(self.max_num_bits, other.max_num_bits) + 1
        _check(mb < P_BITS, "Risk of overflow detected in add")
        return PolyChip(out, mb)

    def scalar_mul(self, ctx, scalar, gate):
        # poly_chip.rs:150-174
        mb = self.max_num_bits + (scalar.value % R_MOD).bit_length()
        _check(mb < P_BITS, "Risk of overflow detected in scalar_mul")
        out = [gate.mul(ctx, c, scalar) for c in self.assigned_coefficients[:self.degree + 1]]
        return PolyChip(out, mb)

    def reduce_by_cyclo(self, cyclo, quotient, quotient_times_cyclo, remainder,
                        range_, ctx_gate, ctx_rlc, rlc, modulus):
        # poly_chip.rs:183-223
        mbits = modulus.bit_length()
        _check(quotient.max_num_bits <= mbits, "quotient.max_num_bits (poly_chip.rs:196)")
        _check(remainder.max_num_bits <= mbits, "remainder.max_num_bits (poly_chip.rs:197)")
        _check(max(quotient_times_cyclo.max_num_bits, remainder.max_num_bits) + 1 < P_BITS,
               "overflow risk (poly_chip.rs:201)")
        cyclo_deg = cyclo.degree
        quotient.constrain_mul(cyclo, quotient_times_cyclo.clone(), ctx_gate, ctx_rlc, rlc)
        s = quotient_times_cyclo.add(ctx_gate, remainder.clone(), range_.gate)
        s_mod = s.reduce_by_modulo(ctx_gate, range_, modulus)
        s_trim = s_mod.safe_trim_leading_zeroes(ctx_gate, range_, self.degree)
        s_trim.constrain_equality(ctx_gate, self.clone(), range_.gate)
        return remainder.safe_trim_leading_zeroes(ctx_gate, range_, cyclo_deg - 1)

    def reduce_by_modulo(self, ctx, range_, modulus):
        # poly_chip.rs:226-252
        nb = self.max_num_bits
        out = [range_.div_mod(ctx, self.assigned_coefficients[i], modulus, nb)[1]
               for i in range(self.degree + 1)]
        return PolyChip(out, modulus.bit_length())

    def constrain_equality(self, ctx, other, gate):
        # poly_chip.rs:255-264
        for i in range(self.degree + 1):
            b = gate.is_equal(ctx, self.assigned_coefficients[i], other.assigned_coefficients[i])
            gate.assert_is_const(ctx, b, 1)

    def constrain_coefficients_in_range(self, ctx, range_, z, y):
        # poly_chip.rs:270-317
        _check(z < y, "z < y (poly_chip.rs:278)")
        y_bits = y.bit_length()
        for coeff in self.assigned_coefficients:
            range_.check_less_than_safe(ctx, coeff, y)
            in1 = range_.is_less_than(ctx, coeff, Constant(z + 1), y_bits)
            not_in2 = range_.is_less_than(ctx, coeff, Constant(y - z), y_bits)
            in2 = range_.gate.not_(ctx, not_in2)
            in_range = range_.gate.or_(ctx, in1, in2)
            range_.gate.assert_is_const(ctx, in_range, 1)

    def constrain_from_distribution_chi_key(self, ctx, gate, z):
        # poly_chip.rs:320-354
        for coeff in self.assigned_coefficients:
            f1 = gate.sub(ctx, coeff, Constant(0))
            f2 = gate.sub(ctx, coeff, Constant(1))
            f3 = gate.sub(ctx, coeff, Constant(z))
            f12 = gate.mul(ctx, f1, f2)
            f123 = gate.mul(ctx, f12, f3)
            gate.assert_is_const(ctx, f123, 0)

    def constrain_coefficients_in_modulus_field(self, ctx, range_, modulus):
        # poly_chip.rs:357-366
        for coeff in self.assigned_coefficients:
            range_.check_less_than_safe(ctx, coeff, modulus)

    def safe_trim_leading_zeroes(self, ctx, range_, degree):
        # poly_chip.rs:374-399
        _check(degree <= self.degree, "degree <= self.degree (poly_chip.rs:380)")
        for i in range(self.degree - degree):
            range_.gate.assert_is_const(ctx, self.assigned_coefficients[i], 0)
        return PolyChip(self.assigned_coefficients[self.degree - degree:], self.max_num_bits)


__all__ = ["PolyChip", "OracleError", "bigint_to_fe"]
