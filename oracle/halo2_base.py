"""Cell-level witness semantics of halo2-base v0.3.0-ce and axiom-eth's RlcChip
(oracle; test-only).

[UPSTREAM-RECALL] The crates are un-vendored (Cargo.toml:9-11 of the
reference); this file restates their published behaviour as summarised in
SURVEY.md Appendix B.  The restatement is PINNED on structure by
/root/reference/configs/bfv.json: column counts and every break point are
reproduced exactly (tests/test_oracle_layout.py).

Vertical gate:  q * (a + b*c - d) = 0  on 4 consecutive cells of one column.
RLC gate:       q * (a*gamma + b - c) = 0 on 3 consecutive cells.
Values are canonical ints mod r.
"""
from collections import namedtuple

from .field import R_MOD, inv

# A cell is (context_id, offset); an AssignedValue carries its value too.
AssignedValue = namedtuple("AssignedValue", "value ctx offset")


class Existing(namedtuple("Existing", "av")):
    @property
    def value(self):
        return self.av.value


class Witness(namedtuple("Witness", "value")):
    pass


class Constant(namedtuple("Constant", "value")):
    pass


def _q(x):
    return Existing(x) if isinstance(x, AssignedValue) else x


class Context:
    """halo2_base::Context (witness_gen_only = False: records selectors, copy
    constraints and constant constraints as well as values)."""

    def __init__(self, context_id, phase):
        self.context_id = context_id
        self.phase = phase
        self.advice = []            # ints mod r
        self.selector = []          # bools, same length as advice
        self.cells_to_lookup = []   # AssignedValue, creation order
        self.advice_equality = []   # ((ctx,off),(ctx,off))
        self.constant_equality = []  # (constant, (ctx,off))
        self._zero = None

    # -- primitive assignment --------------------------------------------
    def _assign_cell(self, qc):
        off = len(self.advice)
        if isinstance(qc, Existing):
            self.advice.append(qc.av.value % R_MOD)
            self.advice_equality.append(((self.context_id, off), (qc.av.ctx, qc.av.offset)))
        elif isinstance(qc, Witness):
            self.advice.append(qc.value % R_MOD)
        elif isinstance(qc, Constant):
            c = qc.value % R_MOD
            self.advice.append(c)
            self.constant_equality.append((c, (self.context_id, off)))
        else:
            raise TypeError(qc)

    def get(self, offset):
        if offset < 0:
            offset += len(self.advice)
        return AssignedValue(self.advice[offset], self.context_id, offset)

    def last(self):
        return self.get(-1)

    def assign_region(self, inputs, gate_offsets, equality_offsets=()):
        row = len(self.advice)
        for qc in inputs:
            self._assign_cell(_q(qc))
        self.selector.extend([False] * (len(self.advice) - len(self.selector)))
        for g in gate_offsets:
            self.selector[row + g] = True
        for (o1, o2) in equality_offsets:
            self.advice_equality.append(((self.context_id, row + o1), (self.context_id, row + o2)))
        return row

    def assign_region_last(self, inputs, gate_offsets):
        self.assign_region(inputs, gate_offsets)
        return self.last()

    def load_witness(self, v):
        self._assign_cell(Witness(v))
        self.selector.append(False)
        return self.last()

    def load_constant(self, c):
        self._assign_cell(Constant(c))
        self.selector.append(False)
        return self.last()

    def load_zero(self):
        if self._zero is None:
            self._zero = self.load_constant(0)
        return self._zero

    def constrain_equal(self, a, b):
        self.advice_equality.append(((a.ctx, a.offset), (b.ctx, b.offset)))


def _val(x):
    return x.value % R_MOD


class GateChip:
    """halo2_base::gates::GateChip, Vertical strategy."""

    def add(self, ctx, a, b):
        a, b = _q(a), _q(b)
        out = (_val(a) + _val(b)) % R_MOD
        return ctx.assign_region_last([a, b, Constant(1), Witness(out)], [0])

    def sub(self, ctx, a, b):
        a, b = _q(a), _q(b)
        out = (_val(a) - _val(b)) % R_MOD
        ctx.assign_region([Witness(out), b, Constant(1), a], [0])
        return ctx.get(-4)

    def mul(self, ctx, a, b):
        a, b = _q(a), _q(b)
        out = _val(a) * _val(b) % R_MOD
        return ctx.assign_region_last([Constant(0), a, b, Witness(out)], [0])

    def not_(self, ctx, a):
        return self.sub(ctx, Constant(1), a)

    def or_(self, ctx, a, b):
        a, b = _q(a), _q(b)
        not_b = (1 - _val(b)) % R_MOD
        out = (_val(a) + _val(b) - _val(a) * _val(b)) % R_MOD
        cells = [Witness(not_b), Constant(1), b, Constant(1), b, a, Witness(not_b), Witness(out)]
        ctx.assign_region(cells, [0, 4], [(0, 6), (2, 4)])
        return ctx.last()

    def assert_bit(self, ctx, x):
        ctx.assign_region([Constant(0), Existing(x), Existing(x), Existing(x)], [0])

    def is_zero(self, ctx, a):
        x = a.value % R_MOD
        if x == 0:
            is_zero, inv_v = 1, 1      # Assigned::Trivial(F::one())
        else:
            is_zero, inv_v = 0, inv(x)  # Assigned::Rational(1, x), batch-inverted at assignment
        cells = [Witness(is_zero), Existing(a), Witness(inv_v), Constant(1),
                 Constant(0), Existing(a), Witness(is_zero), Constant(0)]
        ctx.assign_region(cells, [0, 4], [(0, 6)])
        return ctx.get(-2)

    def is_equal(self, ctx, a, b):
        diff = self.sub(ctx, a, b)
        return self.is_zero(ctx, diff)

    def assert_is_const(self, ctx, a, c):
        ctx.constant_equality.append((c % R_MOD, (a.ctx, a.offset)))

    def inner_product(self, ctx, a_vals, b_consts):
        """inner_product with witness `a` and constant `b` (first b == 1 fast
        path): cells [a0, a1, b1, acc1, a2, b2, acc2, ...]."""
        a_vals = list(a_vals)
        b_consts = list(b_consts)
        assert len(a_vals) == len(b_consts)
        cells = []
        if b_consts[0] % R_MOD == 1:
            s = _val(_q(a_vals[0]))
            cells.append(_q(a_vals[0]))
            rest = zip(a_vals[1:], b_consts[1:])
        else:
            s = 0
            cells.append(Constant(0))
            rest = zip(a_vals, b_consts)
        for a, b in rest:
            a = _q(a)
            s = (s + _val(a) * b) % R_MOD
            cells += [a, Constant(b), Witness(s)]
        n_gates = len(cells) // 3
        ctx.assign_region(cells, [3 * i for i in range(n_gates)])
        return ctx.last()


def bit_length(x):
    return int(x).bit_length()


class RangeChip:
    """halo2_base::gates::RangeChip, Vertical strategy, `lookup_bits` table."""

    def __init__(self, lookup_bits):
        self.lookup_bits = lookup_bits
        self.gate = GateChip()

    def gate_(self):
        return self.gate

    def range_check(self, ctx, a, range_bits):
        lb = self.lookup_bits
        k = (range_bits + lb - 1) // lb
        rem_bits = range_bits % lb
        if k == 1:
            ctx.cells_to_lookup.append(a)
        else:
            v = a.value % R_MOD
            limbs = [Witness((v >> (lb * i)) & ((1 << lb) - 1)) for i in range(k)]
            row = len(ctx.advice)
            acc = self.gate.inner_product(ctx, limbs, [1 << (lb * i) for i in range(k)])
            ctx.constrain_equal(a, acc)
            ctx.cells_to_lookup.append(ctx.get(row))
            for i in range(k - 1):
                ctx.cells_to_lookup.append(ctx.get(row + 1 + 3 * i))
        if rem_bits == 1:
            self.gate.assert_bit(ctx, ctx.cells_to_lookup[-1])
        elif rem_bits > 1:
            check = self.gate.mul(ctx, ctx.cells_to_lookup[-1], Constant(1 << (lb - rem_bits)))
            ctx.cells_to_lookup.append(check)

    def check_less_than(self, ctx, a, b, num_bits):
        a, b = _q(a), _q(b)
        pow2 = 1 << num_bits
        shift_a = (pow2 + _val(a)) % R_MOD
        cells = [Witness((shift_a - _val(b)) % R_MOD), b, Constant(1), Witness(shift_a),
                 Constant((-pow2) % R_MOD), Constant(1), a]
        ctx.assign_region(cells, [0, 3])
        self.range_check(ctx, ctx.get(-7), num_bits)

    def check_less_than_safe(self, ctx, a, b):
        lb = self.lookup_bits
        range_bits = (bit_length(b) + lb - 1) // lb * lb
        self.range_check(ctx, a, range_bits)
        self.check_less_than(ctx, a, Constant(b), range_bits)

    check_big_less_than_safe = check_less_than_safe

    def is_less_than(self, ctx, a, b, num_bits):
        a, b = _q(a), _q(b)
        lb = self.lookup_bits
        k = (num_bits + lb - 1) // lb
        padded_bits = k * lb
        pow_padded = 1 << padded_bits
        shift_a = (pow_padded + _val(a)) % R_MOD
        shifted = (shift_a - _val(b)) % R_MOD
        ctx.assign_region([Witness(shifted), b, Constant(1), Witness(shift_a),
                           Constant((-pow_padded) % R_MOD), Constant(1), a], [0, 3])
        self.range_check(ctx, ctx.get(-7), padded_bits + lb)
        return self.gate.is_zero(ctx, ctx.cells_to_lookup[-1])

    def div_mod(self, ctx, a, b, a_num_bits):
        a = _q(a)
        a_val = _val(a)
        div, rem = divmod(a_val, b)
        ctx.assign_region([Witness(rem), Constant(b), Witness(div), a], [0])
        rem_c = ctx.get(-4)
        div_c = ctx.get(-2)
        self.check_big_less_than_safe(ctx, div_c, (1 << a_num_bits) // b + 1)
        self.check_big_less_than_safe(ctx, rem_c, b)
        return div_c, rem_c


class RlcChip:
    """axiom_eth::rlp::rlc::RlcChip::compute_rlc_fixed_len [UPSTREAM-RECALL]:
    Horner evaluation in gamma, first input is the highest power."""

    def __init__(self, gamma):
        self.gamma = gamma % R_MOD

    def compute_rlc_fixed_len(self, ctx_rlc, inputs):
        inputs = list(inputs)
        assert inputs, "empty RLC"
        running = inputs[0].value % R_MOD
        cells = [Existing(inputs[0])]
        for x in inputs[1:]:
            running = (running * self.gamma + x.value) % R_MOD
            cells += [Existing(x), Witness(running)]
        ctx_rlc.assign_region(cells, [2 * i for i in range(len(inputs) - 1)])
        return ctx_rlc.last()
