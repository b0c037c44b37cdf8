"""Off-circuit polynomial arithmetic (oracle; test-only).

Restates /root/reference/src/poly.rs line by line on Python ints.
Coefficients are big-endian: index 0 is the highest degree (poly.rs:17,43).
Rust `assert!`/panics become `OracleError`.
"""


class OracleError(AssertionError):
    pass


def _check(cond, msg):
    if not cond:
        raise OracleError(msg)


def log2_ceil(x):
    """halo2_base::utils::log2_ceil [UPSTREAM-RECALL]: ceil(log2 x) for x>=1
    computed as 64 - leading_zeros(x) - (x is a power of two)."""
    _check(x > 0, "log2_ceil(0)")
    return x.bit_length() - (1 if x & (x - 1) == 0 else 0)


class Poly:
    """poly.rs:9-13."""

    def __init__(self, coefficients, max_bits):
        # from_big_int, poly.rs:47-59
        _check(len(coefficients) >= 1, "empty polynomial (usize underflow at poly.rs:48)")
        for c in coefficients:
            _check(abs(c).bit_length() <= max_bits, "coefficient exceeds max_bits (poly.rs:51)")
        self.coefficients = list(coefficients)
        self.degree = len(coefficients) - 1
        self.max_bits = max_bits

    @classmethod
    def from_string(cls, coefficients, modulus):
        # poly.rs:21-40
        out = []
        for s in coefficients:
            c = int(s, 10)
            _check(c <= modulus, "coefficient > modulus (poly.rs:28)")
            out.append(c)
        _check(len(out) >= 1, "empty polynomial (usize underflow at poly.rs:33)")
        p = cls.__new__(cls)
        p.coefficients = out
        p.degree = len(out) - 1
        p.max_bits = modulus.bit_length()
        return p

    def deg(self):
        return self.degree

    def mul(self, other):
        # poly.rs:75-103 -- exact integer schoolbook product, no modular reduction
        _check(self.deg() == other.deg(), "degree mismatch (poly.rs:78)")
        da = self.deg()
        c = [0] * (2 * da + 1)
        a, b = self.coefficients, other.coefficients
        for i in range(da + 1):
            ai = a[i]
            if ai == 0:
                continue
            for j in range(da + 1):
                c[i + j] += ai * b[j]
        max_bits = self.max_bits + other.max_bits + log2_ceil(da + 1)
        return Poly(c, max_bits)

    def reduce_by_modulus(self, modulus):
        # poly.rs:180-191 (mod_floor == Python %)
        return Poly([x % modulus for x in self.coefficients], modulus.bit_length())

    def divide_by_cyclo(self, cyclo, modulus):
        # poly.rs:113-177, literal long division
        modulus_bits = modulus.bit_length()
        if len(self.coefficients) == 0 or all(c == 0 for c in self.coefficients):
            return (Poly([0] * (cyclo.deg() + 1), modulus_bits),
                    Poly([0] * (2 * cyclo.deg() + 1), modulus_bits))
        dividend = list(self.coefficients)
        divisor = list(cyclo.coefficients)
        quotient = []
        pos = 0  # stands for the repeated `dividend.remove(0)`
        while len(dividend) - pos > len(divisor) - 1:
            _check(divisor[0] != 0, "division by zero leading coefficient (poly.rs:134)")
            # BigInt `/` truncates toward zero; operands are non-negative here
            q = abs(dividend[pos]) // abs(divisor[0])
            if (dividend[pos] < 0) != (divisor[0] < 0):
                q = -q
            quotient.append(q)
            for i, coeff in enumerate(divisor):
                dividend[pos + i] -= q * coeff
            pos += 1
        remainder = dividend[pos:]
        while quotient and quotient[0] == 0:
            quotient.pop(0)
        while remainder and remainder[0] == 0:
            remainder.pop(0)
        _check(len(quotient) >= 1, "quotient.len()-1 underflows (poly.rs:158)")
        while len(quotient) - 1 < cyclo.deg():
            quotient.insert(0, 0)
        if len(remainder) == 0:
            # `remainder.len() - 1` underflows at poly.rs:164 (overflow-checks on)
            raise OracleError("remainder.len()-1 underflows (poly.rs:164)")
        while len(remainder) - 1 < 2 * cyclo.deg():
            remainder.insert(0, 0)
        remainder = [x % modulus for x in remainder]
        return Poly(quotient, modulus_bits), Poly(remainder, modulus_bits)


# --- closed form of divide_by_cyclo for cyclo = x^N + 1 ------
def divide_by_cyclo_closed_form(D, N, modulus):
    """D: 2N-1 coefficients (big-endian) already reduced mod `modulus`.
    Returns (quotient[N+1], remainder[2N+1]) exactly as poly.rs:113-177 does
    when the divisor is x^N + 1 and D != 0."""
    assert len(D) == 2 * N - 1
    q_raw = D[:N - 1]
    rem = [D[N - 1]] + [D[N + t] - D[t] for t in range(N - 1)]
    quotient = [0] * (N + 1 - len(q_raw)) + list(q_raw)
    remainder = [0] * (2 * N + 1 - len(rem)) + [x % modulus for x in rem]
    return quotient, remainder
