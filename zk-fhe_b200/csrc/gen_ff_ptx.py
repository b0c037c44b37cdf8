#!/usr/bin/env python3
"""Generates ff_ptx_gen.cuh: 256-bit Montgomery arithmetic for BN254 Fr / Fq as
inline-PTX blocks (8 x 32-bit limbs, one asm statement per field op so the
carry flag never lives across statements).

Multiplication is an operand-scanning (CIOS) Montgomery product arranged so
that every 32x32 product is a (mad.lo, madc.hi) pair on the same operands --
ptxas fuses each pair into one IMAD.WIDE with carry-in/out.  Products whose
column index is even accumulate into one 8-limb register array, odd ones into a
second array offset by one limb; dividing by 2^32 after each reduction step
swaps the two roles, so no limb ever moves.

Because there is no GPU in the build container, the instruction list is
*emulated here* (`selfcheck`) against big-int arithmetic before the header is
written: the same list object is both printed as PTX and interpreted.

Usage:  python gen_ff_ptx.py            (rewrites ff_ptx_gen.cuh next to it)
"""
import os
import random
import sys

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
M32 = 0xFFFFFFFF


def limbs(x, n=8):
    return [(x >> (32 * i)) & M32 for i in range(n)]


class Prog:
    """A straight-line PTX program over 32-bit registers + the carry flag."""

    def __init__(self):
        self.ins = []      # (op, dst, [srcs])  srcs are reg names or int immediates
        self.temps = []

    def t(self, name):
        if name not in self.temps:
            self.temps.append(name)
        return name

    def emit(self, op, dst, *srcs):
        self.ins.append((op, dst, list(srcs)))

    # -- emulation -----------------------------------------------------------
    def run(self, env):
        cf = 0
        pred = {}

        def val(s):
            return s if isinstance(s, int) else env[s]

        for op, dst, srcs in self.ins:
            v = [val(s) for s in srcs if not (isinstance(s, str) and s.startswith("%p_"))]
            if op == "mul.lo":
                env[dst] = (v[0] * v[1]) & M32
            elif op == "mul.hi":
                env[dst] = (v[0] * v[1]) >> 32
            elif op in ("mad.lo.cc", "madc.lo.cc", "madc.hi.cc", "mad.hi.cc", "madc.hi", "madc.lo"):
                prod = v[0] * v[1]
                part = (prod & M32) if ".lo" in op else (prod >> 32)
                full = part + v[2] + (cf if op.startswith("madc") else 0)
                env[dst] = full & M32
                if op.endswith(".cc"):
                    cf = full >> 32
            elif op in ("add.cc", "addc.cc", "addc"):
                full = v[0] + v[1] + (cf if op.startswith("addc") else 0)
                env[dst] = full & M32
                if op.endswith(".cc"):
                    cf = full >> 32
            elif op in ("sub.cc", "subc.cc", "subc"):
                full = v[0] - v[1] - (cf if op.startswith("subc") else 0)
                env[dst] = full & M32
                if op.endswith(".cc"):
                    cf = 1 if full < 0 else 0
            elif op == "setp.eq":
                pred[dst] = (v[0] == v[1])
            elif op == "selp":
                env[dst] = v[0] if pred[srcs[2]] else v[1]
            elif op == "mov":
                env[dst] = v[0]
            else:
                raise ValueError(op)
        return env

    # -- printing ------------------------------------------------------------
    def ptx(self, regmap):
        def r(s):
            if isinstance(s, int):
                return "0x%08x" % s
            return regmap.get(s, s)

        lines = []
        for op, dst, srcs in self.ins:
            if op == "setp.eq":
                lines.append(f"setp.eq.u32 {dst}, {r(srcs[0])}, {r(srcs[1])};")
            elif op == "selp":
                lines.append(f"selp.u32 {r(dst)}, {r(srcs[0])}, {r(srcs[1])}, {srcs[2]};")
            elif op == "mov":
                lines.append(f"mov.u32 {r(dst)}, {r(srcs[0])};")
            else:
                lines.append(f"{op}.u32 {r(dst)}, " + ", ".join(r(s) for s in srcs) + ";")
        return lines


def final_reduce(p, src, mod_l, out="r"):
    """out = src - mod if src >= mod else src   (src < 2*mod)."""
    T = [p.t(f"t{j}") for j in range(8)]
    p.emit("sub.cc", T[0], src[0], mod_l[0])
    for j in range(1, 8):
        p.emit("subc.cc", T[j], src[j], mod_l[j])
    brw = p.t("brw")
    p.emit("subc", brw, 0, 0)                 # 0xffffffff iff src < mod
    p.emit("setp.eq", "%p_ge", brw, 0)
    for j in range(8):
        p.emit("selp", f"{out}{j}", T[j], src[j], "%p_ge")


def gen_mul(mod, square=False):
    """r = a*b*2^-256 mod `mod`, inputs < mod (works for < 2*mod too), r < mod."""
    pl = limbs(mod)
    inv = (-pow(mod, -1, 1 << 32)) % (1 << 32)
    p = Prog()
    a = [f"a{j}" for j in range(8)]
    b = a if square else [f"b{j}" for j in range(8)]

    def reduce_step(E, O):
        m = p.t("m")
        p.emit("mul.lo", m, E[0], inv)
        # odd limbs of the modulus into O (limbs 1..8)
        for k in range(4):
            lo = "mad.lo.cc" if k == 0 else "madc.lo.cc"
            p.emit(lo, O[2 * k], m, pl[2 * k + 1], O[2 * k])
            p.emit("madc.hi.cc" if k < 3 else "madc.hi", O[2 * k + 1], m, pl[2 * k + 1], O[2 * k + 1])
        # even limbs into E (limbs 0..7), carry-out to limb 8 = O[7]
        for k in range(4):
            lo = "mad.lo.cc" if k == 0 else "madc.lo.cc"
            p.emit(lo, E[2 * k], m, pl[2 * k], E[2 * k])
            p.emit("madc.hi.cc", E[2 * k + 1], m, pl[2 * k], E[2 * k + 1])
        p.emit("addc", O[7], O[7], 0)

    # two physical arrays X, Y that alternate roles each row
    X = [p.t(f"x{j}") for j in range(8)]
    Y = [p.t(f"y{j}") for j in range(8)]
    # row 0
    E, O = X, Y
    for k in range(4):
        p.emit("mul.lo", E[2 * k], a[2 * k], b[0])
        p.emit("mul.hi", E[2 * k + 1], a[2 * k], b[0])
        p.emit("mul.lo", O[2 * k], a[2 * k + 1], b[0])
        p.emit("mul.hi", O[2 * k + 1], a[2 * k + 1], b[0])
    reduce_step(E, O)
    for i in range(1, 8):
        Ep, Op = E, O
        E, O = Op, Ep          # after /2^32 the odd array is limb-aligned; reuse Ep's registers for the new odd array
        # limb 0: E[0] += Ep[1]; carry continues into limb 1 = O[0]
        p.emit("add.cc", E[0], E[0], Ep[1])
        for k in range(4):
            c_lo = Ep[2 * k + 2] if 2 * k + 2 < 8 else 0
            c_hi = Ep[2 * k + 3] if 2 * k + 3 < 8 else 0
            p.emit("madc.lo.cc", O[2 * k], a[2 * k + 1], b[i], c_lo)
            p.emit("madc.hi.cc" if k < 3 else "madc.hi", O[2 * k + 1], a[2 * k + 1], b[i], c_hi)
        for k in range(4):
            lo = "mad.lo.cc" if k == 0 else "madc.lo.cc"
            p.emit(lo, E[2 * k], a[2 * k], b[i], E[2 * k])
            p.emit("madc.hi.cc", E[2 * k + 1], a[2 * k], b[i], E[2 * k + 1])
        p.emit("addc", O[7], O[7], 0)
        reduce_step(E, O)
    # merge: result limbs = O[0..7] + E[1..7]
    S = [p.t(f"s{j}") for j in range(8)]
    p.emit("add.cc", S[0], O[0], E[1])
    for j in range(1, 7):
        p.emit("addc.cc", S[j], O[j], E[j + 1])
    p.emit("addc", S[7], O[7], 0)
    final_reduce(p, S, pl)
    return p


def gen_add(mod):
    pl = limbs(mod)
    p = Prog()
    S = [p.t(f"s{j}") for j in range(8)]
    p.emit("add.cc", S[0], "a0", "b0")
    for j in range(1, 7):
        p.emit("addc.cc", S[j], f"a{j}", f"b{j}")
    p.emit("addc", S[7], "a7", "b7")          # a+b < 2*mod < 2^256: no carry out
    final_reduce(p, S, pl)
    return p


def gen_sub(mod):
    pl = limbs(mod)
    p = Prog()
    S = [p.t(f"s{j}") for j in range(8)]
    p.emit("sub.cc", S[0], "a0", "b0")
    for j in range(1, 8):
        p.emit("subc.cc", S[j], f"a{j}", f"b{j}")
    brw = p.t("brw")
    p.emit("subc", brw, 0, 0)                 # 0xffffffff iff a < b
    p.emit("setp.eq", "%p_ge", brw, 0)
    T = [p.t(f"t{j}") for j in range(8)]
    p.emit("add.cc", T[0], S[0], pl[0])
    for j in range(1, 7):
        p.emit("addc.cc", T[j], S[j], pl[j])
    p.emit("addc", T[7], S[7], pl[7])
    for j in range(8):
        p.emit("selp", f"r{j}", S[j], T[j], "%p_ge")
    return p


def selfcheck():
    rng = random.Random(0xB200)
    for mod in (R_MOD, P_MOD):
        mul, sqr, add, sub = gen_mul(mod), gen_mul(mod, True), gen_add(mod), gen_sub(mod)
        rinv = pow(1 << 256, -1, mod)
        edge = [0, 1, 2, mod - 1, mod - 2, (1 << 255) % mod, M32, (1 << 32), mod >> 1, (mod + 1) >> 1]
        cases = [(x, y) for x in edge for y in edge]
        cases += [(rng.randrange(mod), rng.randrange(mod)) for _ in range(4000)]
        for x, y in cases:
            env = {f"a{j}": v for j, v in enumerate(limbs(x))}
            env.update({f"b{j}": v for j, v in enumerate(limbs(y))})
            out = mul.run(dict(env))
            got = sum(out[f"r{j}"] << (32 * j) for j in range(8))
            assert got == x * y * rinv % mod, ("mul", hex(mod), hex(x), hex(y))
            out = sqr.run(dict(env))
            got = sum(out[f"r{j}"] << (32 * j) for j in range(8))
            assert got == x * x * rinv % mod, ("sqr", hex(x))
            out = add.run(dict(env))
            got = sum(out[f"r{j}"] << (32 * j) for j in range(8))
            assert got == (x + y) % mod, ("add", hex(x), hex(y))
            out = sub.run(dict(env))
            got = sum(out[f"r{j}"] << (32 * j) for j in range(8))
            assert got == (x - y) % mod, ("sub", hex(x), hex(y))
        # lazy inputs (< 2*mod) still give a correct, fully reduced product
        for _ in range(500):
            x, y = rng.randrange(2 * mod), rng.randrange(2 * mod)
            env = {f"a{j}": v for j, v in enumerate(limbs(x))}
            env.update({f"b{j}": v for j, v in enumerate(limbs(y))})
            out = mul.run(env)
            got = sum(out[f"r{j}"] << (32 * j) for j in range(8))
            assert got == x * y * rinv % mod


def emit_fn(name, prog, binary=True):
    regmap = {f"r{j}": f"%{j}" for j in range(8)}
    regmap.update({f"a{j}": f"%{8 + j}" for j in range(8)})
    if binary:
        regmap.update({f"b{j}": f"%{16 + j}" for j in range(8)})
    decl = ".reg .u32 " + ", ".join(prog.temps) + ";"
    body = ["{", decl, ".reg .pred %p_ge;"] + prog.ptx(regmap) + ["}"]
    text = "\n".join(f'        "{ln}\\n\\t"' for ln in body)
    outs = ", ".join(f'"=r"(r[{j}])' for j in range(8))
    ins = ", ".join(f'"r"(a[{j}])' for j in range(8))
    if binary:
        ins += ", " + ", ".join(f'"r"(b[{j}])' for j in range(8))
        sig = f"uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]"
    else:
        sig = f"uint32_t (&r)[8], const uint32_t (&a)[8]"
    # early-clobber: outputs are written only by the final selp's, after all inputs are dead
    return (f"__device__ __forceinline__ void {name}({sig}) {{\n"
            f"    asm(\n{text}\n        : {outs}\n        : {ins});\n}}\n")


def main():
    selfcheck()
    out = ["// GENERATED by gen_ff_ptx.py -- do not edit.  Instruction lists were emulated",
           "// against big-int arithmetic (4000 random + edge cases per op per field) at",
           "// generation time.",
           "#pragma once", "#include <cstdint>", "namespace zkfhe { namespace ptx {", ""]
    for tag, mod in (("fr", R_MOD), ("fq", P_MOD)):
        out.append(emit_fn(f"{tag}_mul", gen_mul(mod)))
        out.append(emit_fn(f"{tag}_sqr", gen_mul(mod, True), binary=False))
        out.append(emit_fn(f"{tag}_add", gen_add(mod)))
        out.append(emit_fn(f"{tag}_sub", gen_sub(mod)))
    out.append("} }  // namespace zkfhe::ptx")
    out.append("")
    for tag, mod in (("FR", R_MOD), ("FQ", P_MOD)):
        r1 = (1 << 256) % mod
        consts = {"MOD": mod, "ONE": r1, "R2": r1 * r1 % mod, "R3": r1 * r1 * r1 % mod,
                  "MOD_MINUS_2": mod - 2}
        for k, v in consts.items():
            out.append(f"#define ZKFHE_{tag}_{k} {{" + ", ".join("0x%08xu" % x for x in limbs(v)) + "}")
        out.append(f"#define ZKFHE_{tag}_INV32 0x%08xu" % ((-pow(mod, -1, 1 << 32)) % (1 << 32)))
    # Fr domain constants (Montgomery form): 2^28-th root of unity, zeta (cube root of unity), zeta^2
    root = pow(7, (R_MOD - 1) >> 28, R_MOD)
    zeta = pow(pow(7, (R_MOD - 1) // 3, R_MOD), 2, R_MOD)
    r1 = (1 << 256) % R_MOD
    for k, v in (("ROOT_OF_UNITY", root), ("ROOT_OF_UNITY_INV", pow(root, -1, R_MOD)),
                 ("ZETA", zeta), ("ZETA2", zeta * zeta % R_MOD)):
        out.append(f"#define ZKFHE_FR_{k}_MONT {{" + ", ".join("0x%08xu" % x for x in limbs(v * r1 % R_MOD)) + "}")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ff_ptx_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("selfcheck ok; wrote", path)


if __name__ == "__main__":
    sys.exit(main())
