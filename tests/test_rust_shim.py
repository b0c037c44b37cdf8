"""The Rust shim crate (rust/zk-fhe-b200) is source only -- no cargo / rustc in this image -- so what CAN be checked is
checked: ffi.rs is exactly what the generator makes of the C header and declares every symbol of the ABI; the wrapper
sources only call functions ffi.rs declares; PolyChip keeps the reference's method names and argument lists."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRATE = os.path.join(ROOT, "rust", "zk-fhe-b200")


def _read(*parts):
    return open(os.path.join(CRATE, *parts)).read()


def test_ffi_rs_is_generated_from_the_header_and_complete():
    import zk_fhe_b200
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    declared = set(re.findall(r"pub fn (zkfhe_[a-z0-9_]+)\(", _read("src", "ffi.rs")))
    assert declared == set(zk_fhe_b200.declared_symbols())
    assert "..." not in _read("src", "ffi.rs")


def test_wrappers_only_call_declared_functions():
    declared = set(re.findall(r"pub fn (zkfhe_[a-z0-9_]+)\(", _read("src", "ffi.rs")))
    for f in ("lib.rs", "poly.rs", "poly_chip.rs", "halo2_shim.rs", "prover.rs"):
        used = set(re.findall(r"ffi::(zkfhe_[a-z0-9_]+)\(", _read("src", f)))
        assert used and used <= declared, (f, used - declared)


def test_poly_chip_keeps_the_reference_signatures():
    """Method names and parameter lists of reference src/poly_chip.rs (:27, :58, :81-88, :122, :150-155, :183-194,
    :226-231, :255, :270-276, :320-325, :357-362), whitespace-insensitive."""
    want = {
        "from_poly": "poly: Poly, ctx: &mut Context<F>",
        "to_public": "&self, make_public: &mut Vec<AssignedValue<F>>",
        "constrain_mul": "&self, b: PolyChip<F>, c: PolyChip<F>, ctx_gate: &mut Context<F>, ctx_rlc: &mut Context<F>, rlc: &RlcChip<F>",
        "add": "&self, ctx: &mut Context<F>, other: PolyChip<F>, gate: &GateChip<F>",
        "scalar_mul": "&self, ctx: &mut Context<F>, scalar: &AssignedValue<F>, gate: &GateChip<F>",
        "reduce_by_cyclo": "&self, cyclo: PolyChip<F>, quotient: PolyChip<F>, quotient_times_cyclo: PolyChip<F>, remainder: PolyChip<F>, "
                           "range: &RangeChip<F>, ctx_gate: &mut Context<F>, ctx_rlc: &mut Context<F>, rlc: &RlcChip<F>, modulus: u64",
        "reduce_by_modulo": "&self, ctx: &mut Context<F>, range: &RangeChip<F>, modulus: u64",
        "constrain_equality": "&self, ctx: &mut Context<F>, other: PolyChip<F>, gate: &GateChip<F>",
        "constrain_coefficients_in_range": "&self, ctx: &mut Context<F>, range: &RangeChip<F>, z: u64, y: u64",
        "constrain_from_distribution_chi_key": "&self, ctx: &mut Context<F>, gate: &GateChip<F>, z: u64",
        "constrain_coefficients_in_modulus_field": "&self, ctx: &mut Context<F>, range: &RangeChip<F>, modulus: u64",
    }
    src = _read("src", "poly_chip.rs")
    norm = lambda s: re.sub(r"\s+", "", s).replace(",)", ")").rstrip(",")
    for name, params in want.items():
        m = re.search(r"pub fn %s\s*\((.*?)\)\s*(->|\{)" % name, src, flags=re.S)
        assert m, name
        got = re.sub(r"(?<![a-z])_(range|gate|rlc):", r"\1:", norm(m.group(1)))
        assert got == norm(params), (name, got)
    for field in ("pub assigned_coefficients: Vec<AssignedValue<F>>", "pub max_num_bits: u64", "pub degree: usize"):
        assert field in src
    poly = _read("src", "poly.rs")
    for sig in ("pub fn from_string(coefficients: Vec<String>, modulus: u64) -> Self", "pub fn deg(&self) -> usize",
                "pub fn mul(&self, other: &Self) -> Self", "pub fn divide_by_cyclo(&self, cyclo: &Poly, modulus: u64) -> (Self, Self)",
                "pub fn reduce_by_modulus(&mut self, modulus: u64) -> Poly"):
        assert sig in poly, sig
