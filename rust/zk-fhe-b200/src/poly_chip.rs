//! `zk_fhe::poly_chip::PolyChip<F>` (reference src/poly_chip.rs:19-23) with the reference's method signatures
//! (:27, :58, :81-88, :122, :150-155, :183-194, :226-231, :255, :270-276, :320-325, :357-362), each forwarding to the
//! `zkfhe_chip_*` entry point of the same name: one kernel launch over the coefficients instead of one halo2-base
//! gate call per coefficient.  The `assert!`s of the reference (overflow guards :94, :138-141, :161-164, :196-201,
//! `z < y` :278) are evaluated by the library and come back as panics through `Device::check`.
use crate::ffi;
use crate::halo2_shim::{AssignedValue, Context, Field, GateChip, RangeChip, RlcChip};
use crate::poly::Poly;
use std::marker::PhantomData;

#[derive(Clone)]
pub struct PolyChip<F: Field> {
    pub assigned_coefficients: Vec<AssignedValue<F>>,
    pub max_num_bits: u64,
    pub degree: usize,
}

impl<F: Field> PolyChip<F> {
    /// The strided view the library works on: coefficient i at `base + i * stride` of one context.
    fn view(&self) -> ffi::zkfhe_assigned_poly {
        let c = &self.assigned_coefficients;
        let first = c[0].cell;
        let stride = if c.len() > 1 { (c[1].cell.offset - first.offset) as u32 } else { 1 };
        debug_assert!(c.iter().enumerate().all(|(i, v)| v.cell.ctx_id == first.ctx_id && v.cell.offset == first.offset + i as u64 * stride as u64));
        ffi::zkfhe_assigned_poly { ctx_id: first.ctx_id, stride, base: first.offset, len: c.len() as u32, reserved: 0, max_num_bits: self.max_num_bits }
    }
    fn from_view(v: ffi::zkfhe_assigned_poly) -> Self {
        let cells = (0..v.len as u64)
            .map(|i| AssignedValue { cell: ffi::zkfhe_cell { ctx_id: v.ctx_id, reserved: 0, offset: v.base + i * v.stride as u64 }, value_u64: None, _f: PhantomData })
            .collect();
        PolyChip { assigned_coefficients: cells, max_num_bits: v.max_num_bits, degree: v.len as usize - 1 }
    }

    /// src/poly_chip.rs:27-42
    pub fn from_poly(poly: Poly, ctx: &mut Context<F>) -> Self {
        let mut out = ffi::zkfhe_assigned_poly::default();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_from_poly(ctx.wit.raw, ctx.id, poly.raw, &mut out) });
        Self::from_view(out)
    }

    /// src/poly_chip.rs:58-62
    pub fn to_public(&self, make_public: &mut Vec<AssignedValue<F>>) {
        make_public.extend(self.assigned_coefficients.iter().copied());
    }
    /// What the driver does with `make_public` once the circuit function returns: tell the witness which cells are
    /// instances (the library keeps the list; order = order of the calls).
    pub fn register_public(&self, ctx: &Context<F>) {
        let v = self.view();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_to_public(ctx.wit.raw, &v) });
    }

    /// src/poly_chip.rs:81-116
    pub fn constrain_mul(&self, b: PolyChip<F>, c: PolyChip<F>, ctx_gate: &mut Context<F>, ctx_rlc: &mut Context<F>, _rlc: &RlcChip<F>) {
        let (va, vb, vc) = (self.view(), b.view(), c.view());
        ctx_gate.wit.dev.check(unsafe { ffi::zkfhe_chip_constrain_mul(ctx_gate.wit.raw, ctx_gate.id, ctx_rlc.id, &va, &vb, &vc) });
    }

    /// src/poly_chip.rs:122-144
    pub fn add(&self, ctx: &mut Context<F>, other: PolyChip<F>, _gate: &GateChip<F>) -> PolyChip<F> {
        let (va, vb) = (self.view(), other.view());
        let mut out = ffi::zkfhe_assigned_poly::default();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_add(ctx.wit.raw, ctx.id, &va, &vb, &mut out) });
        Self::from_view(out)
    }

    /// src/poly_chip.rs:150-174
    pub fn scalar_mul(&self, ctx: &mut Context<F>, scalar: &AssignedValue<F>, _gate: &GateChip<F>) -> PolyChip<F> {
        let va = self.view();
        let value = scalar.value_u64.expect("scalar_mul: the scalar must be a loaded constant (examples/bfv.rs:115)");
        let mut out = ffi::zkfhe_assigned_poly::default();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_scalar_mul(ctx.wit.raw, ctx.id, &va, &scalar.cell, value, &mut out) });
        Self::from_view(out)
    }

    /// src/poly_chip.rs:183-223
    #[allow(clippy::too_many_arguments)]
    pub fn reduce_by_cyclo(
        &self,
        cyclo: PolyChip<F>,
        quotient: PolyChip<F>,
        quotient_times_cyclo: PolyChip<F>,
        remainder: PolyChip<F>,
        _range: &RangeChip<F>,
        ctx_gate: &mut Context<F>,
        ctx_rlc: &mut Context<F>,
        _rlc: &RlcChip<F>,
        modulus: u64,
    ) -> PolyChip<F> {
        let (vs, vc, vq, vqc, vr) = (self.view(), cyclo.view(), quotient.view(), quotient_times_cyclo.view(), remainder.view());
        let mut out = ffi::zkfhe_assigned_poly::default();
        ctx_gate.wit.dev.check(unsafe {
            ffi::zkfhe_chip_reduce_by_cyclo(ctx_gate.wit.raw, ctx_gate.id, ctx_rlc.id, &vs, &vc, &vq, &vqc, &vr, modulus, &mut out)
        });
        Self::from_view(out)
    }

    /// src/poly_chip.rs:226-252
    pub fn reduce_by_modulo(&self, ctx: &mut Context<F>, _range: &RangeChip<F>, modulus: u64) -> PolyChip<F> {
        let va = self.view();
        let mut out = ffi::zkfhe_assigned_poly::default();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_reduce_by_modulo(ctx.wit.raw, ctx.id, &va, modulus, &mut out) });
        Self::from_view(out)
    }

    /// src/poly_chip.rs:255-264
    pub fn constrain_equality(&self, ctx: &mut Context<F>, other: PolyChip<F>, _gate: &GateChip<F>) {
        let (va, vb) = (self.view(), other.view());
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_constrain_equality(ctx.wit.raw, ctx.id, &va, &vb) });
    }

    /// src/poly_chip.rs:270-317
    pub fn constrain_coefficients_in_range(&self, ctx: &mut Context<F>, _range: &RangeChip<F>, z: u64, y: u64) {
        let va = self.view();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_constrain_coefficients_in_range(ctx.wit.raw, ctx.id, &va, z, y) });
    }

    /// src/poly_chip.rs:320-354
    pub fn constrain_from_distribution_chi_key(&self, ctx: &mut Context<F>, _gate: &GateChip<F>, z: u64) {
        let va = self.view();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_constrain_from_distribution_chi_key(ctx.wit.raw, ctx.id, &va, z) });
    }

    /// src/poly_chip.rs:357-366
    pub fn constrain_coefficients_in_modulus_field(&self, ctx: &mut Context<F>, _range: &RangeChip<F>, modulus: u64) {
        let va = self.view();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_constrain_coefficients_in_modulus_field(ctx.wit.raw, ctx.id, &va, modulus) });
    }

    /// src/poly_chip.rs:374-399 (private in the reference; `reduce_by_cyclo` calls it internally there and here)
    #[allow(dead_code)]
    fn safe_trim_leading_zeroes(&self, ctx: &mut Context<F>, _range: &RangeChip<F>, degree: usize) -> PolyChip<F> {
        let va = self.view();
        let mut out = ffi::zkfhe_assigned_poly::default();
        ctx.wit.dev.check(unsafe { ffi::zkfhe_chip_safe_trim_leading_zeroes(ctx.wit.raw, &va, degree as u32, &mut out) });
        Self::from_view(out)
    }
}
