//! Stand-ins for the halo2-base / axiom-eth types that appear in the reference's `PolyChip` signatures
//! (`halo2_base::{Context, AssignedValue, gates::{GateChip, RangeChip}}`, `axiom_eth::{rlp::rlc::RlcChip, Field}`;
//! reference src/poly_chip.rs:4-11).  Upstream they own CPU cell vectors; here they are handles onto ONE
//! device-resident witness object whose kernels emit exactly the cells halo2-base would.
use crate::{ffi, Device};
use std::marker::PhantomData;
use std::rc::Rc;

/// Marker for the circuit field.  Only BN254 Fr exists on the device (`F::MODULUS` at src/poly_chip.rs:90,135,158,199).
pub trait Field: Copy + 'static {
    const MODULUS: &'static str;
}
#[derive(Clone, Copy, Debug)]
pub struct Fr;
impl Field for Fr {
    const MODULUS: &'static str = "0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001";
}

/// The witness under construction: three flat advice vectors (context 0: phase-0 gate, 1: phase-1 gate, 2: phase-1 RLC)
/// plus the lookup-cell list, all in HBM.
pub struct Witness {
    pub(crate) dev: Rc<Device>,
    pub(crate) raw: *mut ffi::zkfhe_witness,
}
impl Witness {
    pub fn new(dev: Rc<Device>, lookup_bits: u32, record_structure: bool) -> Rc<Self> {
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_witness_new(dev.raw, lookup_bits, &mut raw) });
        if record_structure {
            dev.check(unsafe { ffi::zkfhe_witness_set_recording(raw, 1) });       // keygen / mock
        }
        Rc::new(Witness { dev, raw })
    }
    pub fn reset(&self) {
        self.dev.check(unsafe { ffi::zkfhe_witness_reset(self.raw) });
    }
    /// The phase-0 challenge (Fr, Montgomery bytes) that `RlcChip` uses in phase 1 (examples/bfv.rs:92-98).
    pub fn set_challenge(&self, gamma_fr_mont: &[u8; 32]) {
        self.dev.check(unsafe { ffi::zkfhe_chip_set_challenge(self.raw, gamma_fr_mont.as_ptr()) });
    }
    /// The reference's `mock` subcommand (README.md:16-22).
    pub fn mock(&self) {
        let (mut bad, mut first) = (0u64, 0u64);
        let rc = unsafe { ffi::zkfhe_witness_mock(self.raw, &mut bad, &mut first) };
        assert!(rc == ffi::ZKFHE_OK, "{bad} constraint violations, first at cell {first:#x}");
    }
}
impl Drop for Witness {
    fn drop(&mut self) {
        unsafe { ffi::zkfhe_witness_free(self.raw) }
    }
}

pub const CTX_PHASE0: u32 = 0;
pub const CTX_GATE: u32 = 1;
pub const CTX_RLC: u32 = 2;

/// `halo2_base::Context<F>`: one of the three contexts of a witness.
pub struct Context<F: Field> {
    pub(crate) wit: Rc<Witness>,
    pub(crate) id: u32,
    _f: PhantomData<F>,
}
impl<F: Field> Context<F> {
    pub fn new(wit: Rc<Witness>, id: u32) -> Self {
        Context { wit, id, _f: PhantomData }
    }
    /// `ctx.load_constant` (examples/bfv.rs:115)
    pub fn load_constant(&mut self, value: u64) -> AssignedValue<F> {
        let mut cell = ffi::zkfhe_cell::default();
        self.wit.dev.check(unsafe { ffi::zkfhe_chip_load_constant(self.wit.raw, self.id, value, &mut cell) });
        AssignedValue { cell, value_u64: Some(value), _f: PhantomData }
    }
}

/// `halo2_base::AssignedValue<F>`: a cell; its value stays on the device.
#[derive(Clone, Copy, Debug)]
pub struct AssignedValue<F: Field> {
    pub(crate) cell: ffi::zkfhe_cell,
    pub(crate) value_u64: Option<u64>,       // known on the host only for constants (scalar_mul needs the scalar's bit length)
    pub(crate) _f: PhantomData<F>,
}

/// `halo2_base::gates::GateChip<F>`: stateless upstream, stateless here.
pub struct GateChip<F: Field>(PhantomData<F>);
impl<F: Field> Default for GateChip<F> {
    fn default() -> Self {
        GateChip(PhantomData)
    }
}
/// `halo2_base::gates::RangeChip<F>`: upstream carries `lookup_bits`; the witness object does here.
pub struct RangeChip<F: Field> {
    pub gate: GateChip<F>,
}
impl<F: Field> Default for RangeChip<F> {
    fn default() -> Self {
        RangeChip { gate: GateChip::default() }
    }
}
/// `axiom_eth::rlp::rlc::RlcChip<F>`: upstream carries gamma; `Witness::set_challenge` does here.
pub struct RlcChip<F: Field>(PhantomData<F>);
impl<F: Field> Default for RlcChip<F> {
    fn default() -> Self {
        RlcChip(PhantomData)
    }
}
