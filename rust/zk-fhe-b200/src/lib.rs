//! Drop-in replacement of the reference crate's public surface for the `prove` path (reference `src/lib.rs:2-3`:
//! `pub mod poly; pub mod poly_chip;`), backed by libzkfhe_b200 on one B200.
//!
//! * `poly::Poly`            -- reference `src/poly.rs`: coefficients live in HBM, `BigInt`s only on request
//! * `poly_chip::PolyChip`   -- reference `src/poly_chip.rs`: same fields, same method signatures; the halo2-base /
//!                              axiom-eth types in those signatures are the thin stand-ins of `halo2_shim`
//! * `halo2_shim`            -- `Context`, `AssignedValue`, `GateChip`, `RangeChip`, `RlcChip`, `Field`: handles onto
//!                              one device-resident witness (`zkfhe_witness`) instead of CPU cell vectors
//! * `prover`                -- keygen / prove / verify drivers (what halo2-scaffold's `run_eth` does for the example)
//! * `ffi`                   -- the raw C ABI, generated from include/zkfhe_b200.h
pub mod ffi;
pub mod halo2_shim;
pub mod poly;
pub mod poly_chip;
pub mod prover;

use std::ffi::CStr;

/// One GPU context (one CUDA stream); everything in this crate hangs off it.
pub struct Device {
    pub(crate) raw: *mut ffi::zkfhe_ctx,
}

impl Device {
    pub fn new(device: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { ffi::zkfhe_init(device, &mut raw) };
        assert!(rc == ffi::ZKFHE_OK, "zkfhe_init({device}) failed: no usable CUDA device (there is no CPU fallback)");
        Device { raw }
    }
    /// The reference panics where the library returns an error code (`assert!`, `unwrap`): keep that behaviour.
    pub(crate) fn check(&self, rc: i32) {
        if rc != ffi::ZKFHE_OK {
            let msg = unsafe { CStr::from_ptr(ffi::zkfhe_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("{msg}");
        }
    }
    /// Synchronise and surface the data-dependent `assert!`s the device kernels recorded (src/poly.rs:28,51,158,164).
    pub fn status(&self) {
        self.check(unsafe { ffi::zkfhe_status(self.raw) });
    }
    pub fn load_srs(&self, k: u32, g: &[u8], g_lagrange: &[u8]) {
        assert!(g.len() == 64 << k && g_lagrange.len() == 64 << k);
        self.check(unsafe { ffi::zkfhe_load_srs(self.raw, k, g.as_ptr(), g_lagrange.as_ptr()) });
    }
    /// INSECURE test setup from an explicit trapdoor (halo2 `ParamsKZG::setup` shape); tests and benchmarks only.
    pub fn srs_setup_insecure(&self, k: u32, tau_fr_mont: &[u8; 32]) {
        self.check(unsafe { ffi::zkfhe_srs_setup(self.raw, k, tau_fr_mont.as_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) });
    }
}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe { ffi::zkfhe_destroy(self.raw) }
    }
}

thread_local! {
    static CURRENT: std::cell::RefCell<Option<std::rc::Rc<Device>>> = std::cell::RefCell::new(None);
}

/// The reference's `Poly` API has no context argument (`Poly::from_string(coefficients, modulus)`), so the device a
/// thread works on is ambient: set it once per thread.
pub fn set_device(dev: std::rc::Rc<Device>) {
    CURRENT.with(|c| *c.borrow_mut() = Some(dev));
}
pub(crate) fn device() -> std::rc::Rc<Device> {
    CURRENT.with(|c| c.borrow().clone()).expect("zk_fhe::set_device has not been called on this thread")
}
