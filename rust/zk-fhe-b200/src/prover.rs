//! keygen / prove / verify: what halo2-scaffold's `run_eth` (reference examples/bfv.rs:311) does around the circuit
//! function, on top of the library's proving system (halo2-shaped: multi-phase advice, lookup and permutation
//! arguments, SHPLONK over KZG; Poseidon transcript by default, as snark-verifier-sdk's `gen_snark_shplonk`).
use crate::halo2_shim::Witness;
use crate::{ffi, Device};
use std::rc::Rc;

pub const TRANSCRIPT_BLAKE2B: i32 = 0;
pub const TRANSCRIPT_POSEIDON: i32 = 1;

pub struct ProvingKey {
    dev: Rc<Device>,
    pub(crate) raw: *mut ffi::zkfhe_pk,
}

impl ProvingKey {
    /// `keygen` (README.md:28-38): `wit` was built in recording mode on the keygen input, both phases run.
    pub fn keygen(dev: Rc<Device>, wit: &Witness, k: u32, unusable_rows: u32) -> Self {
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_keygen(wit.raw, k, unusable_rows, &mut raw) });
        ProvingKey { dev, raw }
    }
    /// data/<name>.pk
    pub fn to_bytes(&self) -> Vec<u8> {
        let mut need = 0usize;
        self.dev.check(unsafe { ffi::zkfhe_pk_export(self.raw, std::ptr::null_mut(), 0, &mut need) });
        let mut buf = vec![0u8; need];
        self.dev.check(unsafe { ffi::zkfhe_pk_export(self.raw, buf.as_mut_ptr(), need, std::ptr::null_mut()) });
        buf
    }
    pub fn from_bytes(dev: Rc<Device>, bytes: &[u8]) -> Self {
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_pk_import(dev.raw, bytes.as_ptr(), bytes.len(), &mut raw) });
        ProvingKey { dev, raw }
    }
    /// data/<name>.vk
    pub fn vk_bytes(&self) -> Vec<u8> {
        let mut need = 0usize;
        self.dev.check(unsafe { ffi::zkfhe_vk_export(self.raw, std::ptr::null_mut(), 0, &mut need) });
        let mut buf = vec![0u8; need];
        self.dev.check(unsafe { ffi::zkfhe_vk_export(self.raw, buf.as_mut_ptr(), need, std::ptr::null_mut()) });
        buf
    }
    /// configs/<name>.json (schema of the reference's configs/bfv.json)
    pub fn pinning_json(&self) -> String {
        let mut need = 0usize;
        self.dev.check(unsafe { ffi::zkfhe_pk_pinning_json(self.raw, std::ptr::null_mut(), 0, &mut need) });
        let mut buf = vec![0u8; need];
        self.dev.check(unsafe { ffi::zkfhe_pk_pinning_json(self.raw, buf.as_mut_ptr() as *mut _, need, std::ptr::null_mut()) });
        buf.pop();
        String::from_utf8(buf).expect("pinning is ASCII")
    }
}
impl Drop for ProvingKey {
    fn drop(&mut self) {
        unsafe { ffi::zkfhe_pk_free(self.raw) }
    }
}

/// One proof: `phase0` commits the phase-0 advice and returns the challenge gamma; the caller runs the phase-1
/// callback (examples/bfv.rs:172-301) and `finish`es.
pub struct Prover {
    dev: Rc<Device>,
    raw: *mut ffi::zkfhe_prover,
}
impl Prover {
    /// `seed`: ChaCha20 key of the blinding factors; take it from the OS (`StdRng::from_entropy()` upstream).
    pub fn begin(dev: Rc<Device>, pk: &ProvingKey, seed: &[u8; 32], transcript_kind: i32) -> Self {
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_prove_begin(dev.raw, pk.raw, seed.as_ptr(), transcript_kind, &mut raw) });
        Prover { dev, raw }
    }
    pub fn reset(&self, seed: &[u8; 32]) {
        self.dev.check(unsafe { ffi::zkfhe_prove_reset(self.raw, seed.as_ptr()) });
    }
    pub fn phase0(&self, wit: &Witness) -> [u8; 32] {
        let mut gamma = [0u8; 32];
        self.dev.check(unsafe { ffi::zkfhe_prove_phase0(self.raw, wit.raw, gamma.as_mut_ptr()) });
        gamma
    }
    pub fn finish(&self, wit: &Witness) -> Vec<u8> {
        let (mut p, mut len) = (std::ptr::null_mut(), 0usize);
        self.dev.check(unsafe { ffi::zkfhe_prove_finish(self.raw, wit.raw, &mut p, &mut len) });
        let out = unsafe { std::slice::from_raw_parts(p, len) }.to_vec();
        unsafe { ffi::zkfhe_proof_free(p) };
        out
    }
}
impl Drop for Prover {
    fn drop(&mut self) {
        unsafe { ffi::zkfhe_prover_free(self.raw) }
    }
}

/// `verify` (README.md:48-54).  `instances`: canonical 32-byte little-endian scalars; `s_g2`: [tau]_2 of the SRS.
pub fn verify(dev: &Device, vk: &[u8], instances: &[u8], proof: &[u8], s_g2: &[u8; 128], transcript_kind: i32) -> bool {
    assert!(instances.len() % 32 == 0);
    let mut ok = 0;
    dev.check(unsafe {
        ffi::zkfhe_verify(dev.raw, vk.as_ptr(), vk.len(), instances.as_ptr(), (instances.len() / 32) as u32, proof.as_ptr(), proof.len(),
                          s_g2.as_ptr(), transcript_kind, &mut ok)
    });
    ok == 1
}

/// One proof over several GPUs: bind an NCCL communicator (unique id made on rank 0, handed to every rank) and make
/// the same calls with the same input and seed on every rank; each returns the same proof bytes.
pub fn comm_unique_id() -> [u8; 128] {
    let mut id = [0u8; 128];
    assert!(unsafe { ffi::zkfhe_comm_unique_id(id.as_mut_ptr()) } == ffi::ZKFHE_OK, "libnccl.so.2 is not loadable");
    id
}
pub fn comm_init(dev: &Device, rank: i32, n_ranks: i32, id: &[u8; 128]) {
    dev.check(unsafe { ffi::zkfhe_comm_init(dev.raw, rank, n_ranks, id.as_ptr()) });
}
