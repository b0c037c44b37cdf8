//! `zk_fhe::poly::Poly` (reference src/poly.rs:9-13) over device-resident coefficients.
//! Same constructor and method signatures; coefficients are big-endian (index 0 = highest degree, src/poly.rs:17,43).
use crate::{device, ffi};
use num_bigint::{BigInt, Sign};

pub struct Poly {
    pub(crate) raw: *mut ffi::zkfhe_poly,
    pub degree: usize,
    pub max_bits: u64,
}

impl Poly {
    fn wrap(raw: *mut ffi::zkfhe_poly) -> Self {
        let len = unsafe { ffi::zkfhe_poly_len(raw) } as usize;
        Poly { raw, degree: len - 1, max_bits: unsafe { ffi::zkfhe_poly_max_bits(raw) } }
    }

    /// src/poly.rs:21-40.  The decimal strings are parsed on the C side in one pass; a malformed number or a
    /// coefficient above `modulus` panics, as `parse().unwrap()` (:25) and the assert (:28) do.
    pub fn from_string(coefficients: Vec<String>, modulus: u64) -> Self {
        let text = coefficients.join(",");
        let dev = device();
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe {
            ffi::zkfhe_poly_from_decimal(dev.raw, text.as_ptr() as *const _, text.len(), coefficients.len() as u32, modulus, &mut raw)
        });
        Self::wrap(raw)
    }

    /// src/poly.rs:47-59 (the reference keeps this private; public here for completeness).
    pub fn from_big_int(coefficients: Vec<BigInt>, max_bits: u64) -> Self {
        let mut limbs = vec![0u64; 4 * coefficients.len()];
        for (i, c) in coefficients.iter().enumerate() {
            let (sign, digits) = c.to_u64_digits();
            assert!(sign != Sign::Minus && digits.len() <= 4, "coefficient does not fit the u256 the ABI carries");
            limbs[4 * i..4 * i + digits.len()].copy_from_slice(&digits);
        }
        let dev = device();
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_poly_from_u256(dev.raw, limbs.as_ptr(), coefficients.len() as u32, max_bits, &mut raw) });
        Self::wrap(raw)
    }

    /// src/poly.rs:62
    pub fn deg(&self) -> usize {
        self.degree
    }

    /// src/poly.rs:75-103: the exact integer product (an NTT over Fr on the GPU instead of the O(N^2) BigInt loop).
    pub fn mul(&self, other: &Self) -> Self {
        assert_eq!(self.deg(), other.deg());                                  // :78
        let dev = device();
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_poly_mul(dev.raw, self.raw, other.raw, &mut raw) });
        Self::wrap(raw)
    }

    /// src/poly.rs:113-177 for cyclo = x^N + 1 (the documented assumption, :111).
    pub fn divide_by_cyclo(&self, cyclo: &Poly, modulus: u64) -> (Self, Self) {
        let dev = device();
        let (mut q, mut r) = (std::ptr::null_mut(), std::ptr::null_mut());
        dev.check(unsafe { ffi::zkfhe_poly_divide_by_cyclo(dev.raw, self.raw, cyclo.raw, modulus, &mut q, &mut r) });
        (Self::wrap(q), Self::wrap(r))
    }

    /// src/poly.rs:180-191 (`&mut self` kept from the reference signature; `self` is not modified there either).
    pub fn reduce_by_modulus(&mut self, modulus: u64) -> Poly {
        let dev = device();
        let mut raw = std::ptr::null_mut();
        dev.check(unsafe { ffi::zkfhe_poly_reduce_by_modulus(dev.raw, self.raw, modulus, &mut raw) });
        Self::wrap(raw)
    }

    /// The reference's public `coefficients: Vec<BigInt>` field, materialised on request (a device -> host copy).
    pub fn coefficients(&self) -> Vec<BigInt> {
        let len = self.degree + 1;
        let mut limbs = vec![0u64; 4 * len];
        let dev = device();
        dev.check(unsafe { ffi::zkfhe_poly_download(dev.raw, self.raw, limbs.as_mut_ptr()) });
        (0..len)
            .map(|i| {
                let bytes: Vec<u8> = limbs[4 * i..4 * i + 4].iter().flat_map(|l| l.to_le_bytes()).collect();
                BigInt::from_bytes_le(Sign::Plus, &bytes)
            })
            .collect()
    }
}

impl Drop for Poly {
    fn drop(&mut self) {
        unsafe { ffi::zkfhe_poly_free(self.raw) }
    }
}
