/* libzkfhe_b200 -- C ABI of the B200-native `prove` hot path for zk-fhe's BFV
 * encryption circuit.
 *
 * The reference (enricobottazzi/zk-fhe) is pure Rust with no FFI of its own; these
 * are the entry points a Rust `extern "C"` block would bind to replace, for the
 * prove path only, the arithmetic it reaches today through:
 *
 *   stage (1)  src/poly.rs:75-103 (Poly::mul), :113-177 (divide_by_cyclo),
 *              :180-191 (reduce_by_modulus), called from examples/bfv.rs:131-150;
 *              and the per-cell witness values computed as a side effect of the
 *              halo2-base gate calls in src/poly_chip.rs:81-399, called from
 *              examples/bfv.rs:172-301;
 *   stage (2)  halo2-axiom `ParamsKZG::commit_lagrange/commit` -> `best_multiexp`
 *              (un-vendored dependency, Cargo.toml:9-11; reached via
 *              examples/bfv.rs:311 `run_eth`);
 *   stage (3)  halo2-axiom `EvaluationDomain::{lagrange_to_coeff, coeff_to_extended,
 *              extended_to_coeff}` -> `best_fft` (same dependency).
 *
 * Data layout (zero-copy from Rust slices):
 *   Fr / Fq    32 bytes = 4 x u64 little-endian limbs, Montgomery form, R = 2^256
 *              (halo2curves `bn256::Fr` / `Fq` in memory).
 *   G1Affine   64 bytes = x || y (each Fq as above); identity = 64 zero bytes.
 *   u256       32 bytes = 4 x u64 little-endian limbs of a plain non-negative integer
 *              (stands in for num-bigint `BigInt` polynomial coefficients).
 *   Polynomials are big-endian in the coefficient index (index 0 = highest degree),
 *   as in src/poly.rs:17,43 and data/bfv/bfv.in.
 *
 * Conventions: every function returns ZKFHE_OK (0) or a negative error code and
 * never aborts or unwinds across the boundary (the reference uses panic!/assert!,
 * release profile panic = "abort", Cargo.toml:47).  `zkfhe_last_error` returns a
 * message for the last failure on that context.  A context is bound to one GPU
 * and one CUDA stream and must be used from one host thread at a time.  Pointers
 * named `h_*` are host memory, `d_*` are device memory on the context's GPU.
 * There is no CPU fallback: without a CUDA device `zkfhe_init` fails.
 */
#ifndef ZKFHE_B200_H
#define ZKFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKFHE_OK 0
#define ZKFHE_ERR_CUDA (-1)        /* CUDA runtime failure (message has details)            */
#define ZKFHE_ERR_ARG (-2)         /* invalid argument / unsupported size                   */
#define ZKFHE_ERR_STATE (-3)       /* call order (e.g. MSM before zkfhe_load_srs)           */
#define ZKFHE_ERR_ASSERT (-4)      /* a reference `assert!`/panic condition was hit         */
#define ZKFHE_ERR_OVERFLOW (-5)    /* reference overflow guard (max_bits >= 254) tripped    */
#define ZKFHE_ERR_UNSATISFIED (-6) /* witness does not satisfy the circuit (mock / prove)   */

typedef struct zkfhe_ctx zkfhe_ctx;

/* ---- lifecycle -------------------------------------------------------------------------- */
int zkfhe_init(int device, zkfhe_ctx** ctx);
void zkfhe_destroy(zkfhe_ctx* ctx);
const char* zkfhe_last_error(const zkfhe_ctx* ctx);
const char* zkfhe_version(void);
/* Run all subsequent work on `cuda_stream` (a cudaStream_t); NULL = the context's own stream. */
int zkfhe_set_stream(zkfhe_ctx* ctx, void* cuda_stream);
int zkfhe_sync(zkfhe_ctx* ctx);
/* Number of kernels launched through this context so far (bench.py's gpu_launches). */
uint64_t zkfhe_launch_count(const zkfhe_ctx* ctx);
/* On-device self test of the generated PTX field arithmetic against an independent plain-C
 * Montgomery product and algebraic identities; `mismatches` receives the failure count. */
int zkfhe_selftest(zkfhe_ctx* ctx, uint32_t n_cases, uint64_t seed, uint32_t* mismatches);

/* ---- stage (3): NTT over BN254 Fr --------------------------------------------------------
 * Replaces halo2 `best_fft` + EvaluationDomain scaling.  `batch` columns of n = 2^log_n
 * elements each, natural order in, natural order out, in place.
 *   inverse = 0: out[j] = sum_i in[i] * w^(i*j)          (lagrange <- coeff)
 *   inverse = 1: out    = n^-1 * sum_i in[i] * w^(-i*j)  (coeff <- lagrange)
 *   coset   = 1: forward: in[i] is first multiplied by zeta^i (coeff_to_extended);
 *                inverse: out[i] is finally multiplied by zeta^-i (extended_to_coeff);
 *                zeta = Fr::ZETA (cube root of unity, so only zeta^(i mod 3) is needed).
 * 1 <= log_n <= 22. */
int zkfhe_ntt_fr(zkfhe_ctx* ctx, uint8_t* h_data, uint32_t log_n, uint32_t batch, int inverse, int coset);
int zkfhe_ntt_fr_dev(zkfhe_ctx* ctx, uint8_t* d_data, uint32_t log_n, uint32_t batch, int inverse, int coset);
/* coeff_to_extended: each input column has n_in = 2^log_n_in coefficients (column stride
 * n_in), zero-extended to 2^log_n_out, multiplied by zeta^i and transformed; output column
 * stride 2^log_n_out. */
int zkfhe_coeff_to_extended_dev(zkfhe_ctx* ctx, const uint8_t* d_coeffs, uint32_t log_n_in, uint8_t* d_ext,
                                uint32_t log_n_out, uint32_t batch);

/* ---- stage (2): MSM over the KZG commitment key ------------------------------------------
 * Replaces `ParamsKZG::{commit, commit_lagrange}` -> `best_multiexp`.
 * zkfhe_load_srs uploads both bases (n = 2^k points each) and builds the resident
 * fixed-base window tables; `basis` selects 0 = g (coefficient form), 1 = g_lagrange. */
int zkfhe_load_srs(zkfhe_ctx* ctx, uint32_t k, const uint8_t* h_g, const uint8_t* h_g_lagrange);
/* out[b] = sum_i scalars[b][i] * basis[i], b < batch; scalars are batch x 2^k Fr. */
int zkfhe_msm_g1(zkfhe_ctx* ctx, const uint8_t* h_scalars, uint32_t batch, int basis, uint8_t* h_out_affine);
int zkfhe_msm_g1_dev(zkfhe_ctx* ctx, const uint8_t* d_scalars, uint32_t batch, int basis, uint8_t* d_out_affine);

/* ---- timing hook -------------------------------------------------------------------------
 * Device time (ms, CUDA events on the context's stream) of the dominant kernel of the last
 * NTT / MSM call: the butterfly passes for NTT, the bucket-accumulation kernel for MSM. */
float zkfhe_last_kernel_ms(const zkfhe_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ZKFHE_B200_H */
