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
/* Run all subsequent work on `cuda_stream` (a cudaStream_t; NULL is the CUDA legacy default
 * stream).  Until this is called the context uses a private non-blocking stream. */
int zkfhe_set_stream(zkfhe_ctx* ctx, void* cuda_stream);
int zkfhe_sync(zkfhe_ctx* ctx);
/* How the host thread waits for the stream inside the library (commitment read-backs, status checks): on != 0
 * (default) sleeps on a blocking-sync CUDA event, leaving the core to the other proofs' host work (the transcript's
 * Poseidon sponge); 0 spins in cudaStreamSynchronize, a few microseconds quicker per wait for a lone proof. */
int zkfhe_set_blocking_sync(zkfhe_ctx* ctx, int on);
/* Number of kernels launched through this context so far (bench.py's gpu_launches). */
uint64_t zkfhe_launch_count(const zkfhe_ctx* ctx);
/* On-device self test of the generated PTX field arithmetic against an independent plain-C
 * Montgomery product and algebraic identities; `mismatches` receives the failure count. */
int zkfhe_selftest(zkfhe_ctx* ctx, uint32_t n_cases, uint64_t seed, uint32_t* mismatches);
/* Arithmetic ceilings measured on this GPU, for the rooflines bench.py reports (the prove path is
 * 256-bit modular integer work bound by the INT32 multiply pipe, not by HBM).  kind 0: Montgomery
 * products with every SM full (`ops` = products executed in `ms`); kinds 1..5: a dependent chain of
 * `iters` operations on one warp -- 1 XYZZ point addition, 2 mixed addition, 3 field product,
 * 4 field inversion (binary Euclid), 5 field inversion (Fermat).  kinds 6, 7: an experiment the prover does not use --
 * the instruction mix of a Montgomery product on the FP64 pipe (52-bit limbs, fma_rz splitting), on every warp (6)
 * or on the odd warps beside the IMAD product on the even ones (7); `ops` = products of both kinds together.  kinds 8, 9:
 * point additions per second on a full GPU over points gathered from the loaded SRS's window table -- 8 = mixed XYZZ
 * additions (what the MSM runs), 9 = affine additions in batches of 16 pairs per thread sharing one inversion (an
 * experiment; the prover does not use it); `ops` = additions. */
int zkfhe_microbench(zkfhe_ctx* ctx, int kind, uint32_t iters, float* ms, uint64_t* ops);

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
/* Test SRS, the shape of halo2 `ParamsKZG::setup` that halo2-scaffold's `gen_srs(k)` falls back to
 * when no params file exists: g[i] = tau^i * G1, g_lagrange[i] = l_i(tau) * G1, computed on the
 * GPU from an explicit tau (Fr, Montgomery) and loaded as by zkfhe_load_srs.  The bases are also
 * copied to the host when the output pointers are non-NULL (n x 64 bytes each).  INSECURE. */
int zkfhe_srs_setup(zkfhe_ctx* ctx, uint32_t k, const uint8_t* h_tau_fr, uint8_t* h_g_out, uint8_t* h_g_lagrange_out);
/* The tau of the reference's OWN test SRS: halo2-scaffold `gen_srs(k)` without a params file runs
 * `ParamsKZG::setup(k, ChaCha20Rng::from_seed([0; 32]))`, whose first draw `Fr::random` is the first 64 ChaCha20
 * keystream bytes (zero key / counter / nonce: RFC 7539 A.1 #1) as a little-endian 512-bit integer mod r
 * [UPSTREAM-RECALL].  Writes tau (Fr, Montgomery, 32 bytes) and, when `keystream64` is non-NULL, the 64 keystream
 * bytes.  Host only; what `bfv --insecure-test-srs` and bench.py feed to zkfhe_srs_setup.  INSECURE by construction. */
int zkfhe_reference_test_tau(uint8_t* tau_fr32, uint8_t* keystream64);
/* Make `dst` use the resident commitment-key tables of `src` (same GPU; `src` must outlive `dst`). */
int zkfhe_share_srs(zkfhe_ctx* dst, const zkfhe_ctx* src);
/* In place: canonical 256-bit integers -> Montgomery Fr (to_montgomery = 1) or back (0). */
int zkfhe_fr_convert_dev(zkfhe_ctx* ctx, uint8_t* d_data, uint64_t count, int to_montgomery);
/* out[b] = sum_i scalars[b][i] * basis[i], b < batch; scalars are batch x 2^k Fr. */
int zkfhe_msm_g1(zkfhe_ctx* ctx, const uint8_t* h_scalars, uint32_t batch, int basis, uint8_t* h_out_affine);
int zkfhe_msm_g1_dev(zkfhe_ctx* ctx, const uint8_t* d_scalars, uint32_t batch, int basis, uint8_t* d_out_affine);
/* Same, with a hint: `small_values` != 0 says the columns hold witness cells / lookup inputs (values
 * far below the field size, a few full-size ones allowed).  The result is identical; the MSM then
 * buckets with a narrower window whose reduction is 8x cheaper -- what `create_proof` pays for on
 * the 269 advice / permuted-lookup columns of a config-1 proof. */
int zkfhe_msm_g1_dev_ex(zkfhe_ctx* ctx, const uint8_t* d_scalars, uint32_t batch, int basis, int small_values,
                        uint8_t* d_out_affine);

/* ---- stage (1a): off-circuit polynomial arithmetic ----------------------------------------
 * Device-resident mirror of `zk_fhe::poly::Poly` (src/poly.rs:9-13): `len` plain integer
 * coefficients (u256), big-endian, plus the `max_bits` bookkeeping.  All calls are
 * asynchronous on the context's stream; data-dependent `assert!`s of the reference are
 * recorded in a sticky status word that zkfhe_status() reads back (ZKFHE_ERR_ASSERT). */
typedef struct zkfhe_poly zkfhe_poly;
/* Poly::from_string after decimal parsing (poly.rs:21-40): asserts coeff <= modulus (:28). */
int zkfhe_poly_from_u64(zkfhe_ctx* ctx, const uint64_t* h_coeffs, uint32_t len, uint64_t modulus, zkfhe_poly** out);
/* Poly::from_string including the decimal parsing (poly.rs:21-40): `text` holds `len` non-negative decimal
 * integers separated by single commas (no spaces); a malformed number is ZKFHE_ERR_ARG, as the reference's
 * `parse().unwrap()` panics (:25). */
int zkfhe_poly_from_decimal(zkfhe_ctx* ctx, const char* text, size_t text_len, uint32_t len, uint64_t modulus, zkfhe_poly** out);
/* Poly::from_big_int (poly.rs:47-59): asserts bits(coeff) <= max_bits (:51). */
int zkfhe_poly_from_u256(zkfhe_ctx* ctx, const uint64_t* h_coeffs_u256, uint32_t len, uint64_t max_bits, zkfhe_poly** out);
/* Poly::mul (poly.rs:75-103): exact integer product of two equal-degree polynomials, computed
 * as an NTT over Fr (exact while max_bits_a + max_bits_b + log2_ceil(len) < 254, else
 * ZKFHE_ERR_OVERFLOW -- the bound the circuit itself asserts at src/poly_chip.rs:90-94). */
int zkfhe_poly_mul(zkfhe_ctx* ctx, const zkfhe_poly* a, const zkfhe_poly* b, zkfhe_poly** out);
/* Poly::reduce_by_modulus (poly.rs:180-191). */
int zkfhe_poly_reduce_by_modulus(zkfhe_ctx* ctx, const zkfhe_poly* a, uint64_t modulus, zkfhe_poly** out);
/* Poly::divide_by_cyclo (poly.rs:113-177) for cyclo = x^N + 1 (the documented assumption, :111);
 * quotient has N+1, remainder 2N+1 coefficients; the all-zero shortcut (:118-123) is kept. */
int zkfhe_poly_divide_by_cyclo(zkfhe_ctx* ctx, const zkfhe_poly* a, const zkfhe_poly* cyclo, uint64_t modulus,
                               zkfhe_poly** quotient, zkfhe_poly** remainder);
uint32_t zkfhe_poly_len(const zkfhe_poly* p);
uint64_t zkfhe_poly_max_bits(const zkfhe_poly* p);
int zkfhe_poly_download(zkfhe_ctx* ctx, const zkfhe_poly* p, uint64_t* h_out_u256);
void zkfhe_poly_free(zkfhe_poly* p);
/* Synchronise and return the sticky status (ZKFHE_OK or the first recorded reference assert). */
int zkfhe_status(zkfhe_ctx* ctx);

/* ---- stage (1b): in-circuit witness generation --------------------------------------------
 * Mirror of `zk_fhe::poly_chip::PolyChip<F>` (src/poly_chip.rs:19-23) over halo2-base
 * `Context`s: a witness object owns the flat advice vector of each context
 * (context 0: phase-0 gate, 1: phase-1 gate, 2: phase-1 RLC) and the lookup-cell list, all
 * resident in HBM.  Each chip call assigns exactly the cells the CPU builder assigns, in the
 * same order, by one kernel launch over the coefficients. */
typedef struct zkfhe_witness zkfhe_witness;
typedef struct {
    uint32_t ctx_id;    /* which Context holds the cells                                 */
    uint32_t stride;    /* coefficient i lives at advice[base + i * stride]              */
    uint64_t base;
    uint32_t len;       /* degree + 1                                                     */
    uint32_t reserved;
    uint64_t max_num_bits;
} zkfhe_assigned_poly;  /* = PolyChip { assigned_coefficients, max_num_bits, degree }     */
typedef struct {
    uint32_t ctx_id;
    uint32_t reserved;
    uint64_t offset;
} zkfhe_cell;           /* = AssignedValue                                                */

int zkfhe_witness_new(zkfhe_ctx* ctx, uint32_t lookup_bits, zkfhe_witness** out);
void zkfhe_witness_free(zkfhe_witness* w);
int zkfhe_witness_reset(zkfhe_witness* w);   /* keep the buffers, forget the cells (next proof) */
/* PolyChip::from_poly (poly_chip.rs:27-42) */
int zkfhe_chip_from_poly(zkfhe_witness* w, uint32_t ctx_id, const zkfhe_poly* p, zkfhe_assigned_poly* out);
/* ctx.load_constant (examples/bfv.rs:115) */
int zkfhe_chip_load_constant(zkfhe_witness* w, uint32_t ctx_id, uint64_t value, zkfhe_cell* out);
/* PolyChip::to_public (poly_chip.rs:58-62) */
int zkfhe_chip_to_public(zkfhe_witness* w, const zkfhe_assigned_poly* p);
/* The phase-0 challenge gamma (Fr, Montgomery) used by RlcChip in phase 1 (examples/bfv.rs:92-98) */
int zkfhe_chip_set_challenge(zkfhe_witness* w, const uint8_t* h_gamma_fr);
/* PolyChip::constrain_mul (poly_chip.rs:81-116) */
int zkfhe_chip_constrain_mul(zkfhe_witness* w, uint32_t ctx_gate, uint32_t ctx_rlc, const zkfhe_assigned_poly* a,
                             const zkfhe_assigned_poly* b, const zkfhe_assigned_poly* c);
/* PolyChip::add (poly_chip.rs:122-144) */
int zkfhe_chip_add(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a, const zkfhe_assigned_poly* b,
                   zkfhe_assigned_poly* out);
/* PolyChip::scalar_mul (poly_chip.rs:150-174); scalar_value is the constant held by `scalar` */
int zkfhe_chip_scalar_mul(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a, const zkfhe_cell* scalar,
                          uint64_t scalar_value, zkfhe_assigned_poly* out);
/* PolyChip::reduce_by_cyclo (poly_chip.rs:183-223) */
int zkfhe_chip_reduce_by_cyclo(zkfhe_witness* w, uint32_t ctx_gate, uint32_t ctx_rlc, const zkfhe_assigned_poly* self,
                               const zkfhe_assigned_poly* cyclo, const zkfhe_assigned_poly* quotient,
                               const zkfhe_assigned_poly* quotient_times_cyclo, const zkfhe_assigned_poly* remainder,
                               uint64_t modulus, zkfhe_assigned_poly* out);
/* PolyChip::reduce_by_modulo (poly_chip.rs:226-252) */
int zkfhe_chip_reduce_by_modulo(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a, uint64_t modulus,
                                zkfhe_assigned_poly* out);
/* PolyChip::constrain_equality (poly_chip.rs:255-264) */
int zkfhe_chip_constrain_equality(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a,
                                  const zkfhe_assigned_poly* b);
/* PolyChip::constrain_coefficients_in_range (poly_chip.rs:270-317) */
int zkfhe_chip_constrain_coefficients_in_range(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a,
                                               uint64_t z, uint64_t y);
/* PolyChip::constrain_from_distribution_chi_key (poly_chip.rs:320-354) */
int zkfhe_chip_constrain_from_distribution_chi_key(zkfhe_witness* w, uint32_t ctx_gate, const zkfhe_assigned_poly* a,
                                                   uint64_t z);
/* PolyChip::constrain_coefficients_in_modulus_field (poly_chip.rs:357-366) */
int zkfhe_chip_constrain_coefficients_in_modulus_field(zkfhe_witness* w, uint32_t ctx_gate,
                                                       const zkfhe_assigned_poly* a, uint64_t modulus);
/* PolyChip::safe_trim_leading_zeroes (poly_chip.rs:374-399) */
int zkfhe_chip_safe_trim_leading_zeroes(zkfhe_witness* w, const zkfhe_assigned_poly* a, uint32_t degree,
                                        zkfhe_assigned_poly* out);
/* Structure recording (keygen / mock): when switched on before the first chip call, every chip
 * call also records, per advice cell, a flag byte (1: a gate selector is enabled at this cell,
 * 2: Constant cell, 4 / 8: assert_is_const(cell, 0 / 1)) and the cell it is copy-constrained to
 * (0 = none; else ((ctx_id + 1) << 60) | offset), and per lookup cell the advice cell it copies.
 * This is what halo2-base's builder keeps in `Context::{selector, advice_equality_constraints,
 * constant_equality_constraints, cells_to_lookup}` when `witness_gen_only` is false. */
int zkfhe_witness_set_recording(zkfhe_witness* w, int on);
int zkfhe_witness_download_structure(zkfhe_witness* w, uint32_t ctx_id, uint8_t* h_flags, uint64_t* h_copy);
int zkfhe_witness_download_lookup_sources(zkfhe_witness* w, uint64_t* h_src);
int zkfhe_witness_public_cells(const zkfhe_witness* w, uint64_t* h_cell_ids);
/* The reference's `mock` subcommand (README.md:16-22, halo2 MockProver): checks every gate,
 * RLC gate, copy / constant constraint and lookup on the device.  Returns ZKFHE_OK or
 * ZKFHE_ERR_UNSATISFIED; the counts are written either way. */
int zkfhe_witness_mock(zkfhe_witness* w, uint64_t* n_violations, uint64_t* first_bad_cell);
/* Sizes so far: advice cells per context, lookup cells, public instances. */
int zkfhe_witness_counts(const zkfhe_witness* w, uint64_t advice_cells[3], uint64_t* lookup_cells, uint64_t* instances);
/* Copy a flat vector to the host as Fr (Montgomery): which = 0,1,2 advice of that context,
 * 3 = lookup cells (creation order), 4 = instance values. */
int zkfhe_witness_download(zkfhe_witness* w, uint32_t which, uint8_t* h_out_fr);
/* Device address of the same vectors (valid until the next chip call grows the buffer). */
int zkfhe_witness_device_ptr(zkfhe_witness* w, uint32_t which, uint8_t** d_ptr);

/* ---- keygen ------------------------------------------------------------------------------
 * `cargo run --example bfv -- ... keygen` (README.md:28-38): from a witness object built in
 * recording mode on the keygen input (data/bfv/bfv_empty.in) to a proving key resident in HBM:
 * column counts and break points for 2^k rows (the `configs/<name>.json` pinning), selector /
 * constant / table columns, the permutation, their commitments (the verifying key).
 * The SRS for k must be loaded first. */
typedef struct zkfhe_pk zkfhe_pk;
int zkfhe_keygen(zkfhe_witness* w, uint32_t k, uint32_t unusable_rows, zkfhe_pk** out);
void zkfhe_pk_free(zkfhe_pk* pk);
/* The pinning in the reference's configs/bfv.json schema; `needed` receives the size incl. NUL. */
int zkfhe_pk_pinning_json(const zkfhe_pk* pk, char* buf, size_t cap, size_t* needed);
/* out = {k, n_gate0, n_gate1, n_rlc, n_lookup, n_advice, n_perm, n_fixed, n_chunks, usable_rows,
 *        max_rows, lookup_bits, instances, fx_sigma, fx_const, fx_table} */
int zkfhe_pk_info(const zkfhe_pk* pk, uint32_t out[16]);
/* Fixed polynomial `index` as Fr (Montgomery): form 0 = Lagrange values (n), 1 = coefficients (n),
 * 2 = evaluations on the extended coset zeta * H_ext (4n). */
int zkfhe_pk_download_fixed(const zkfhe_pk* pk, uint32_t index, uint32_t form, uint8_t* h_out);
/* The verifying key's commitments: n_fixed G1Affine (Montgomery, 64 bytes each). */
int zkfhe_pk_fixed_commitments(const zkfhe_pk* pk, uint8_t* h_out);

/* data/<name>.pk (README.md:38: `keygen` writes it once, `prove` reads it many times): the layout of the circuit,
 * the public cells, the fixed columns in Lagrange form and their commitments.  Coefficient / extended forms are
 * recomputed on import (two batched NTTs).  Call export with buf = NULL to get the size.  The SRS for the key's k must
 * be loaded on `ctx` before import. */
int zkfhe_pk_export(const zkfhe_pk* pk, uint8_t* buf, size_t cap, size_t* needed);
int zkfhe_pk_import(zkfhe_ctx* ctx, const uint8_t* buf, size_t len, zkfhe_pk** out);

/* ---- prove -------------------------------------------------------------------------------
 * `cargo run --example bfv -- ... prove` (README.md:40-46): snark-verifier-sdk `gen_snark_shplonk`
 * -> halo2 `create_proof`.  The Challenge API makes it two-phase (examples/bfv.rs:92-98): the
 * phase-0 witness is committed first, the challenge gamma comes back, the caller then runs the
 * phase-1 chip calls (the reference's callback) and finishes the proof.
 *   seed32            ChaCha20 key for the blinding factors (the reference uses OS entropy: callers should too)
 *   transcript_kind   1 = Poseidon (what the reference's `prove` runs: snark-verifier's PoseidonTranscript, t = 5,
 *                     rate 4, R_F = 8, R_P = 60; ~1,700 permutations per config-1 proof on the host),
 *                     0 = BLAKE2b (halo2's native transcript; microseconds)
 * The proof is a malloc'd byte string (free with zkfhe_proof_free), in round order: commitments as halo2 writes a
 * bn256 G1Affine (32 bytes: x little-endian, bit 6 of the last byte = parity of y, bit 7 = identity), scalars
 * canonical little-endian (32 bytes). */
typedef struct zkfhe_prover zkfhe_prover;
/* `ctx` is the context (stream) the proof runs on; it may differ from the one keygen ran on (same
 * GPU), so several proofs can be in flight on one GPU, one context per host thread, all sharing one
 * proving key and -- via zkfhe_share_srs -- one copy of the commitment-key tables. */
int zkfhe_prove_begin(zkfhe_ctx* ctx, zkfhe_pk* pk, const uint8_t* seed32, int transcript_kind, zkfhe_prover** out);
int zkfhe_prove_phase0(zkfhe_prover* pr, zkfhe_witness* w, uint8_t* h_gamma_fr_out);
int zkfhe_prove_finish(zkfhe_prover* pr, zkfhe_witness* w, uint8_t** proof_out, size_t* proof_len);
/* Host wall-clock (ms) of the rounds of the last proof; each round ends with a synchronising
 * read-back of its commitments, so these are true latencies: [0] phase-0 commit, [1] phase-1 advice
 * commit, [2] lookup permutations, [3] grand products, [4] quotient, [5] evaluations, [6] SHPLONK
 * quotient commit, [7] final opening commit. */
int zkfhe_prover_round_ms(const zkfhe_prover* pr, double out[8]);
/* Start the next proof with the same prover object (keeps its device buffers). */
int zkfhe_prove_reset(zkfhe_prover* pr, const uint8_t* seed32);
void zkfhe_prover_free(zkfhe_prover* pr);
void zkfhe_proof_free(uint8_t* proof);

/* ---- one proof over several GPUs (nothing of the kind exists in the reference) ---------
 * SPMD: every rank holds the same proving key and SRS on its own GPU and makes the SAME sequence of witness / prove
 * calls on the same input and seed.  Inside zkfhe_prove_* the work then splits into one shard per rank:
 *   - every commitment phase by column (rank r commits the block zkfhe_shard_range gives it; one ncclAllGather of
 *     64 bytes per column);
 *   - the quotient by expression: rank r owns a block of the gate groups, permutation chunks and lookups, transforms
 *     (lagrange -> coeff -> extended coset) only the columns those expressions read, evaluates them on the whole
 *     extended domain and contributes 4n x 32 bytes to one ncclAllGather, after which every rank sums the shares;
 *   - the openings by column: rank r evaluates at x, and adds into the SHPLONK sums, the polynomials it already holds
 *     in coefficient form (two more all-gathers: ~900 / ranks scalars, 6n x 32 bytes).
 * Witness generation, lookup permutations, grand products and the transcript are replicated.  Shares are combined by
 * field additions in rank order on every rank, so every rank returns the same proof, byte-identical to the
 * single-GPU proof.  libnccl.so.2 is loaded with dlopen on first use.  Any rank count works. */
int zkfhe_comm_unique_id(uint8_t* out128);                 /* ncclGetUniqueId: call on one rank, hand the bytes to all */
int zkfhe_comm_init(zkfhe_ctx* ctx, int rank, int n_ranks, const uint8_t* id128);   /* collective: ncclCommInitRank */
int zkfhe_comm_destroy(zkfhe_ctx* ctx);
int zkfhe_comm_info(const zkfhe_ctx* ctx, int* rank, int* n_ranks, int* virtual_ranks);
/* Testing on ONE GPU: compute the `n_ranks` shards one after the other on this context, no collective (0 or 1
 * switches it off).  The proof must not change. */
int zkfhe_set_virtual_ranks(zkfhe_ctx* ctx, int n_ranks);
/* The contiguous block [lo, hi) of `count` items that shard `rank` of `n_ranks` owns (ceil(count / n_ranks) each). */
int zkfhe_shard_range(uint32_t count, uint32_t n_ranks, uint32_t rank, uint32_t* lo, uint32_t* hi);

/* ---- verify (reference `verify` subcommand, README.md:48-54) --------------------------------
 * The pairing check halo2-axiom's VerifierSHPLONK ends with: is prod_i e(P_i, Q_i) == 1 ?
 * P_i are 64-byte G1 affine points (x | y), Q_i 128-byte G2 affine points (x.c0 | x.c1 | y.c0 | y.c1,
 * Fq2 = Fq[u]/(u^2+1)); every coordinate Montgomery (R = 2^256), identity = all zeros.  Host
 * arithmetic: one check per proof. */
int zkfhe_pairing_check(const uint8_t* g1_points, const uint8_t* g2_points, uint32_t count, int* is_one);
/* e(P, Q) itself: 12 Fq coefficients (Montgomery, 384 bytes) of an element of Fq[w]/(w^12 - 18 w^6 + 82), low degree
 * first.  `reference_construction` != 0 computes it the plain way (Fq12 curve arithmetic, one 2790-bit
 * exponentiation -- what oracle/pairing.py restates); 0 is the production path (Q on the twist over Fq2, sparse
 * lines, Frobenius maps, the BN final-exponentiation chain).  Both return the same bytes. */
int zkfhe_pairing(const uint8_t* g1_point, const uint8_t* g2_point, int reference_construction, uint8_t* out384);
/* [tau]_2 of the test SRS made by zkfhe_srs_setup (halo2 `ParamsKZG::setup` keeps s_g2 beside the G1
 * powers): tau as a Montgomery Fr element in, 128 bytes (x.c0 | x.c1 | y.c0 | y.c1, Montgomery) out. */
int zkfhe_srs_g2(const uint8_t* tau_mont32, uint8_t* out128);
/* The proof's point encoding, host only: canonical affine x || y (64 bytes little-endian, identity = zeros) <-> the 32
 * bytes halo2 writes for a bn256 G1Affine (x little-endian, bit 6 of the last byte = parity of y, bit 7 = identity;
 * SURVEY.md App. C.2).  Decompression recovers y = sqrt(x^3 + 3) with that parity and returns ZKFHE_ERR_ARG when the
 * bytes do not encode a curve point (x >= p, no square root, stray bits beside the identity flag). */
int zkfhe_point_compress(const uint8_t* xy_canon64, uint8_t* out32);
int zkfhe_point_decompress(const uint8_t* in32, uint8_t* xy_canon64);
/* The verifying key as bytes -- the reference's data/<name>.vk written by `keygen`: the layout numbers
 * of the circuit and the commitments of the fixed columns.  Call with buf = NULL to get the size. */
int zkfhe_vk_export(const zkfhe_pk* pk, uint8_t* buf, size_t cap, size_t* needed);
/* The reference's `verify`: replay the transcript, check every gate / permutation / lookup identity at
 * the challenge point, fold the SHPLONK multi-open into one MSM over the proof's and the key's
 * commitments (GPU) and finish with the pairing check (host).  `instances`: n canonical 32-byte
 * little-endian scalars; `s_g2`: [tau]_2 as above.  Returns ZKFHE_OK with *accepted = 1 or 0 (the
 * reason for a rejection is zkfhe_last_error); error codes are for malformed arguments only. */
int zkfhe_verify(zkfhe_ctx* ctx, const uint8_t* vk, size_t vk_len, const uint8_t* instances, uint32_t n_instances,
                 const uint8_t* proof, size_t proof_len, const uint8_t* s_g2, int transcript_kind, int* accepted);

/* ---- transcript (host arithmetic; callable without a GPU) ----------------------------------
 * The Poseidon permutation behind transcript kind 1 on five Fr elements (Montgomery, 160 bytes, in place):
 * plain = 0 is the form the transcript runs (sparse partial rounds; AVX-512 IFMA where the CPU has it, else scalar),
 * 1 the textbook form over (round constants, MDS), 2 the scalar optimised form, 3 the IFMA form (ZKFHE_ERR_STATE
 * without AVX-512 IFMA); all return the same state.  Non-canonical input is ZKFHE_ERR_ARG. */
int zkfhe_poseidon_permute(uint8_t* state160, int plain);
/* Replay a message script through a fresh transcript of the given kind: op 1 + 32 bytes = common_scalar (canonical
 * little-endian), op 2 + 64 bytes = common_point (canonical x | y), op 3 = squeeze; the challenges are written to
 * `out` as canonical 32-byte scalars (out = NULL only counts them). */
int zkfhe_transcript_replay(int kind, const uint8_t* script, size_t len, uint8_t* out, size_t cap, size_t* n_challenges);

/* Host arithmetic speed on the calling thread, ns per operation: kind 0 = one Poseidon permutation (the form the
 * transcript runs), 1 = one dependent Fr product, 2 = one plain-form permutation, 3 = scalar optimised form, 4 = the
 * AVX-512 IFMA form.  `features` (optional) receives a short description of the code path. */
int zkfhe_host_microbench(int kind, uint32_t iters, double* ns_per_op, char* features, size_t cap);

/* ---- timing hook -------------------------------------------------------------------------
 * Device time (ms, CUDA events on the context's stream) of the dominant kernel of the last
 * NTT / MSM call: the butterfly passes for NTT, the bucket-accumulation kernel for MSM. */
float zkfhe_last_kernel_ms(const zkfhe_ctx* ctx);
/* Accumulated device time per kernel category since zkfhe_timing_reset (CUDA events recorded on
 * the context's stream around every launch of that category, so whole-proof shares can be read
 * without a profiler).  category 0: MSM bucket accumulation (units = scalar/point pairs),
 * 1: NTT passes (units = field elements transformed), 2: MSM counting sort, 3: MSM bucket folding,
 * 4: MSM final reduction, 7: NCCL collectives of a sharded proof (units = bytes gathered),
 * 5: no time -- units = point additions issued by the accumulate kernel
 * (non-zero signed digits), the numerator of the IMAD-pipe roofline; 6: no time -- units = field
 * products issued by the NTT passes (butterflies + 4-step twiddles + coset / n^-1 factors). */
int zkfhe_timing_reset(zkfhe_ctx* ctx);
int zkfhe_timing_get(zkfhe_ctx* ctx, int category, float* ms, uint32_t* spans, uint64_t* units);

#ifdef __cplusplus
}
#endif
#endif /* ZKFHE_B200_H */
