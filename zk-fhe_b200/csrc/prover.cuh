// Shared declarations of the proving system (keygen.cu, prover.cu): the proving key and the
// column layout of the BFV circuit.
//
// Column model (the reference's `EthConfigParams`, configs/bfv.json):
//   advice columns, in this order:
//     [0, n_gate0)                     phase-0 gate columns  (context 0 cut at its break points)
//     [.., + n_gate1)                  phase-1 gate columns  (context 1)
//     [.., + n_rlc)                    phase-1 RLC columns   (context 2)
//     [.., + n_lookup)                 phase-1 lookup-advice columns (cells_to_lookup, max_rows each)
//   fixed columns: one selector per gate column, one per RLC column, the constants column, the
//   lookup table column; plus l_0, l_last, l_active and the permutation polynomials sigma_c.
//   one instance column.
// Gates:  q_c(X) * (a_c(X) + a_c(wX) * a_c(w^2 X) - a_c(w^3 X)) = 0          (halo2-base vertical gate)
//         q_r(X) * (a_r(X) * gamma + a_r(wX) - a_r(w^2 X)) = 0               (axiom-eth RLC gate)
// Lookups: every lookup-advice column is looked up in the table column [0, 2^lookup_bits).
// Permutation over all advice columns, the constants column and the instance column.
#pragma once
#include <array>
#include <string>
#include <vector>
#include "common.cuh"
#include "host_ff.h"

namespace zkfhe {

static constexpr uint32_t BLINDING_FACTORS = 6;   // halo2 ConstraintSystem::blinding_factors() for this shape
static constexpr uint32_t PERM_CHUNK = 2;         // permutation columns per grand product (cs.degree() - 2, degree 4)
static constexpr uint32_t EXT_SHIFT = 2;          // extended domain = 2^(k + 2): quotient degree 3n - 4

struct ColumnCut {                 // one context cut into columns
    std::vector<uint32_t> break_points;
    std::vector<uint64_t> start;   // flat offset of row 0 of each column
    std::vector<uint32_t> rows;    // rows used in each column
};

}  // namespace zkfhe

struct zkfhe_pk {
    zkfhe_ctx* ctx = nullptr;
    uint32_t k = 0, n = 0, unusable_rows = 0, lookup_bits = 0, max_rows = 0, usable = 0 /* u = n - bf - 1 */;
    uint64_t cells[3] = {0, 0, 0}, lookups = 0, instances = 0;
    zkfhe::ColumnCut cut[3];
    uint32_t n_gate0 = 0, n_gate1 = 0, n_rlc = 0, n_lookup = 0;
    uint32_t n_advice = 0, n_perm = 0, n_fixed = 0, n_chunks = 0;
    // fixed polynomial index: q_gate[n_gate0 + n_gate1], q_rlc[n_rlc], constants, table, l0, l_last, l_active, sigma[n_perm]
    uint32_t fx_qgate = 0, fx_qrlc = 0, fx_const = 0, fx_table = 0, fx_l0 = 0, fx_llast = 0, fx_lactive = 0, fx_sigma = 0;
    zkfhe::fr_t* fixed_lagrange = nullptr;   // [n_fixed][n]
    zkfhe::fr_t* fixed_coeff = nullptr;      // [n_fixed][n]
    zkfhe::fr_t* fixed_ext = nullptr;        // [n_fixed][4n]  evaluations on zeta * H_ext
    zkfhe::fr_t* delta_pow = nullptr;        // [n_perm] delta^c (Montgomery)
    std::vector<zkfhe::g1_affine> fixed_commitments;   // host copy (Montgomery affine), the verifying key
    std::vector<std::array<uint64_t, 8>> fixed_commitments_canon;   // canonical x || y, as hashed and serialised
    std::vector<uint64_t> public_cells;      // cell ids exposed as instances, in order
    zkfhe::host::Fr vk_digest;
    std::string pinning_json;
};
