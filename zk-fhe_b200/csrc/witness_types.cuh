// Objects behind the opaque witness handles of the C ABI (shared by witness.cu, keygen.cu, prover.cu).
#pragma once
#include <vector>
#include "common.cuh"

struct zkfhe_poly {
    zkfhe_ctx* ctx = nullptr;
    zkfhe::fr_t* d = nullptr;       // canonical integers in 32-byte slots, big-endian coefficient order
    uint32_t len = 0;
    uint64_t max_bits = 0;
};

struct DevVec {
    zkfhe::fr_t* p = nullptr;
    size_t cap = 0, size = 0;
};

template <class T> struct DevArr {
    T* p = nullptr;
    size_t cap = 0;
};

struct zkfhe_witness {
    zkfhe_ctx* ctx = nullptr;
    uint32_t lookup_bits = 8;
    bool record = false;            // keygen / mock mode: record selectors, copies, constants
    DevVec adv[3];                  // flat advice of context 0 (phase-0 gate), 1 (phase-1 gate), 2 (phase-1 RLC)
    DevVec lk[3];                   // cells_to_lookup values, per context, creation order
    DevArr<uint8_t> flags[3];       // per advice cell (record mode)
    DevArr<uint64_t> copy[3];
    DevArr<uint64_t> lk_src[3];     // per lookup cell
    std::vector<zkfhe_cell> make_public;
    zkfhe::fr_t gamma;
    bool have_gamma = false;
};
