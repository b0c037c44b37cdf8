// Modular inverse by the binary extended Euclidean algorithm on 8 x 32-bit limbs (plain integers,
// odd modulus p < 2^255, 0 < a < p).  Shifts, adds and compares only -- none of it touches the
// IMAD pipe that every other kernel of the prover is bound by, and it replaces the 380 dependent
// Montgomery products of a Fermat inversion (~260 us on a lone warp) with ~370 short iterations.
// `__host__ __device__` so the host unit test (tests/test_capi_cpu.py builds tools/inv_host_test.cu)
// can check it against Python big-int arithmetic without a GPU.
#pragma once
#include <cstdint>

namespace zkfhe {

__host__ __device__ __forceinline__ uint32_t u256_add(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
}
__host__ __device__ __forceinline__ uint32_t u256_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t brw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - brw;
        r[i] = (uint32_t)d;
        brw = (uint32_t)(d >> 63);
    }
    return brw;
}
__host__ __device__ __forceinline__ void u256_shr1(uint32_t* a, uint32_t top) {
#pragma unroll
    for (int i = 0; i < 7; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
    a[7] = (a[7] >> 1) | (top << 31);
}
__host__ __device__ __forceinline__ bool u256_geq(const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return true;
}
__host__ __device__ __forceinline__ bool u256_is_one(const uint32_t* a) {
    uint32_t o = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < 8; i++) o |= a[i];
    return o == 0;
}
// x = x / 2 mod p
__host__ __device__ __forceinline__ void u256_half_mod(uint32_t* x, const uint32_t* p) {
    uint32_t c = 0;
    if (x[0] & 1u) c = u256_add(x, x, p);
    u256_shr1(x, c);
}

// r = a^-1 mod p  (plain integers; a in [1, p), p odd).  a == 0 returns 0.
__host__ __device__ inline void u256_inv_odd(uint32_t* r, const uint32_t* a, const uint32_t* p) {
    uint32_t u[8], v[8], x1[8], x2[8];
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = a[i]; v[i] = p[i]; x1[i] = 0; x2[i] = 0; nz |= a[i]; }
    x1[0] = 1;
    if (!nz) {
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = 0;
        return;
    }
#pragma unroll 1
    while (!u256_is_one(u) && !u256_is_one(v)) {
#pragma unroll 1
        while (!(u[0] & 1u)) { u256_shr1(u, 0); u256_half_mod(x1, p); }
#pragma unroll 1
        while (!(v[0] & 1u)) { u256_shr1(v, 0); u256_half_mod(x2, p); }
        if (u256_geq(u, v)) {
            u256_sub(u, u, v);
            if (u256_sub(x1, x1, x2)) u256_add(x1, x1, p);
        } else {
            u256_sub(v, v, u);
            if (u256_sub(x2, x2, x1)) u256_add(x2, x2, p);
        }
    }
    const bool first = u256_is_one(u);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = first ? x1[i] : x2[i];
}

}  // namespace zkfhe
