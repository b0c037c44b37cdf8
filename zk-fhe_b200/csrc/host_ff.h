// Host-side BN254 Fr arithmetic (4 x u64 Montgomery, the ABI layout) and the Poseidon
// Fiat-Shamir transcript.  Only the sequential, tiny part of the prover runs here: hashing
// commitments into challenges and a few hundred scalar operations of the opening argument;
// everything proportional to the domain size runs on the GPU.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "poseidon_consts.h"

namespace zkfhe { namespace host {

typedef unsigned __int128 u128;

struct Fr {
    uint64_t l[4];
    bool operator==(const Fr& o) const { return !memcmp(l, o.l, 32); }
    bool operator!=(const Fr& o) const { return !(*this == o); }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
};

static const Fr FR_MOD = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
static const Fr FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const Fr FR_ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fr FR_ZERO = {{0, 0, 0, 0}};
static const uint64_t FR_INV = 0xc2e1f593efffffffULL;
// canonical (non-Montgomery) constants
static const Fr FR_ROOT_OF_UNITY_CANON = {{0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL}};
static const Fr FR_DELTA_CANON = {{0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL}};
static const Fr FR_ZETA_CANON = {{0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL}};

inline bool geq(const Fr& a, const Fr& b) {
    for (int i = 3; i >= 0; i--) {
        if (a.l[i] > b.l[i]) return true;
        if (a.l[i] < b.l[i]) return false;
    }
    return true;
}
inline Fr sub_raw(const Fr& a, const Fr& b) {
    Fr r;
    u128 brw = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - brw;
        r.l[i] = (uint64_t)d;
        brw = (d >> 64) & 1;
    }
    return r;
}
inline Fr add(const Fr& a, const Fr& b) {
    Fr t;
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; t.l[i] = (uint64_t)c; c >>= 64; }
    return (c || geq(t, FR_MOD)) ? sub_raw(t, FR_MOD) : t;
}
inline Fr sub(const Fr& a, const Fr& b) { return geq(a, b) ? sub_raw(a, b) : sub_raw(FR_MOD, sub_raw(b, a)); }
inline Fr neg(const Fr& a) { return a.is_zero() ? a : sub_raw(FR_MOD, a); }
inline Fr mul(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    const uint64_t* p = FR_MOD.l;
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[j] * b.l[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV;
        c = ((u128)m * p[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr o = {{t[0], t[1], t[2], t[3]}};
    return (t[4] || geq(o, FR_MOD)) ? sub_raw(o, FR_MOD) : o;
}
inline Fr sqr(const Fr& a) { return mul(a, a); }
inline Fr to_mont(const Fr& canon) { return mul(canon, FR_R2); }
inline Fr from_mont(const Fr& m) { Fr one = {{1, 0, 0, 0}}; return mul(m, one); }
inline Fr from_u64(uint64_t v) { Fr c = {{v, 0, 0, 0}}; return to_mont(c); }
inline Fr pow_u64(Fr a, uint64_t e) {
    Fr acc = FR_ONE;
    while (e) {
        if (e & 1) acc = mul(acc, a);
        a = sqr(a);
        e >>= 1;
    }
    return acc;
}
inline Fr inv(const Fr& a) {   // Fermat; inv(0) = 0
    Fr e = FR_MOD;
    e.l[0] -= 2;
    Fr acc = FR_ONE;
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            acc = sqr(acc);
            if ((e.l[i] >> b) & 1) acc = mul(acc, a);
        }
    return acc;
}
inline Fr omega(uint32_t k) {   // generator of the 2^k-th roots of unity (Montgomery)
    Fr w = to_mont(FR_ROOT_OF_UNITY_CANON);
    for (uint32_t s = k; s < 28; s++) w = sqr(w);
    return w;
}

// ---- Poseidon permutation (t = 5, full rounds 8, partial rounds 60, x^5) ---------------------
inline void poseidon_permute(Fr s[POSEIDON_T]) {
    const int T = POSEIDON_T, half = POSEIDON_RF / 2, rounds = POSEIDON_RF + POSEIDON_RP;
    for (int r = 0; r < rounds; r++) {
        for (int i = 0; i < T; i++) s[i] = add(s[i], *(const Fr*)POSEIDON_RC[r * T + i]);
        bool full = r < half || r >= half + POSEIDON_RP;
        for (int i = 0; i < (full ? T : 1); i++) {
            Fr x2 = sqr(s[i]);
            s[i] = mul(mul(x2, x2), s[i]);
        }
        Fr n[POSEIDON_T];
        for (int i = 0; i < T; i++) {
            Fr acc = FR_ZERO;
            for (int j = 0; j < T; j++) acc = add(acc, mul(*(const Fr*)POSEIDON_MDS[i * T + j], s[j]));
            n[i] = acc;
        }
        for (int i = 0; i < T; i++) s[i] = n[i];
    }
}

// ---- BLAKE2b-512 (RFC 7693), streaming, clonable ----------------------------------------------
struct Blake2b {
    uint64_t h[8];
    uint8_t buf[128];
    size_t buflen = 0;
    u128 total = 0;
    static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    Blake2b() {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        memcpy(h, iv, sizeof h);
        h[0] ^= 0x01010000ULL ^ 64;   // digest length 64, no key, fanout = depth = 1
    }
    void compress(const uint8_t* block, bool last) {
        static const uint64_t iv[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL,
                                       0xa54ff53a5f1d36f1ULL, 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL,
                                       0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
        static const uint8_t sigma[12][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
        uint64_t m[16], v[16];
        memcpy(m, block, 128);
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = iv[i]; }
        v[12] ^= (uint64_t)total;
        v[13] ^= (uint64_t)(total >> 64);
        if (last) v[14] = ~v[14];
#define ZK_B2G(a, b, c, d, x, y)                                                   \
    v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); v[c] = v[c] + v[d];    \
    v[b] = rotr(v[b] ^ v[c], 24); v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d]; v[b] = rotr(v[b] ^ v[c], 63);
        for (int r = 0; r < 12; r++) {
            const uint8_t* s = sigma[r];
            ZK_B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) ZK_B2G(1, 5, 9, 13, m[s[2]], m[s[3]])
            ZK_B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) ZK_B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
            ZK_B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) ZK_B2G(1, 6, 11, 12, m[s[10]], m[s[11]])
            ZK_B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) ZK_B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
        }
#undef ZK_B2G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    void update(const void* data, size_t len) {
        const uint8_t* p = (const uint8_t*)data;
        while (len) {
            if (buflen == 128) {            // keep the last block for finalisation
                total += 128;
                compress(buf, false);
                buflen = 0;
            }
            size_t take = 128 - buflen < len ? 128 - buflen : len;
            memcpy(buf + buflen, p, take);
            buflen += take; p += take; len -= take;
        }
    }
    void finalize(uint8_t out[64]) const {   // does not disturb the running state
        Blake2b c = *this;
        c.total += c.buflen;
        memset(c.buf + c.buflen, 0, 128 - c.buflen);
        c.compress(c.buf, true);
        memcpy(out, c.h, 64);
    }
};

static const Fr FR_R3 = {{0x5e94d8e1b4bf0040ULL, 0x2a489cbe1cfbb6b8ULL, 0x893cc664a19fcfedULL, 0x0cf8594b7fcc657cULL}};
// 64 uniform bytes (little-endian 512-bit integer) -> Fr (Montgomery), as halo2curves `from_uniform_bytes`
inline Fr from_uniform_bytes(const uint8_t b[64]) {
    Fr d0, d1;
    memcpy(d0.l, b, 32);
    memcpy(d1.l, b + 32, 32);
    return add(mul(d0, FR_R2), mul(d1, FR_R3));
}

// ---- transcript ------------------------------------------------------------------------------
// Two interchangeable Fiat-Shamir hashes over the same message sequence:
//   BLAKE2B  (default) halo2's own `Blake2bWrite` / `Challenge255` shape: every absorbed item is fed
//            to a running BLAKE2b-512 with a one-byte tag; a challenge is the digest of the state so
//            far (tag 0 appended), reduced from 512 bits.  Microseconds per proof on the host.
//   POSEIDON the hash family the reference reaches through snark-verifier's PoseidonTranscript
//            (t = 5, rate 4, R_F = 8, R_P = 60): sponge over state[1..4]; a squeeze pads the buffer
//            with a single 1, absorbs it rate-by-rate and returns state[1].  Points are absorbed as
//            four Fr elements (low / high 128 bits of canonical x and y).  ~60 us per permutation on
//            the host, i.e. tens of ms per proof -- selectable, not the default.
// Every written item is also appended to the proof in canonical little-endian form (points
// uncompressed, x || y, identity = 64 zero bytes), so the verifier replays the same sequence.
enum TranscriptKind { TRANSCRIPT_BLAKE2B = 0, TRANSCRIPT_POSEIDON = 1 };

struct Transcript {
    int kind;
    Blake2b b2;
    Fr state[POSEIDON_T];
    std::vector<Fr> buf;
    std::vector<uint8_t> proof;

    explicit Transcript(int kind_ = TRANSCRIPT_BLAKE2B) : kind(kind_) {
        for (auto& s : state) s = FR_ZERO;
        state[0] = from_u64(0x7a6b666865ULL);   // domain tag "zkfhe"
        const char tag[] = "zkfhe-b200-transcript-v1";
        b2.update(tag, sizeof tag - 1);
    }
    void common_scalar(const Fr& x_mont) {
        if (kind == TRANSCRIPT_POSEIDON) { buf.push_back(x_mont); return; }
        Fr c = from_mont(x_mont);
        uint8_t t = 2;
        b2.update(&t, 1);
        b2.update(c.l, 32);
    }
    void write_scalar(const Fr& x_mont) {
        common_scalar(x_mont);
        Fr c = from_mont(x_mont);
        const uint8_t* b = (const uint8_t*)c.l;
        proof.insert(proof.end(), b, b + 32);
    }
    // canonical affine coordinates (4 x u64 each), identity = all zero
    void common_point(const uint64_t x_canon[4], const uint64_t y_canon[4]) {
        if (kind == TRANSCRIPT_POSEIDON) {
            const uint64_t* cs[2] = {x_canon, y_canon};
            for (auto c : cs) {
                Fr lo = {{c[0], c[1], 0, 0}}, hi = {{c[2], c[3], 0, 0}};
                buf.push_back(to_mont(lo));
                buf.push_back(to_mont(hi));
            }
            return;
        }
        uint8_t t = 1;
        b2.update(&t, 1);
        b2.update(x_canon, 32);
        b2.update(y_canon, 32);
    }
    void write_point(const uint64_t x_canon[4], const uint64_t y_canon[4]) {
        common_point(x_canon, y_canon);
        const uint8_t* bx = (const uint8_t*)x_canon;
        const uint8_t* by = (const uint8_t*)y_canon;
        proof.insert(proof.end(), bx, bx + 32);
        proof.insert(proof.end(), by, by + 32);
    }
    Fr squeeze() {
        if (kind == TRANSCRIPT_POSEIDON) {
            buf.push_back(FR_ONE);
            while (buf.size() % 4) buf.push_back(FR_ZERO);
            for (size_t i = 0; i < buf.size(); i += 4) {
                for (int j = 0; j < 4; j++) state[1 + j] = add(state[1 + j], buf[i + j]);
                poseidon_permute(state);
            }
            buf.clear();
            return state[1];
        }
        uint8_t t = 0;
        b2.update(&t, 1);
        uint8_t d[64];
        b2.finalize(d);
        return from_uniform_bytes(d);
    }
};

} }  // namespace zkfhe::host
