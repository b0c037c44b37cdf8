// The `verify` path (reference: `cargo run --example bfv -- ... verify`, README.md:48-54, which
// reaches halo2-axiom `verify_proof` + `VerifierSHPLONK` through halo2-scaffold [UPSTREAM,
// un-vendored; SURVEY.md §3.4, §8(f) rank 1]).
//
// A verifier for the proofs prover.cu writes: transcript replay, every gate / permutation / lookup
// identity at the challenge point x against the quotient commitment, the SHPLONK multi-open folded
// into ONE multi-scalar multiplication over the ~770 commitments of the proof and the verifying key
// (run on the GPU through the same fixed-base Pippenger pipeline as the prover: the commitments are
// expanded into a throw-away window table), and the final pairing check
//     e(F + u W', [1]_2) * e(-W', [tau]_2) == 1
// on the host (host_pairing.h).  Scalar work is a few thousand Fr operations on the host, like the
// prover's transcript.  oracle/verifier.py is the independent Python restatement the tests compare
// accept / reject decisions with.
#include "prover.cuh"
#include "host_pairing.h"

using namespace zkfhe;
using host::Fr;

namespace zkfhe {
int msm_variable_base(zkfhe_ctx* ctx, const g1_affine* h_points, const fr_t* h_scalars, uint32_t count, g1_affine* h_out);
}

namespace {

struct VkHeader {
    char magic[8];
    uint32_t k, n_gate0, n_gate1, n_rlc, n_lookup, n_advice, n_perm, n_fixed, n_chunks, usable, lookup_bits, instances,
        unusable_rows, reserved[3];
};
static_assert(sizeof(VkHeader) == 72, "vk header layout");
const char VK_MAGIC[8] = {'Z', 'K', 'F', 'H', 'E', 'V', 'K', '1'};

constexpr int ROT_LAST = 1000;
const int SET_ROTS[6][4] = {{0, 0, 0, 0}, {0, 1, 2, 3}, {0, 1, 2, 0}, {0, -1, 0, 0}, {0, 1, 0, 0}, {0, 1, ROT_LAST, 0}};
const int SET_SIZE[6] = {1, 4, 3, 2, 2, 3};
enum { SET_0 = 0, SET_0123 = 1, SET_012 = 2, SET_0m1 = 3, SET_01 = 4, SET_01L = 5 };
int point_index(int rot) { return rot == ROT_LAST ? 5 : rot + 1; }

struct Point { uint64_t c[8]; };       // canonical x || y, identity = zeros

bool on_curve(const Point& p) {
    host::Fq x, y;
    memcpy(x.l, p.c, 32);
    memcpy(y.l, p.c + 4, 32);
    if (x.is_zero() && y.is_zero()) return true;
    if (host::fq_geq(x, host::FQ_MOD) || host::fq_geq(y, host::FQ_MOD)) return false;
    x = host::fq_to_mont(x);
    y = host::fq_to_mont(y);
    return host::fq_mul(y, y) == host::fq_add(host::fq_mul(host::fq_mul(x, x), x), host::fq_from_u64(3));
}
// 32-byte compressed point (host::Transcript::compress_point) -> canonical x || y; false when the bytes are not the
// encoding of a curve point (x >= p, x^3 + 3 not a square, stray bits on the identity).  p = 3 mod 4: sqrt(a) = a^((p+1)/4).
bool decompress_point(const uint8_t in[32], Point& out) {
    uint8_t xb[32];
    memcpy(xb, in, 32);
    const bool inf = (xb[31] & 0x80) != 0, odd = (xb[31] & 0x40) != 0;
    xb[31] &= 0x3f;
    host::Fq x;
    memcpy(x.l, xb, 32);
    memset(out.c, 0, sizeof out.c);
    if (inf) return x.is_zero() && !odd;
    if (host::fq_geq(x, host::FQ_MOD)) return false;
    const host::Fq xm = host::fq_to_mont(x);
    const host::Fq rhs = host::fq_add(host::fq_mul(host::fq_mul(xm, xm), xm), host::fq_from_u64(3));
    host::Fq e = host::FQ_MOD;                          // (p + 1) / 4: p + 1 does not overflow 256 bits (p < 2^254)
    e.l[0] += 1;                                        // p is odd and p = 3 mod 4, so the low limb does not carry
    for (int i = 0; i < 4; i++) e.l[i] = (e.l[i] >> 2) | (i < 3 ? e.l[i + 1] << 62 : 0);
    host::Fq y = host::FQ_ONE;
    for (int i = 3; i >= 0; i--)
        for (int b = 63; b >= 0; b--) {
            y = host::fq_mul(y, y);
            if ((e.l[i] >> b) & 1) y = host::fq_mul(y, rhs);
        }
    if (!(host::fq_mul(y, y) == rhs)) return false;
    host::Fq yc = host::fq_from_mont(y);
    if (yc.is_zero() && odd) return false;              // no point with y = 0 on this curve anyway (-3 is not a cube)
    if (((yc.l[0] & 1) != 0) != odd) yc = host::fq_sub_raw(host::FQ_MOD, yc);
    memcpy(out.c, x.l, 32);
    memcpy(out.c + 4, yc.l, 32);
    return true;
}
g1_affine to_device_point(const Point& p) {
    host::Fq x, y;
    memcpy(x.l, p.c, 32);
    memcpy(y.l, p.c + 4, 32);
    x = host::fq_to_mont(x);
    y = host::fq_to_mont(y);
    g1_affine a;
    memcpy(a.x.v, x.l, 32);
    memcpy(a.y.v, y.l, 32);
    return a;
}

struct Reject { std::string why; };

}  // namespace

extern "C" {

int zkfhe_pairing_check(const uint8_t* g1_points, const uint8_t* g2_points, uint32_t count, int* is_one) {
    if ((!g1_points || !g2_points) && count) return ZKFHE_ERR_ARG;
    if (!is_one) return ZKFHE_ERR_ARG;
    std::vector<host::G1Aff> p(count);
    std::vector<host::G2Aff> q(count);
    for (uint32_t i = 0; i < count; i++) {
        memcpy(&p[i].x, g1_points + 64 * (size_t)i, 32);
        memcpy(&p[i].y, g1_points + 64 * (size_t)i + 32, 32);
        p[i].inf = p[i].x.is_zero() && p[i].y.is_zero();
        const uint8_t* g = g2_points + 128 * (size_t)i;
        memcpy(&q[i].x.c0, g, 32);
        memcpy(&q[i].x.c1, g + 32, 32);
        memcpy(&q[i].y.c0, g + 64, 32);
        memcpy(&q[i].y.c1, g + 96, 32);
        q[i].inf = q[i].x.c0.is_zero() && q[i].x.c1.is_zero() && q[i].y.c0.is_zero() && q[i].y.c1.is_zero();
        for (const host::Fq* c : {&p[i].x, &p[i].y, &q[i].x.c0, &q[i].x.c1, &q[i].y.c0, &q[i].y.c1})
            if (host::fq_geq(*c, host::FQ_MOD)) return ZKFHE_ERR_ARG;       // not a reduced field element
    }
    *is_one = host::pairing_product_is_one(p.data(), q.data(), (int)count) ? 1 : 0;
    return ZKFHE_OK;
}

// e(P, Q) as an element of GT: 12 Fq coefficients (Montgomery) of Fq[w]/(w^12 - 18 w^6 + 82), low degree first.
// reference_construction != 0 selects the plain construction (Fq12 curve arithmetic, one long exponentiation) that
// oracle/pairing.py restates; both give the same 384 bytes.
int zkfhe_pairing(const uint8_t* g1_point, const uint8_t* g2_point, int reference_construction, uint8_t* out384) {
    if (!g1_point || !g2_point || !out384) return ZKFHE_ERR_ARG;
    host::G1Aff p;
    host::G2Aff q;
    memcpy(&p.x, g1_point, 32);
    memcpy(&p.y, g1_point + 32, 32);
    p.inf = p.x.is_zero() && p.y.is_zero();
    memcpy(&q.x.c0, g2_point, 32);
    memcpy(&q.x.c1, g2_point + 32, 32);
    memcpy(&q.y.c0, g2_point + 64, 32);
    memcpy(&q.y.c1, g2_point + 96, 32);
    q.inf = q.x.c0.is_zero() && q.x.c1.is_zero() && q.y.c0.is_zero() && q.y.c1.is_zero();
    for (const host::Fq* c : {&p.x, &p.y, &q.x.c0, &q.x.c1, &q.y.c0, &q.y.c1})
        if (host::fq_geq(*c, host::FQ_MOD)) return ZKFHE_ERR_ARG;
    host::Fq12 e = host::fq12_one();
    if (!p.inf && !q.inf)
        e = reference_construction ? host::final_exponentiate_reference(host::miller_loop(q, p))
                                   : host::final_exponentiate(host::miller_loop_fast(q, p));
    memcpy(out384, e.c, 384);
    return ZKFHE_OK;
}

// The proof's point encoding on its own (host only): canonical x || y (64 bytes, identity = zeros) <-> the 32 bytes
// halo2 writes.  Compression does not check the curve equation; decompression returns ZKFHE_ERR_ARG for bytes that
// are not the encoding of a curve point.
int zkfhe_point_compress(const uint8_t* xy_canon64, uint8_t* out32) {
    if (!xy_canon64 || !out32) return ZKFHE_ERR_ARG;
    uint64_t c[8];
    memcpy(c, xy_canon64, 64);
    host::Transcript::compress_point(c, c + 4, out32);
    return ZKFHE_OK;
}
int zkfhe_point_decompress(const uint8_t* in32, uint8_t* xy_canon64) {
    if (!in32 || !xy_canon64) return ZKFHE_ERR_ARG;
    Point p;
    if (!decompress_point(in32, p) || !on_curve(p)) return ZKFHE_ERR_ARG;
    memcpy(xy_canon64, p.c, 64);
    return ZKFHE_OK;
}

// [tau]_2 for the test SRS (`ParamsKZG::setup` keeps s_g2 next to the G1 powers): 128 bytes, Montgomery.
int zkfhe_srs_g2(const uint8_t* tau_mont32, uint8_t* out128) {
    if (!tau_mont32 || !out128) return ZKFHE_ERR_ARG;
    Fr t;
    memcpy(t.l, tau_mont32, 32);
    if (host::geq(t, host::FR_MOD)) return ZKFHE_ERR_ARG;
    const Fr canon = host::from_mont(t);
    const host::G2Aff q = host::g2_mul(host::g2_generator(), canon.l);
    memset(out128, 0, 128);
    if (!q.inf) {
        memcpy(out128, q.x.c0.l, 32);
        memcpy(out128 + 32, q.x.c1.l, 32);
        memcpy(out128 + 64, q.y.c0.l, 32);
        memcpy(out128 + 96, q.y.c1.l, 32);
    }
    return ZKFHE_OK;
}

// The verifying key as bytes (the reference's data/<name>.vk): layout numbers + fixed commitments.
int zkfhe_vk_export(const zkfhe_pk* pk, uint8_t* buf, size_t cap, size_t* needed) {
    if (!pk) return ZKFHE_ERR_ARG;
    const size_t need = sizeof(VkHeader) + (size_t)pk->n_fixed * 64;
    if (needed) *needed = need;
    if (!buf) return ZKFHE_OK;
    if (cap < need) return ZKFHE_ERR_ARG;
    VkHeader h{};
    memcpy(h.magic, VK_MAGIC, 8);
    h.k = pk->k; h.n_gate0 = pk->n_gate0; h.n_gate1 = pk->n_gate1; h.n_rlc = pk->n_rlc; h.n_lookup = pk->n_lookup;
    h.n_advice = pk->n_advice; h.n_perm = pk->n_perm; h.n_fixed = pk->n_fixed; h.n_chunks = pk->n_chunks;
    h.usable = pk->usable; h.lookup_bits = pk->lookup_bits; h.instances = (uint32_t)pk->instances;
    h.unusable_rows = pk->unusable_rows;
    memcpy(buf, &h, sizeof h);
    for (uint32_t f = 0; f < pk->n_fixed; f++) memcpy(buf + sizeof h + 64 * (size_t)f, pk->fixed_commitments_canon[f].data(), 64);
    return ZKFHE_OK;
}

// Returns ZKFHE_OK with *accepted = 1 / 0 (the reason for a rejection is zkfhe_last_error); error codes
// are for malformed arguments and CUDA failures only.
//   instances : n_instances canonical 32-byte little-endian scalars
//   s_g2      : [tau]_2, 128 bytes Montgomery (zkfhe_srs_g2 for the test SRS)
int zkfhe_verify(zkfhe_ctx* ctx, const uint8_t* vk, size_t vk_len, const uint8_t* instances, uint32_t n_instances,
                 const uint8_t* proof, size_t proof_len, const uint8_t* s_g2, int transcript_kind, int* accepted) {
    if (!ctx || !vk || !proof || !s_g2 || !accepted || (!instances && n_instances)) return fail(ctx, ZKFHE_ERR_ARG, "verify: null pointer");
    *accepted = 0;
    if (vk_len < sizeof(VkHeader)) return fail(ctx, ZKFHE_ERR_ARG, "verify: verifying key too short");
    VkHeader h;
    memcpy(&h, vk, sizeof h);
    if (memcmp(h.magic, VK_MAGIC, 8)) return fail(ctx, ZKFHE_ERR_ARG, "verify: not a zkfhe verifying key");
    if (vk_len != sizeof(VkHeader) + (size_t)h.n_fixed * 64) return fail(ctx, ZKFHE_ERR_ARG, "verify: verifying key length mismatch");
    if (transcript_kind != host::TRANSCRIPT_BLAKE2B && transcript_kind != host::TRANSCRIPT_POSEIDON)
        return fail(ctx, ZKFHE_ERR_ARG, "verify: unknown transcript kind %d", transcript_kind);
    const uint32_t k = h.k, n = 1u << k, usable = h.usable, n_gate = h.n_gate0 + h.n_gate1, n_sel = n_gate + h.n_rlc;
    const uint32_t fx_qgate = 0, fx_qrlc = n_gate, fx_const = n_sel, fx_table = n_sel + 1, fx_l0 = n_sel + 2, fx_sigma = n_sel + 5;
    if (k < 4 || k > 24 || h.n_fixed != fx_sigma + h.n_perm || h.n_advice != n_sel + h.n_lookup ||
        h.n_perm != h.n_advice + 2 || h.instances > usable)
        return fail(ctx, ZKFHE_ERR_ARG, "verify: inconsistent verifying key");
    // `usable` and `n_chunks` are not free parameters (and not part of the digest): the prover fixes
    // usable = n - BLINDING_FACTORS - 1 and n_chunks = ceil(n_perm / PERM_CHUNK); a key that says otherwise
    // would move l_last / l_active and the rotation w^usable without changing the digest
    if (usable != n - BLINDING_FACTORS - 1 || h.n_chunks != (h.n_perm + PERM_CHUNK - 1) / PERM_CHUNK)
        return fail(ctx, ZKFHE_ERR_ARG, "verify: verifying key carries usable_rows = %u / n_chunks = %u, the layout implies %u / %u",
                    usable, h.n_chunks, n - BLINDING_FACTORS - 1, (h.n_perm + PERM_CHUNK - 1) / PERM_CHUNK);
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<Point> fixed_cm(h.n_fixed);
    for (uint32_t f = 0; f < h.n_fixed; f++) memcpy(fixed_cm[f].c, vk + sizeof h + 64 * (size_t)f, 64);

    try {
        if (n_instances != h.instances) throw Reject{"wrong number of instances"};
        // vk digest, exactly as keygen hashes it
        Fr digest;
        {
            host::Transcript t(host::TRANSCRIPT_BLAKE2B);     // the key digest is always BLAKE2b (keygen.cu pk_finalize)
            const uint32_t shape[] = {k, h.n_gate0, h.n_gate1, h.n_rlc, h.n_lookup, h.unusable_rows, h.lookup_bits,
                                      h.instances, BLINDING_FACTORS, PERM_CHUNK};
            for (uint32_t s : shape) t.common_scalar(host::from_u64(s));
            for (const auto& cm : fixed_cm) t.common_point(cm.c, cm.c + 4);
            digest = t.squeeze();
        }
        host::Transcript tr(transcript_kind);
        size_t pos = 0;
        auto rd_point = [&]() {
            if (pos + 32 > proof_len) throw Reject{"proof truncated"};
            Point p;
            const bool ok = decompress_point(proof + pos, p);
            pos += 32;
            if (!ok || !on_curve(p)) throw Reject{"commitment is not a curve point"};
            tr.common_point(p.c, p.c + 4);
            return p;
        };
        auto rd_scalar = [&]() {
            if (pos + 32 > proof_len) throw Reject{"proof truncated"};
            Fr c;
            memcpy(c.l, proof + pos, 32);
            pos += 32;
            if (host::geq(c, host::FR_MOD)) throw Reject{"non-canonical scalar"};
            const Fr m = host::to_mont(c);
            tr.common_scalar(m);
            return m;
        };
        tr.common_scalar(digest);
        std::vector<Fr> inst(n_instances);
        for (uint32_t i = 0; i < n_instances; i++) {
            Fr c;
            memcpy(c.l, instances + 32 * (size_t)i, 32);
            if (host::geq(c, host::FR_MOD)) throw Reject{"non-canonical instance"};
            inst[i] = host::to_mont(c);
            tr.common_scalar(inst[i]);
        }
        std::vector<Point> advice_cm;
        for (uint32_t i = 0; i < h.n_gate0; i++) advice_cm.push_back(rd_point());
        const Fr gamma_rlc = tr.squeeze();
        for (uint32_t i = h.n_gate0; i < h.n_advice; i++) advice_cm.push_back(rd_point());
        tr.squeeze();                                                   // theta (single-column lookups)
        std::vector<Point> lookup_cm, zp_cm, zl_cm;
        for (uint32_t i = 0; i < 2 * h.n_lookup; i++) lookup_cm.push_back(rd_point());     // A'_l, S'_l interleaved
        const Fr beta = tr.squeeze(), gamma = tr.squeeze();
        for (uint32_t i = 0; i < h.n_chunks; i++) zp_cm.push_back(rd_point());
        for (uint32_t i = 0; i < h.n_lookup; i++) zl_cm.push_back(rd_point());
        const Point r_cm = rd_point();
        const Fr y = tr.squeeze();
        Point h_cm[3];
        for (auto& p : h_cm) p = rd_point();
        const Fr x = tr.squeeze();

        // ---- opening table, in the prover's order ------------------------------------------------------
        struct Entry { Point cm; int set; };
        std::vector<Entry> table;
        for (uint32_t c = 0; c < h.n_advice; c++)
            table.push_back({advice_cm[c], c < n_gate ? SET_0123 : c < n_gate + h.n_rlc ? SET_012 : SET_0});
        std::vector<uint32_t> fixed_idx;
        for (uint32_t f = 0; f < h.n_fixed; f++)
            if (!(f >= fx_l0 && f < fx_sigma)) fixed_idx.push_back(f);
        for (uint32_t f : fixed_idx) table.push_back({fixed_cm[f], SET_0});
        for (uint32_t l = 0; l < h.n_lookup; l++) {
            table.push_back({lookup_cm[2 * l], SET_0m1});
            table.push_back({lookup_cm[2 * l + 1], SET_0});
            table.push_back({zl_cm[l], SET_01});
        }
        for (uint32_t j = 0; j < h.n_chunks; j++) table.push_back({zp_cm[j], j + 1 < h.n_chunks ? SET_01L : SET_01});
        table.push_back({r_cm, SET_0});
        const size_t n_written = table.size();                        // h_comb(x) is recomputed, not read
        std::vector<std::vector<Fr>> evals(n_written + 1);
        for (size_t i = 0; i < n_written; i++)
            for (int r = 0; r < SET_SIZE[table[i].set]; r++) evals[i].push_back(rd_scalar());

        // ---- named evaluations ----------------------------------------------------------------------------
        size_t ei = 0;
        std::vector<std::vector<Fr>> adv(evals.begin(), evals.begin() + h.n_advice);
        ei = h.n_advice;
        std::vector<Fr> fixed(h.n_fixed, host::FR_ZERO);
        for (uint32_t f : fixed_idx) fixed[f] = evals[ei++][0];
        struct Lk { Fr ap, ap_m1, sp, z, z_w; };
        std::vector<Lk> lk(h.n_lookup);
        for (uint32_t l = 0; l < h.n_lookup; l++) {
            lk[l].ap = evals[ei][0]; lk[l].ap_m1 = evals[ei][1]; ei++;
            lk[l].sp = evals[ei][0]; ei++;
            lk[l].z = evals[ei][0]; lk[l].z_w = evals[ei][1]; ei++;
        }
        std::vector<std::vector<Fr>> zp(evals.begin() + ei, evals.begin() + ei + h.n_chunks);
        ei += h.n_chunks;
        // evals[ei] is the random polynomial's evaluation: it only enters the multi-open

        using namespace host;
        const Fr w = omega(k), xn = pow_u64(x, n), zh = sub(xn, FR_ONE);
        if (zh.is_zero()) throw Reject{"challenge x lies in the domain"};
        // Lagrange basis at x: l_i(x) = w^i (x^n - 1) / (n (x - w^i)), for the instance rows, row 0 and rows >= usable
        std::vector<uint32_t> rows;
        for (uint32_t i = 0; i < h.instances; i++) rows.push_back(i);
        if (h.instances == 0) rows.push_back(0);
        for (uint32_t i = usable; i < n; i++) rows.push_back(i);
        std::vector<Fr> wi(rows.size()), den(rows.size());
        {
            Fr cur = FR_ONE;
            uint32_t at = 0;
            for (size_t j = 0; j < rows.size(); j++) {
                if (rows[j] != at) { cur = mul(cur, pow_u64(w, rows[j] - at)); at = rows[j]; }
                wi[j] = cur;
                den[j] = sub(x, cur);
            }
            // batch inversion
            std::vector<Fr> pre(rows.size());
            Fr run = FR_ONE;
            for (size_t j = 0; j < rows.size(); j++) { pre[j] = run; run = mul(run, den[j]); }
            Fr irun = inv(run);
            for (size_t j = rows.size(); j-- > 0;) { const Fr d = den[j]; den[j] = mul(irun, pre[j]); irun = mul(irun, d); }
        }
        const Fr scale = mul(zh, inv(from_u64(n)));
        auto lag = [&](size_t j) { return mul(mul(wi[j], scale), den[j]); };
        Fr l0 = FR_ZERO, l_last = FR_ZERO, l_blind = FR_ZERO, inst_eval = FR_ZERO;
        for (size_t j = 0; j < rows.size(); j++) {
            const Fr lj = lag(j);
            if (rows[j] == 0) l0 = lj;
            if (rows[j] < h.instances) inst_eval = add(inst_eval, mul(inst[rows[j]], lj));
            if (rows[j] == usable) l_last = lj;
            if (rows[j] > usable) l_blind = add(l_blind, lj);
        }
        const Fr l_act = sub(sub(FR_ONE, l_last), l_blind);

        // ---- every identity at x, folded with powers of y (Horner, expression 0 first) ---------------------------
        Fr acc = FR_ZERO;
        auto push = [&](const Fr& e) { acc = add(mul(acc, y), e); };
        for (uint32_t c = 0; c < n_gate; c++)
            push(mul(fixed[fx_qgate + c], sub(add(adv[c][0], mul(adv[c][1], adv[c][2])), adv[c][3])));
        for (uint32_t j = 0; j < h.n_rlc; j++) {
            const auto& a = adv[n_gate + j];
            push(mul(fixed[fx_qrlc + j], sub(add(mul(a[0], gamma_rlc), a[1]), a[2])));
        }
        const uint32_t m = h.n_chunks;
        auto perm_val = [&](uint32_t c) { return c < h.n_advice ? adv[c][0] : c == h.n_advice ? fixed[fx_const] : inst_eval; };
        push(mul(l0, sub(FR_ONE, zp[0][0])));
        push(mul(l_last, sub(sqr(zp[m - 1][0]), zp[m - 1][0])));
        for (uint32_t j = 1; j < m; j++) push(mul(l0, sub(zp[j][0], zp[j - 1][2])));
        {
            const Fr delta = to_mont(FR_DELTA_CANON);
            Fr dpow = FR_ONE;                                              // delta^c
            for (uint32_t j = 0; j < m; j++) {
                Fr left = zp[j][1], right = zp[j][0];
                for (uint32_t c = j * PERM_CHUNK; c < (j + 1) * PERM_CHUNK && c < h.n_perm; c++) {
                    const Fr v = perm_val(c);
                    left = mul(left, add(add(v, mul(beta, fixed[fx_sigma + c])), gamma));
                    right = mul(right, add(add(v, mul(mul(beta, dpow), x)), gamma));
                    dpow = mul(dpow, delta);
                }
                push(mul(l_act, sub(left, right)));
            }
        }
        const uint32_t lookup_adv_base = n_gate + h.n_rlc;
        for (uint32_t l = 0; l < h.n_lookup; l++) {
            const Lk& q = lk[l];
            const Fr a = adv[lookup_adv_base + l][0], s = fixed[fx_table];
            push(mul(l0, sub(FR_ONE, q.z)));
            push(mul(l_last, sub(sqr(q.z), q.z)));
            push(mul(l_act, sub(mul(mul(q.z_w, add(q.ap, beta)), add(q.sp, gamma)), mul(mul(q.z, add(a, beta)), add(s, gamma)))));
            push(mul(l0, sub(q.ap, q.sp)));
            push(mul(l_act, mul(sub(q.ap, q.sp), sub(q.ap, q.ap_m1))));
        }
        evals[n_written] = {mul(acc, inv(zh))};                           // h_comb(x)

        // ---- SHPLONK -------------------------------------------------------------------------------------------------
        const Fr yq = tr.squeeze(), v = tr.squeeze();
        const Point W = rd_point();
        const Fr u = tr.squeeze();
        const Point Wp = rd_point();
        if (pos != proof_len) throw Reject{"trailing bytes in proof"};
        const Fr winv = inv(w);
        const Fr pts[6] = {mul(x, winv), x, mul(x, w), mul(x, sqr(w)), mul(x, mul(w, sqr(w))), mul(x, pow_u64(w, usable))};
        Fr zt = FR_ONE;
        for (const Fr& t : pts) zt = mul(zt, sub(u, t));
        std::vector<Point> mpts;
        std::vector<Fr> mscal;
        Fr g_coeff = FR_ZERO, vpow = FR_ONE;
        // the combined quotient commitment h_0 + x^n h_1 + x^2n h_2 is the last member of set 0
        for (int s = 0; s < 6; s++) {
            const int sz = SET_SIZE[s];
            Fr T[4];
            for (int i = 0; i < sz; i++) T[i] = pts[point_index(SET_ROTS[s][i])];
            Fr comb[4] = {FR_ZERO, FR_ZERO, FR_ZERO, FR_ZERO};
            std::vector<std::pair<size_t, Fr>> members;                   // (table index, yq^j)
            Fr yp = FR_ONE;
            for (size_t i = 0; i <= n_written; i++) {
                const int set = i < n_written ? table[i].set : SET_0;
                if (set != s) continue;
                for (int r = 0; r < sz; r++) comb[r] = add(comb[r], mul(yp, evals[i][r]));
                members.push_back({i, yp});
                yp = mul(yp, yq);
            }
            if (!members.empty()) {
                Fr r_u = FR_ZERO;                                          // r_s(u): Lagrange interpolation at u
                for (int i = 0; i < sz; i++) {
                    Fr num = FR_ONE, dn = FR_ONE;
                    for (int q = 0; q < sz; q++)
                        if (q != i) { num = mul(num, sub(u, T[q])); dn = mul(dn, sub(T[i], T[q])); }
                    r_u = add(r_u, mul(mul(comb[i], num), inv(dn)));
                }
                Fr z_s = FR_ONE;
                for (int i = 0; i < sz; i++) z_s = mul(z_s, sub(u, T[i]));
                const Fr zc = mul(zt, inv(z_s)), vz = mul(vpow, zc);
                for (const auto& mb : members) {
                    const Fr sc = mul(vz, mb.second);
                    if (mb.first < n_written) {
                        mpts.push_back(table[mb.first].cm);
                        mscal.push_back(sc);
                    } else {
                        mpts.push_back(h_cm[0]); mscal.push_back(sc);
                        mpts.push_back(h_cm[1]); mscal.push_back(mul(sc, xn));
                        mpts.push_back(h_cm[2]); mscal.push_back(mul(sc, sqr(xn)));
                    }
                }
                g_coeff = sub(g_coeff, mul(vz, r_u));
            }
            vpow = mul(vpow, v);
        }
        Point gen{};
        gen.c[0] = 1;
        gen.c[4] = 2;
        mpts.push_back(gen); mscal.push_back(g_coeff);
        mpts.push_back(W); mscal.push_back(neg(zt));
        mpts.push_back(Wp); mscal.push_back(u);                           // lhs = F + u W'

        // ---- one MSM on the GPU, one pairing check on the host --------------------------------------------------
        std::vector<g1_affine> dev_pts(mpts.size());
        std::vector<fr_t> dev_sc(mpts.size());
        for (size_t i = 0; i < mpts.size(); i++) {
            dev_pts[i] = to_device_point(mpts[i]);
            memcpy(dev_sc[i].v, mscal[i].l, 32);
        }
        g1_affine lhs;
        ZK_TRY(msm_variable_base(ctx, dev_pts.data(), dev_sc.data(), (uint32_t)dev_pts.size(), &lhs));
        host::G1Aff pp[2];
        host::G2Aff qq[2];
        memcpy(pp[0].x.l, lhs.x.v, 32);
        memcpy(pp[0].y.l, lhs.y.v, 32);
        pp[0].inf = pp[0].x.is_zero() && pp[0].y.is_zero();
        const g1_affine wp_m = to_device_point(Wp);
        memcpy(pp[1].x.l, wp_m.x.v, 32);
        memcpy(pp[1].y.l, wp_m.y.v, 32);
        pp[1].inf = pp[1].x.is_zero() && pp[1].y.is_zero();
        pp[1].y = host::fq_neg(pp[1].y);
        qq[0] = host::g2_generator();
        memcpy(qq[1].x.c0.l, s_g2, 32);
        memcpy(qq[1].x.c1.l, s_g2 + 32, 32);
        memcpy(qq[1].y.c0.l, s_g2 + 64, 32);
        memcpy(qq[1].y.c1.l, s_g2 + 96, 32);
        qq[1].inf = false;
        for (const host::Fq* c : {&qq[1].x.c0, &qq[1].x.c1, &qq[1].y.c0, &qq[1].y.c1})
            if (host::fq_geq(*c, host::FQ_MOD)) return fail(ctx, ZKFHE_ERR_ARG, "verify: s_g2 is not reduced");
        if (!host::pairing_product_is_one(pp, qq, 2)) throw Reject{"KZG opening check failed (pairing)"};
        *accepted = 1;
        return ZKFHE_OK;
    } catch (const Reject& r) {
        ctx->err = "verify: rejected: " + r.why;
        *accepted = 0;
        return ZKFHE_OK;
    }
}

}  // extern "C"
