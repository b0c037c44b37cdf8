// keygen for the BFV circuit: from the structure recorded by the witness kernels (run on the
// keygen input, data/bfv/bfv_empty.in in the reference, README.md:28-38) to the proving key.
//
// Mirrors what `cargo run --example bfv -- ... keygen` does through halo2-base / axiom-eth /
// halo2 [UPSTREAM, un-vendored]: auto-configure the column counts for k, record the break points
// (configs/bfv.json -- reproduced exactly, tests/test_gpu_prover.py), build the fixed columns and
// the permutation, commit them.  The Keccak sub-circuit the reference carries unused is not built
// (SURVEY.md §7 H3); `unusable_rows` is kept so that the break points match.
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include "prover.cuh"
#include "witness.cuh"
#include "witness_types.cuh"

using namespace zkfhe;

namespace zkfhe {

// halo2-base GateThreadBuilder::assign_all / axiom-eth assign_rlc: cut a flat context into columns.
static void cut_context(const uint8_t* flags, uint64_t ncells, uint32_t max_rows, uint32_t gate_span, ColumnCut& cut) {
    cut.break_points.clear();
    cut.start.clear();
    cut.rows.clear();
    if (ncells == 0) return;
    cut.start.push_back(0);
    uint32_t row = 0;
    for (uint64_t i = 0; i < ncells; i++) {
        bool q = flags[i] & META_SELECTOR;
        if ((q && row + gate_span > max_rows) || row >= max_rows - 1) {
            cut.break_points.push_back(row);
            cut.rows.push_back(row + 1);
            cut.start.push_back(i);      // the break cell is duplicated at row 0 of the next column
            row = 0;
        }
        row++;
    }
    cut.rows.push_back(row);
}

struct UnionFind {
    std::vector<uint32_t> p;
    explicit UnionFind(size_t n) : p(n) { for (size_t i = 0; i < n; i++) p[i] = (uint32_t)i; }
    uint32_t find(uint32_t x) {
        while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
        return x;
    }
    void unite(uint32_t a, uint32_t b) {
        a = find(a); b = find(b);
        if (a != b) p[a > b ? a : b] = a > b ? b : a;
    }
};

__global__ void k_build_fixed(fr_t* fixed, uint32_t n, uint32_t usable, uint32_t n_sel, const uint8_t* sel_mask,
                              uint32_t fx_const, const fr_t* const_vals, uint32_t n_const, uint32_t fx_table,
                              uint32_t lookup_bits, uint32_t fx_l0, uint32_t fx_sigma, uint32_t n_perm,
                              const uint32_t* sig_col, const uint32_t* sig_row, const fr_t* delta_pow, const fr_t* tw) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y;
    if (row >= n) return;
    fr_t v = fe_zero<FR>();
    if (col < n_sel) {
        if (sel_mask[(size_t)col * n + row]) v = fe_one<FR>();
    } else if (col == fx_const) {
        if (row < n_const) v = fe_load(const_vals + row);
    } else if (col == fx_table) {
        if (row < (1u << lookup_bits)) v = mont_u64(row);
    } else if (col == fx_l0) {
        if (row == 0) v = fe_one<FR>();
    } else if (col == fx_l0 + 1) {
        if (row == usable) v = fe_one<FR>();
    } else if (col == fx_l0 + 2) {
        if (row < usable) v = fe_one<FR>();
    } else if (col >= fx_sigma && col < fx_sigma + n_perm) {
        size_t p = (size_t)(col - fx_sigma) * n + row;
        v = mul(fe_load(delta_pow + sig_col[p]), fe_load(tw + sig_row[p]));
    }
    fe_store(fixed + (size_t)col * n + row, v);
}

static std::string json_list(const std::vector<uint32_t>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) s += (i ? "," : "") + std::to_string(v[i]);
    return s + "]";
}

static std::string hex32(const uint64_t l[4]) {
    char b[67];
    snprintf(b, sizeof b, "0x%016llx%016llx%016llx%016llx", (unsigned long long)l[3], (unsigned long long)l[2],
             (unsigned long long)l[1], (unsigned long long)l[0]);
    return b;
}


// Shared tail of keygen and of zkfhe_pk_import: everything that is a function of the layout numbers, the Lagrange
// form of the fixed columns and their commitments -- coefficient and extended forms, the pinning JSON, the vk digest.
static int pk_finalize(zkfhe_pk* pk) {
    zkfhe_ctx* ctx = pk->ctx;
    const uint32_t n = pk->n, k = pk->k;
    const size_t fbytes = (size_t)pk->n_fixed * n * sizeof(fr_t);
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->fixed_coeff, pk->fixed_lagrange, fbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ZK_TRY(ntt_run(ctx, pk->fixed_coeff, n, n, pk->fixed_coeff, n, k, pk->n_fixed, 1, 0));
    ZK_TRY(ntt_run(ctx, pk->fixed_coeff, n, n, pk->fixed_ext, (uint64_t)n << EXT_SHIFT, k + EXT_SHIFT, pk->n_fixed, 0, 1));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));

    // ---- pinning (configs/<name>.json schema of the reference) and vk digest -----------------------------
    pk->pinning_json =
        "{\"params\":{\"degree\":" + std::to_string(k) + ",\"num_rlc_columns\":" + std::to_string(pk->n_rlc) +
        ",\"num_range_advice\":[" + std::to_string(pk->n_gate0) + "," + std::to_string(pk->n_gate1) + ",0]" +
        ",\"num_lookup_advice\":[0," + std::to_string(pk->n_lookup) + ",0],\"num_fixed\":1,\"unusable_rows\":" +
        std::to_string(pk->unusable_rows) + ",\"keccak_rows_per_round\":50,\"lookup_bits\":" + std::to_string(pk->lookup_bits) +
        "},\"break_points\":{\"gate\":[" + json_list(pk->cut[0].break_points) + "," + json_list(pk->cut[1].break_points) +
        ",[]],\"rlc\":" + json_list(pk->cut[2].break_points) + "}}";
    host::Transcript t(host::TRANSCRIPT_BLAKE2B);
    const uint32_t shape[] = {k, pk->n_gate0, pk->n_gate1, pk->n_rlc, pk->n_lookup, pk->unusable_rows, pk->lookup_bits,
                              (uint32_t)pk->instances, BLINDING_FACTORS, PERM_CHUNK};
    for (uint32_t s : shape) t.common_scalar(host::from_u64(s));
    for (const auto& cm : pk->fixed_commitments_canon) t.common_point(cm.data(), cm.data() + 4);
    pk->vk_digest = t.squeeze();
    return ZKFHE_OK;
}

// layout numbers that follow from the column cuts and the lookup count
static void pk_derive_layout(zkfhe_pk* pk) {
    pk->n_gate0 = (uint32_t)pk->cut[0].rows.size();
    pk->n_gate1 = (uint32_t)pk->cut[1].rows.size();
    pk->n_rlc = (uint32_t)pk->cut[2].rows.size();
    pk->n_lookup = (uint32_t)((pk->lookups + pk->max_rows - 1) / pk->max_rows);
    pk->n_advice = pk->n_gate0 + pk->n_gate1 + pk->n_rlc + pk->n_lookup;
    pk->n_perm = pk->n_advice + 2;
    pk->n_chunks = (pk->n_perm + PERM_CHUNK - 1) / PERM_CHUNK;
    const uint32_t n_gate = pk->n_gate0 + pk->n_gate1, n_sel = n_gate + pk->n_rlc;
    pk->fx_qgate = 0; pk->fx_qrlc = n_gate; pk->fx_const = n_sel; pk->fx_table = n_sel + 1;
    pk->fx_l0 = n_sel + 2; pk->fx_llast = n_sel + 3; pk->fx_lactive = n_sel + 4; pk->fx_sigma = n_sel + 5;
    pk->n_fixed = pk->fx_sigma + pk->n_perm;
}

}  // namespace zkfhe

extern "C" {

void zkfhe_pk_free(zkfhe_pk* pk) {
    if (!pk) return;
    cudaSetDevice(pk->ctx->device);
    cudaStreamSynchronize(pk->ctx->stream);
    if (pk->fixed_lagrange) cudaFree(pk->fixed_lagrange);
    if (pk->fixed_coeff) cudaFree(pk->fixed_coeff);
    if (pk->fixed_ext) cudaFree(pk->fixed_ext);
    if (pk->delta_pow) cudaFree(pk->delta_pow);
    delete pk;
}

int zkfhe_keygen(zkfhe_witness* w, uint32_t k, uint32_t unusable_rows, zkfhe_pk** out) {
    if (!w || !out) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = w->ctx;
    if (!w->record) return fail(ctx, ZKFHE_ERR_STATE, "keygen: the witness must be built in recording mode");
    if (ctx->srs_k != k) return fail(ctx, ZKFHE_ERR_STATE, "keygen: SRS for k=%u is not loaded (have k=%u)", k, ctx->srs_k);
    if (k < 4 || k > 20) return fail(ctx, ZKFHE_ERR_ARG, "keygen: k=%u out of range", k);
    const uint32_t n = 1u << k;
    if (unusable_rows < BLINDING_FACTORS + 3 || unusable_rows >= n / 2)
        return fail(ctx, ZKFHE_ERR_ARG, "keygen: unusable_rows=%u must be in [%u, n/2)", unusable_rows, BLINDING_FACTORS + 3);
    if ((1u << w->lookup_bits) > n - unusable_rows) return fail(ctx, ZKFHE_ERR_ARG, "keygen: lookup table does not fit");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    zkfhe_pk* pk = new (std::nothrow) zkfhe_pk();
    if (!pk) return fail(ctx, ZKFHE_ERR_CUDA, "out of host memory");
    struct Guard { zkfhe_pk* p; ~Guard() { if (p) zkfhe_pk_free(p); } } guard{pk};
    pk->ctx = ctx;
    pk->k = k; pk->n = n; pk->unusable_rows = unusable_rows; pk->lookup_bits = w->lookup_bits;
    pk->max_rows = n - unusable_rows;
    pk->usable = n - BLINDING_FACTORS - 1;

    // ---- download the recorded structure ------------------------------------------------------
    std::vector<uint8_t> flags[3];
    std::vector<uint64_t> copy[3];
    std::vector<host::Fr> vals[3];
    for (int c = 0; c < 3; c++) {
        const size_t m = w->adv[c].size;
        pk->cells[c] = m;
        flags[c].resize(m);
        copy[c].resize(m);
        vals[c].resize(m);
        if (!m) continue;
        ZK_CUDA(ctx, cudaMemcpyAsync(flags[c].data(), w->flags[c].p, m, cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(ctx, cudaMemcpyAsync(copy[c].data(), w->copy[c].p, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        ZK_CUDA(ctx, cudaMemcpyAsync(vals[c].data(), w->adv[c].p, m * 32, cudaMemcpyDeviceToHost, ctx->stream));
    }
    std::vector<uint64_t> lk_src;
    for (int c = 0; c < 3; c++) {
        size_t m = w->lk[c].size, off = lk_src.size();
        lk_src.resize(off + m);
        if (m) ZK_CUDA(ctx, cudaMemcpyAsync(lk_src.data() + off, w->lk_src[c].p, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    pk->lookups = lk_src.size();
    pk->instances = w->make_public.size();
    for (auto& c : w->make_public) pk->public_cells.push_back(cell_id(c.ctx_id, c.offset));
    if (pk->instances > pk->usable) return fail(ctx, ZKFHE_ERR_ARG, "keygen: %llu instances do not fit %u rows",
                                                (unsigned long long)pk->instances, pk->usable);

    // ---- column layout -------------------------------------------------------------------------
    cut_context(flags[0].data(), flags[0].size(), pk->max_rows, 4, pk->cut[0]);
    cut_context(flags[1].data(), flags[1].size(), pk->max_rows, 4, pk->cut[1]);
    cut_context(flags[2].data(), flags[2].size(), pk->max_rows, 3, pk->cut[2]);
    pk_derive_layout(pk);
    const uint32_t n_gate = pk->n_gate0 + pk->n_gate1, n_sel = n_gate + pk->n_rlc;
    const uint32_t col_base[3] = {0, pk->n_gate0, n_gate};
    const uint32_t lookup_base = n_gate + pk->n_rlc;
    const uint32_t perm_const = pk->n_advice, perm_inst = pk->n_advice + 1;

    auto locate = [&](uint64_t id, uint32_t& col, uint32_t& row) {   // cell id -> (advice column, row); break cells -> row 0 of the later column
        const uint32_t c = cell_ctx(id);
        const uint64_t off = cell_off(id);
        const auto& st = pk->cut[c].start;
        size_t j = std::upper_bound(st.begin(), st.end(), off) - st.begin() - 1;
        col = col_base[c] + (uint32_t)j;
        row = (uint32_t)(off - st[j]);
    };

    // ---- selectors, constants, permutation ---------------------------------------------------------
    std::vector<uint8_t> sel_mask((size_t)n_sel * n, 0);
    std::map<std::array<uint64_t, 4>, uint32_t> const_index;
    std::vector<host::Fr> const_vals;
    auto const_row = [&](const host::Fr& v) {
        std::array<uint64_t, 4> key = {v.l[0], v.l[1], v.l[2], v.l[3]};
        auto it = const_index.find(key);
        if (it != const_index.end()) return it->second;
        uint32_t r = (uint32_t)const_vals.size();
        const_index[key] = r;
        const_vals.push_back(v);
        return r;
    };
    UnionFind uf((size_t)pk->n_perm * n);
    auto pos = [&](uint32_t col, uint32_t row) { return (uint32_t)((size_t)col * n + row); };
    for (int c = 0; c < 3; c++) {
        const auto& cut = pk->cut[c];
        for (size_t j = 0; j + 1 < cut.start.size(); j++)      // break duplicates
            uf.unite(pos(col_base[c] + (uint32_t)j, cut.break_points[j]), pos(col_base[c] + (uint32_t)j + 1, 0));
        for (uint64_t off = 0; off < flags[c].size(); off++) {
            const uint8_t f = flags[c][off];
            if (f & META_COPY_CONFLICT)
                return fail(ctx, ZKFHE_ERR_STATE, "keygen: cell %llu of context %d was given two different copy sources "
                            "(an equality constraint would be dropped)", (unsigned long long)off, c);
            uint32_t col, row;
            locate(cell_id(c, off), col, row);
            if (f & META_SELECTOR) {
                const uint32_t sc = c == 2 ? pk->fx_qrlc + (col - col_base[2]) : pk->fx_qgate + col;
                sel_mask[(size_t)sc * n + row] = 1;
            }
            if (f & META_CONSTANT) uf.unite(pos(col, row), pos(perm_const, const_row(vals[c][off])));
            if (f & META_ASSERT_ZERO) uf.unite(pos(col, row), pos(perm_const, const_row(host::FR_ZERO)));
            if (f & META_ASSERT_ONE) uf.unite(pos(col, row), pos(perm_const, const_row(host::FR_ONE)));
            if (copy[c][off] != CELL_NONE) {
                uint32_t c2, r2;
                locate(copy[c][off], c2, r2);
                uf.unite(pos(col, row), pos(c2, r2));
            }
        }
    }
    if (const_vals.size() > pk->usable) return fail(ctx, ZKFHE_ERR_ARG, "keygen: too many distinct constants");
    for (uint64_t i = 0; i < lk_src.size(); i++) {               // lookup-advice cells copy their source cell
        uint32_t c2, r2;
        locate(lk_src[i], c2, r2);
        uf.unite(pos(lookup_base + (uint32_t)(i / pk->max_rows), (uint32_t)(i % pk->max_rows)), pos(c2, r2));
    }
    for (uint64_t i = 0; i < pk->public_cells.size(); i++) {     // instance row i = i-th public cell
        uint32_t c2, r2;
        locate(pk->public_cells[i], c2, r2);
        uf.unite(pos(perm_inst, (uint32_t)i), pos(c2, r2));
    }
    // cycles: sigma(p) = next member of p's class in increasing position order
    const size_t npos = (size_t)pk->n_perm * n;
    std::vector<uint32_t> sig(npos), first(npos, 0xffffffffu), last(npos, 0xffffffffu);
    for (size_t p = 0; p < npos; p++) {
        uint32_t r = uf.find((uint32_t)p);
        if (last[r] == 0xffffffffu) first[r] = (uint32_t)p; else sig[last[r]] = (uint32_t)p;
        last[r] = (uint32_t)p;
    }
    for (size_t p = 0; p < npos; p++)
        if (uf.p[p] == p) sig[last[p]] = first[p];
    std::vector<uint32_t> sig_col(npos), sig_row(npos);
    for (size_t p = 0; p < npos; p++) { sig_col[p] = sig[p] / n; sig_row[p] = sig[p] % n; }

    // ---- fixed columns on the device ---------------------------------------------------------------
    const size_t fbytes = (size_t)pk->n_fixed * n * sizeof(fr_t);
    ZK_CUDA(ctx, cudaMalloc(&pk->fixed_lagrange, fbytes));
    ZK_CUDA(ctx, cudaMalloc(&pk->fixed_coeff, fbytes));
    ZK_CUDA(ctx, cudaMalloc(&pk->fixed_ext, fbytes << EXT_SHIFT));
    ZK_CUDA(ctx, cudaMalloc(&pk->delta_pow, (size_t)pk->n_perm * sizeof(fr_t)));
    std::vector<host::Fr> dpow(pk->n_perm);
    {
        host::Fr d = host::to_mont(host::FR_DELTA_CANON), acc = host::FR_ONE;
        for (uint32_t c = 0; c < pk->n_perm; c++) { dpow[c] = acc; acc = host::mul(acc, d); }
    }
    uint8_t* d_mask;
    uint32_t *d_sc, *d_sr;
    fr_t* d_const;
    ZK_TRY(ws_get(ctx, "kg_mask", sel_mask.size(), (void**)&d_mask));
    ZK_TRY(ws_get(ctx, "kg_sigc", npos * 4, (void**)&d_sc));
    ZK_TRY(ws_get(ctx, "kg_sigr", npos * 4, (void**)&d_sr));
    ZK_TRY(ws_get(ctx, "kg_const", (const_vals.size() + 1) * 32, (void**)&d_const));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_mask, sel_mask.data(), sel_mask.size(), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_sc, sig_col.data(), npos * 4, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_sr, sig_row.data(), npos * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (!const_vals.empty())
        ZK_CUDA(ctx, cudaMemcpyAsync(d_const, const_vals.data(), const_vals.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->delta_pow, dpow.data(), dpow.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    NttDomain* dom;
    ZK_TRY(ntt_domain(ctx, k, &dom));
    dim3 grid((n + 255) / 256, pk->n_fixed);
    k_build_fixed<<<grid, 256, 0, ctx->stream>>>(pk->fixed_lagrange, n, pk->usable, n_sel, d_mask, pk->fx_const, d_const,
                                                 (uint32_t)const_vals.size(), pk->fx_table, pk->lookup_bits, pk->fx_l0,
                                                 pk->fx_sigma, pk->n_perm, d_sc, d_sr, pk->delta_pow, dom->tw_fwd);
    ZK_CHECK_LAUNCH(ctx);
    // commitments (the verifying key), coefficient and extended forms
    g1_affine* d_comm;
    ZK_TRY(ws_get(ctx, "kg_comm", (size_t)pk->n_fixed * sizeof(g1_affine), (void**)&d_comm));
    ZK_TRY(msm_run(ctx, pk->fixed_lagrange, n, k, pk->n_fixed, 1, d_comm));
    pk->fixed_commitments.resize(pk->n_fixed);
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->fixed_commitments.data(), d_comm, (size_t)pk->n_fixed * sizeof(g1_affine),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    ZK_TRY(points_to_canonical(ctx, d_comm, pk->n_fixed));          // canonical coordinates for the transcript
    pk->fixed_commitments_canon.resize(pk->n_fixed);
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->fixed_commitments_canon.data(), d_comm, (size_t)pk->n_fixed * sizeof(g1_affine),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    ZK_TRY(pk_finalize(pk));
    guard.p = nullptr;
    *out = pk;
    return ZKFHE_OK;
}


// ---- data/<name>.pk: the proving key as a file (README.md:38 -- keygen once, prove many) -----------------------
// "ZKFHEPK1" | u32 k, unusable_rows, lookup_bits, reserved | u64 cells[3], lookups, instances |
// 3 x { u32 columns; u32 break_points[columns-1]; u64 start[columns]; u32 rows[columns] } | u64 public_cells[instances] |
// fixed commitments (n_fixed x 64 B, Montgomery affine) | fixed columns in Lagrange form (n_fixed x n x 32 B, Montgomery).
// Everything else (coefficient / extended forms, sigma-independent tables, the digest) is recomputed on import.
static const char PK_MAGIC[8] = {'Z', 'K', 'F', 'H', 'E', 'P', 'K', '1'};

int zkfhe_pk_export(const zkfhe_pk* pk, uint8_t* buf, size_t cap, size_t* needed) {
    if (!pk) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = pk->ctx;
    size_t need = 8 + 16 + 40;
    for (int c = 0; c < 3; c++) {
        const size_t cols = pk->cut[c].rows.size();
        need += 4 + (cols ? (cols - 1) * 4 : 0) + cols * 8 + cols * 4;
    }
    need += pk->public_cells.size() * 8 + (size_t)pk->n_fixed * 64 + (size_t)pk->n_fixed * pk->n * 32;
    if (needed) *needed = need;
    if (!buf) return ZKFHE_OK;
    if (cap < need) return fail(ctx, ZKFHE_ERR_ARG, "pk_export: buffer of %zu bytes, %zu needed", cap, need);
    uint8_t* p = buf;
    auto put = [&](const void* src, size_t len) { memcpy(p, src, len); p += len; };
    put(PK_MAGIC, 8);
    const uint32_t head[4] = {pk->k, pk->unusable_rows, pk->lookup_bits, 0};
    put(head, 16);
    const uint64_t counts[5] = {pk->cells[0], pk->cells[1], pk->cells[2], pk->lookups, pk->instances};
    put(counts, 40);
    for (int c = 0; c < 3; c++) {
        const uint32_t cols = (uint32_t)pk->cut[c].rows.size();
        put(&cols, 4);
        if (cols) put(pk->cut[c].break_points.data(), (size_t)(cols - 1) * 4);
        put(pk->cut[c].start.data(), (size_t)cols * 8);
        put(pk->cut[c].rows.data(), (size_t)cols * 4);
    }
    put(pk->public_cells.data(), pk->public_cells.size() * 8);
    put(pk->fixed_commitments.data(), (size_t)pk->n_fixed * 64);
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    ZK_CUDA(ctx, cudaMemcpyAsync(p, pk->fixed_lagrange, (size_t)pk->n_fixed * pk->n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_pk_import(zkfhe_ctx* ctx, const uint8_t* buf, size_t len, zkfhe_pk** out) {
    if (!ctx || !buf || !out) return ZKFHE_ERR_ARG;
    const uint8_t* p = buf;
    const uint8_t* end = buf + len;
    bool short_read = false;
    auto get = [&](void* dst, size_t m) {
        if ((size_t)(end - p) < m) { short_read = true; return; }
        memcpy(dst, p, m);
        p += m;
    };
    char magic[8] = {0};
    uint32_t head[4] = {0, 0, 0, 0};
    uint64_t counts[5] = {0, 0, 0, 0, 0};
    get(magic, 8); get(head, 16); get(counts, 40);
    if (short_read || memcmp(magic, PK_MAGIC, 8)) return fail(ctx, ZKFHE_ERR_ARG, "pk_import: not a zkfhe proving key");
    const uint32_t k = head[0];
    if (k < 4 || k > 20) return fail(ctx, ZKFHE_ERR_ARG, "pk_import: k=%u out of range", k);
    if (ctx->srs_k != k) return fail(ctx, ZKFHE_ERR_STATE, "pk_import: SRS for k=%u is not loaded (have k=%u)", k, ctx->srs_k);
    const uint32_t n = 1u << k;
    if (head[1] < BLINDING_FACTORS + 3 || head[1] >= n / 2 || head[2] == 0 || head[2] > 12)
        return fail(ctx, ZKFHE_ERR_ARG, "pk_import: inconsistent header");
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    zkfhe_pk* pk = new (std::nothrow) zkfhe_pk();
    if (!pk) return fail(ctx, ZKFHE_ERR_CUDA, "out of host memory");
    struct Guard { zkfhe_pk* p; ~Guard() { if (p) zkfhe_pk_free(p); } } guard{pk};
    pk->ctx = ctx;
    pk->k = k; pk->n = n; pk->unusable_rows = head[1]; pk->lookup_bits = head[2];
    pk->max_rows = n - pk->unusable_rows;
    pk->usable = n - BLINDING_FACTORS - 1;
    for (int c = 0; c < 3; c++) pk->cells[c] = counts[c];
    pk->lookups = counts[3];
    pk->instances = counts[4];
    for (int c = 0; c < 3; c++) {
        uint32_t cols = 0;
        get(&cols, 4);
        if (short_read || cols > 65536) return fail(ctx, ZKFHE_ERR_ARG, "pk_import: truncated or corrupt column table");
        pk->cut[c].break_points.resize(cols ? cols - 1 : 0);
        pk->cut[c].start.resize(cols);
        pk->cut[c].rows.resize(cols);
        if (cols) get(pk->cut[c].break_points.data(), (size_t)(cols - 1) * 4);
        get(pk->cut[c].start.data(), (size_t)cols * 8);
        get(pk->cut[c].rows.data(), (size_t)cols * 4);
        uint64_t covered = 0;
        for (uint32_t j = 0; j < cols && !short_read; j++) {
            if (pk->cut[c].rows[j] > pk->max_rows || pk->cut[c].start[j] + pk->cut[c].rows[j] > pk->cells[c])
                return fail(ctx, ZKFHE_ERR_ARG, "pk_import: column %u of context %d lies outside its context", j, c);
            covered = pk->cut[c].start[j] + pk->cut[c].rows[j];
        }
        if (!short_read && covered != pk->cells[c]) return fail(ctx, ZKFHE_ERR_ARG, "pk_import: the columns of context %d do not cover it", c);
    }
    if (short_read || pk->instances > pk->usable) return fail(ctx, ZKFHE_ERR_ARG, "pk_import: truncated or corrupt key");
    pk->public_cells.resize(pk->instances);
    get(pk->public_cells.data(), pk->instances * 8);
    for (uint64_t id : pk->public_cells)
        if (id == CELL_NONE || cell_ctx(id) > 2 || cell_off(id) >= pk->cells[cell_ctx(id)])
            return fail(ctx, ZKFHE_ERR_ARG, "pk_import: a public cell lies outside the witness");
    pk_derive_layout(pk);
    pk->fixed_commitments.resize(pk->n_fixed);
    get(pk->fixed_commitments.data(), (size_t)pk->n_fixed * 64);
    const size_t fbytes = (size_t)pk->n_fixed * n * sizeof(fr_t);
    if (short_read || (size_t)(end - p) != fbytes) return fail(ctx, ZKFHE_ERR_ARG, "pk_import: length mismatch (%zu bytes of fixed columns expected)", fbytes);
    ZK_CUDA(ctx, cudaMalloc(&pk->fixed_lagrange, fbytes));
    ZK_CUDA(ctx, cudaMalloc(&pk->fixed_coeff, fbytes));
    ZK_CUDA(ctx, cudaMalloc(&pk->fixed_ext, fbytes << EXT_SHIFT));
    ZK_CUDA(ctx, cudaMalloc(&pk->delta_pow, (size_t)pk->n_perm * sizeof(fr_t)));
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->fixed_lagrange, p, fbytes, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<host::Fr> dpow(pk->n_perm);
    {
        host::Fr d = host::to_mont(host::FR_DELTA_CANON), acc = host::FR_ONE;
        for (uint32_t c = 0; c < pk->n_perm; c++) { dpow[c] = acc; acc = host::mul(acc, d); }
    }
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->delta_pow, dpow.data(), dpow.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    g1_affine* d_comm;
    ZK_TRY(ws_get(ctx, "kg_comm", (size_t)pk->n_fixed * sizeof(g1_affine), (void**)&d_comm));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_comm, pk->fixed_commitments.data(), (size_t)pk->n_fixed * 64, cudaMemcpyHostToDevice, ctx->stream));
    ZK_TRY(points_to_canonical(ctx, d_comm, pk->n_fixed));
    pk->fixed_commitments_canon.resize(pk->n_fixed);
    ZK_CUDA(ctx, cudaMemcpyAsync(pk->fixed_commitments_canon.data(), d_comm, (size_t)pk->n_fixed * 64, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_TRY(pk_finalize(pk));      // synchronises: the host buffers above are complete
    guard.p = nullptr;
    *out = pk;
    return ZKFHE_OK;
}

int zkfhe_pk_pinning_json(const zkfhe_pk* pk, char* buf, size_t cap, size_t* needed) {
    if (!pk) return ZKFHE_ERR_ARG;
    if (needed) *needed = pk->pinning_json.size() + 1;
    if (buf && cap) {
        size_t m = pk->pinning_json.size() < cap - 1 ? pk->pinning_json.size() : cap - 1;
        memcpy(buf, pk->pinning_json.data(), m);
        buf[m] = 0;
    }
    return ZKFHE_OK;
}

int zkfhe_pk_info(const zkfhe_pk* pk, uint32_t out[16]) {
    if (!pk || !out) return ZKFHE_ERR_ARG;
    const uint32_t v[16] = {pk->k, pk->n_gate0, pk->n_gate1, pk->n_rlc, pk->n_lookup, pk->n_advice, pk->n_perm, pk->n_fixed,
                            pk->n_chunks, pk->usable, pk->max_rows, pk->lookup_bits, (uint32_t)pk->instances,
                            pk->fx_sigma, pk->fx_const, pk->fx_table};
    memcpy(out, v, sizeof v);
    return ZKFHE_OK;
}

int zkfhe_pk_download_fixed(const zkfhe_pk* pk, uint32_t index, uint32_t form, uint8_t* h_out) {
    if (!pk || !h_out || index >= pk->n_fixed || form > 2) return ZKFHE_ERR_ARG;
    zkfhe_ctx* ctx = pk->ctx;
    ZK_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = pk->n;
    const fr_t* src = form == 0 ? pk->fixed_lagrange + index * n : form == 1 ? pk->fixed_coeff + index * n
                                                                           : pk->fixed_ext + ((index * n) << EXT_SHIFT);
    ZK_CUDA(ctx, cudaMemcpyAsync(h_out, src, (form == 2 ? n << EXT_SHIFT : n) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    return ZKFHE_OK;
}

int zkfhe_pk_fixed_commitments(const zkfhe_pk* pk, uint8_t* h_out) {
    if (!pk || !h_out) return ZKFHE_ERR_ARG;
    memcpy(h_out, pk->fixed_commitments.data(), pk->fixed_commitments.size() * sizeof(g1_affine));
    return ZKFHE_OK;
}

}  // extern "C"
