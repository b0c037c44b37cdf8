// Batched Pippenger MSM over BN254 G1 for sm_100a (stage (2) of the prove path).
//
// Replaces halo2-axiom `ParamsKZG::{commit, commit_lagrange}` ->
// `arithmetic::best_multiexp` [UPSTREAM, un-vendored; SURVEY.md §8 a19]: for each
// of `batch` columns, out = sum_i s_i * P_i over the n = 2^k points of one SRS
// basis.  The affine result is mathematically unique, so parity with the CPU
// prover is equality with oracle/curve.py `msm_naive`.
//
// Design (B200-first, not the CPU algorithm):
//  * The SRS is fixed for the life of the prover and HBM is large, so every base
//    point is expanded once into W = ceil(255/c) fixed-base multiples
//    2^(c*w) * P_i (affine).  All W windows of a scalar then fall into ONE set of
//    2^(c-1) signed-digit buckets per column: one bucket reduction per column
//    instead of W.
//  * Per column: (1) one CTA recodes the scalars into signed c-bit digits and
//    counting-sorts the (point, sign) references by bucket in shared memory;
//    (2) the sorted list is cut into fixed slices of SEG references; one thread sums a
//    slice with mixed XYZZ additions and emits a partial sum at every bucket boundary it
//    crosses -- witness columns are full of repeated small values, so bucket sizes are
//    heavily skewed and per-bucket threads would serialise or leave lanes idle;
//    (3) one CTA per column folds the partial sums with
//    the running-sum trick, each thread owning a contiguous bucket range, then a
//    shared-memory tree reduction and a single inversion give the affine point.
//  * The hot loop (2) gathers 64-byte affine points from the L2-resident table;
//    scalars are read once, coalesced.
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace zkfhe {

// point references summed by one accumulate thread: long slices for big batches (fewer partial sums),
// short ones when only a few columns are committed (shorter dependent chain, more threads)
// (small-valued columns have a few references per scalar: short slices keep enough threads in flight)
static inline uint32_t pick_seg(uint32_t batch, bool narrow) { return batch < 32 || narrow ? 16 : 64; }

__device__ __noinline__ void xyzz_add_ni2(g1_xyzz& acc, const g1_xyzz& p) { xyzz_add(acc, p); }

// ---- fixed-base table ---------------------------------------------------------------------
__global__ void k_msm_precompute(const g1_affine* bases, g1_affine* table, uint32_t n, uint32_t c, uint32_t W) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine p = affine_load(bases + i);
    affine_store(table + i, p);
    g1_xyzz q = xyzz_from_affine(p);
#pragma unroll 1
    for (uint32_t w = 1; w < W; w++) {
#pragma unroll 1
        for (uint32_t d = 0; d < c; d++) q = xyzz_dbl(q);
        g1_affine a = xyzz_to_affine(q);
        affine_store(table + (size_t)w * n + i, a);
        q = xyzz_from_affine(a);     // keep Z = 1 so the next doublings stay cheap
    }
}

// Prefix sums of one window row of the table, prefix[w][i] = sum_{j<i} table[w][j], i <= n (affine; prefix[w][0] is the
// identity).  Three launches: chunk totals, a serial scan of the totals per window, the prefixes inside each chunk.
static constexpr uint32_t PFX_CHUNK = 32;
__global__ void k_msm_prefix_totals(const g1_affine* table, uint32_t n, uint32_t W, g1_xyzz* totals) {
    const uint32_t chunks = (n + PFX_CHUNK - 1) / PFX_CHUNK;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * W) return;
    const uint32_t w = t / chunks, ch = t % chunks;
    const g1_affine* row = table + (size_t)w * n;
    g1_xyzz acc = xyzz_identity();
#pragma unroll 1
    for (uint32_t i = ch * PFX_CHUNK; i < min((ch + 1) * PFX_CHUNK, n); i++) xyzz_madd(acc, affine_load(row + i), false);
    xyzz_store(totals + t, acc);
}
__global__ void k_msm_prefix_scan(g1_xyzz* totals, uint32_t chunks, uint32_t W) {     // exclusive scan, one thread per window
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    g1_xyzz run = xyzz_identity();
#pragma unroll 1
    for (uint32_t ch = 0; ch < chunks; ch++) {
        g1_xyzz t = xyzz_load(totals + (size_t)w * chunks + ch);
        xyzz_store(totals + (size_t)w * chunks + ch, run);
        xyzz_add_ni2(run, t);
    }
}
__global__ void k_msm_prefix_fill(const g1_affine* table, uint32_t n, uint32_t W, const g1_xyzz* totals, g1_affine* prefix) {
    const uint32_t chunks = (n + PFX_CHUNK - 1) / PFX_CHUNK;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * W) return;
    const uint32_t w = t / chunks, ch = t % chunks;
    const g1_affine* row = table + (size_t)w * n;
    g1_affine* out = prefix + (size_t)w * (n + 1);
    g1_xyzz acc = xyzz_load(totals + t);
#pragma unroll 1
    for (uint32_t i = ch * PFX_CHUNK; i < min((ch + 1) * PFX_CHUNK, n); i++) {
        affine_store(out + i, xyzz_to_affine(acc));              // prefix[i] excludes point i
        xyzz_madd(acc, affine_load(row + i), false);
    }
    if (ch + 1 == chunks) affine_store(out + n, xyzz_to_affine(acc));
}

// ---- digit recoding -----------------------------------------------------------------------
// Calls f(w, bucket_index, negative) for every non-zero signed c-bit digit of the scalar
// (W = ceil(255/c) windows, c <= 16).  The limbs stream through a 64-bit shift register with
// compile-time limb indices: indexing the limb array by a run-time window offset made ptxas
// select limbs with compare chains, ~5x the instructions of the whole sort kernel.
template <class Fn>
__device__ __forceinline__ void for_each_digit(const fr_t& s_canon, uint32_t c, uint32_t W, Fn f) {
    const uint32_t mask = (1u << c) - 1, halfw = 1u << (c - 1);
    uint64_t buf = 0;
    uint32_t nbits = 0, w = 0, carry = 0;
    auto emit = [&](uint32_t raw) {
        uint32_t d = raw + carry;
        if (d > halfw) {
            carry = 1;
            uint32_t mag = (1u << c) - d;           // digit = d - 2^c < 0
            if (mag) f(w, mag - 1, true);
        } else {
            carry = 0;
            if (d) f(w, d - 1, false);
        }
        w++;
    };
#pragma unroll
    for (int l = 0; l < 8; l++) {
        buf |= (uint64_t)s_canon.v[l] << nbits;
        nbits += 32;
        while (nbits >= c && w < W) {
            emit((uint32_t)buf & mask);
            buf >>= c;
            nbits -= c;
        }
    }
    if (w < W) emit((uint32_t)buf & mask);          // the last, short window
}

// The references of scalar i of a column: f(table index, bucket, negative) for every non-zero digit.  With a prefix
// table behind the window table (ps_base != 0) a run of equal scalars over rows [a, b] is emitted at its two ends only:
// -prefix[w][a] at row a and +prefix[w][b + 1] at row b, since sum_{a<=i<=b} 2^(cw) P_i = prefix[w][b+1] - prefix[w][a];
// the rows in between emit nothing.  Both passes of the counting sort call this, so they agree by construction.
template <class Fn>
__device__ __forceinline__ void for_each_ref(const fr_t* sc, uint32_t i, uint32_t n, uint32_t c, uint32_t W, uint32_t ps_base, Fn f) {
    const fr_t raw = fe_load(sc + i);
    if (is_zero(raw)) return;
    uint32_t kind = 0;                                    // 0 alone, 1 first row of a run, 2 last row, 3 interior
    if (ps_base) {
        if (i > 0 && eq(raw, fe_load(sc + i - 1))) kind |= 2;
        if (i + 1 < n && eq(raw, fe_load(sc + i + 1))) kind |= 1;
        if (kind == 3) return;
    }
    const fr_t s = from_mont(raw);
    for_each_digit(s, c, W, [&](uint32_t w, uint32_t b, bool negative) {
        if (kind == 0) f(w * n + i, b, negative);
        else if (kind == 1) f(ps_base + w * (n + 1) + i, b, !negative);
        else f(ps_base + w * (n + 1) + i + 1, b, negative);
    });
}

// One CTA per column: histogram -> exclusive scans -> scatter (counting sort by bucket).
//   bucket_off[col][NB+1] : start of each bucket in sorted[col]; bucket_off[NB] = number of references
//   rank[col][NB+1]       : number of non-empty buckets before bucket b
//   sorted[col][..]       : (w*n + i) | sign<<31, grouped by bucket
extern __shared__ uint32_t msm_smem[];

__global__ void __launch_bounds__(1024) k_msm_sort(const fr_t* scalars, uint64_t stride, uint32_t n, uint32_t c,
                                                   uint32_t W, uint32_t* bucket_off, uint32_t* rank_out,
                                                   uint32_t* sorted, uint64_t sorted_stride, uint32_t skew_limit,
                                                   uint32_t* skew_out, uint32_t ps_base) {
    const uint32_t NB = 1u << (c - 1);
    uint32_t* cnt = msm_smem;                 // [NB] counts, then running cursors
    __shared__ uint32_t warp_tot[2][32];
    const uint32_t col = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const fr_t* sc = scalars + (uint64_t)col * stride;
    uint32_t* boff = bucket_off + (size_t)col * (NB + 1);
    uint32_t* rnk = rank_out + (size_t)col * (NB + 1);
    uint32_t* out = sorted + (uint64_t)col * sorted_stride;

    for (uint32_t b = tid; b < NB; b += nt) cnt[b] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += nt)
        for_each_ref(sc, i, n, c, W, ps_base, [&](uint32_t, uint32_t b, bool) { atomicAdd(&cnt[b], 1u); });
    __syncthreads();

    // exclusive scans of the counts (entries) and of the non-empty flags (ranks)
    const uint32_t per = (NB + nt - 1) / nt;
    const uint32_t b0 = tid * per;
    uint32_t sum_e = 0, sum_s = 0, max_c = 0;
    for (uint32_t k = 0; k < per; k++) {
        uint32_t b = b0 + k;
        if (b < NB) { uint32_t v = cnt[b]; sum_e += v; sum_s += (v != 0); max_c = max(max_c, v); }
    }
    // a column is "skewed" when some bucket is long enough to span many accumulate slices (witness columns
    // repeating one value thousands of times); only those columns go through the level-0 combine kernel
    if (__syncthreads_or(max_c > skew_limit) && tid == 0) skew_out[col] = 1;
    else if (tid == 0) skew_out[col] = 0;
    uint32_t lane = tid & 31, wid = tid >> 5;
    uint32_t inc_e = sum_e, inc_s = sum_s;
    for (uint32_t d = 1; d < 32; d <<= 1) {
        uint32_t te = __shfl_up_sync(0xffffffffu, inc_e, d);
        uint32_t ts = __shfl_up_sync(0xffffffffu, inc_s, d);
        if (lane >= d) { inc_e += te; inc_s += ts; }
    }
    if (lane == 31) { warp_tot[0][wid] = inc_e; warp_tot[1][wid] = inc_s; }
    __syncthreads();
    if (wid == 0) {
        uint32_t nw = (nt + 31) / 32;
        uint32_t ve = lane < nw ? warp_tot[0][lane] : 0, vs = lane < nw ? warp_tot[1][lane] : 0;
        uint32_t ie = ve, is = vs;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            uint32_t te = __shfl_up_sync(0xffffffffu, ie, d);
            uint32_t ts = __shfl_up_sync(0xffffffffu, is, d);
            if (lane >= d) { ie += te; is += ts; }
        }
        warp_tot[0][lane] = ie - ve;   // exclusive warp offsets
        warp_tot[1][lane] = is - vs;
    }
    __syncthreads();
    uint32_t run_e = warp_tot[0][wid] + inc_e - sum_e;
    uint32_t run_s = warp_tot[1][wid] + inc_s - sum_s;
    for (uint32_t k = 0; k < per; k++) {
        uint32_t b = b0 + k;
        if (b < NB) {
            uint32_t v = cnt[b];
            boff[b] = run_e;
            rnk[b] = run_s;
            cnt[b] = run_e;            // becomes the scatter cursor
            run_e += v;
            run_s += (v != 0);
        }
    }
    if (tid == nt - 1) { boff[NB] = run_e; rnk[NB] = run_s; }
    __syncthreads();

    for (uint32_t i = tid; i < n; i += nt)
        for_each_ref(sc, i, n, c, W, ps_base, [&](uint32_t ref, uint32_t b, bool negative) {
            uint32_t pos = atomicAdd(&cnt[b], 1u);
            out[pos] = ref | (negative ? 0x80000000u : 0u);
        });
}

// The same counting sort spread over a thread-block CLUSTER of SORT_CS CTAs per column (distributed shared memory):
// CTA q histograms scalars [q n/CS, (q+1) n/CS) into its own shared memory; after a cluster barrier it owns the
// bucket slice [q NB/CS, (q+1) NB/CS), reads the CS histograms of those buckets through DSMEM, scans them, and writes
// back -- into every CTA's shared memory -- that CTA's first output position inside each bucket; a second barrier,
// and every CTA scatters its own scalars with shared-memory cursors.  One CTA per column is the right shape when a
// commit phase has 100+ columns of 2^13; a 1-3 column commit, or columns of 2^16..2^19 scalars, left 140 SMs idle for
// the whole sort (74 of 480 ms at k = 19).
static constexpr uint32_t SORT_CS = 8;
__global__ void __cluster_dims__(SORT_CS, 1, 1) __launch_bounds__(1024)
k_msm_sort_cluster(const fr_t* scalars, uint64_t stride, uint32_t n, uint32_t c, uint32_t W, uint32_t* bucket_off,
                   uint32_t* rank_out, uint32_t* sorted, uint64_t sorted_stride, uint32_t skew_limit, uint32_t* skew_out,
                   uint32_t ps_base) {
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t NB = 1u << (c - 1);
    uint32_t* cnt = msm_smem;                 // [NB] this CTA's histogram, then its scatter cursors
    __shared__ uint32_t warp_tot[2][32];
    __shared__ uint32_t blk_tot[3];
    __shared__ uint32_t slice_sum[SORT_CS][3];                // (references, non-empty buckets, skewed) per bucket slice
    const uint32_t q = cluster.block_rank(), col = blockIdx.x / SORT_CS, tid = threadIdx.x, nt = blockDim.x;
    const fr_t* sc = scalars + (uint64_t)col * stride;
    uint32_t* boff = bucket_off + (size_t)col * (NB + 1);
    uint32_t* rnk = rank_out + (size_t)col * (NB + 1);
    uint32_t* out = sorted + (uint64_t)col * sorted_stride;
    const uint32_t i0 = (uint32_t)((uint64_t)n * q / SORT_CS), i1 = (uint32_t)((uint64_t)n * (q + 1) / SORT_CS);

    for (uint32_t b = tid; b < NB; b += nt) cnt[b] = 0;
    __syncthreads();
    for (uint32_t i = i0 + tid; i < i1; i += nt)
        for_each_ref(sc, i, n, c, W, ps_base, [&](uint32_t, uint32_t b, bool) { atomicAdd(&cnt[b], 1u); });
    cluster.sync();

    // bucket slice of this CTA: totals over the CS histograms, exclusive scans inside the slice
    uint32_t* rc[SORT_CS];
#pragma unroll
    for (uint32_t r = 0; r < SORT_CS; r++) rc[r] = cluster.map_shared_rank(cnt, r);
    const uint32_t slice = NB / SORT_CS, b_lo = q * slice, b_hi = b_lo + slice;
    const uint32_t per = (slice + nt - 1) / nt;
    const uint32_t b0 = b_lo + tid * per;
    uint32_t sum_e = 0, sum_s = 0, max_c = 0;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t b = b0 + k;
        if (b < b_hi) {
            uint32_t v = 0;
#pragma unroll
            for (uint32_t r = 0; r < SORT_CS; r++) v += rc[r][b];
            sum_e += v; sum_s += (v != 0); max_c = max(max_c, v);
        }
    }
    const uint32_t any_skew = __syncthreads_or(max_c > skew_limit);
    const uint32_t lane = tid & 31, wid = tid >> 5;
    uint32_t inc_e = sum_e, inc_s = sum_s;
    for (uint32_t d = 1; d < 32; d <<= 1) {
        uint32_t te = __shfl_up_sync(0xffffffffu, inc_e, d);
        uint32_t ts = __shfl_up_sync(0xffffffffu, inc_s, d);
        if (lane >= d) { inc_e += te; inc_s += ts; }
    }
    if (lane == 31) { warp_tot[0][wid] = inc_e; warp_tot[1][wid] = inc_s; }
    __syncthreads();
    if (wid == 0) {
        const uint32_t nw = (nt + 31) / 32;
        uint32_t ve = lane < nw ? warp_tot[0][lane] : 0, vs = lane < nw ? warp_tot[1][lane] : 0;
        uint32_t ie = ve, is = vs;
        for (uint32_t d = 1; d < 32; d <<= 1) {
            uint32_t te = __shfl_up_sync(0xffffffffu, ie, d);
            uint32_t ts = __shfl_up_sync(0xffffffffu, is, d);
            if (lane >= d) { ie += te; is += ts; }
        }
        warp_tot[0][lane] = ie - ve;   // exclusive warp offsets
        warp_tot[1][lane] = is - vs;
        if (lane == 31) { blk_tot[0] = ie; blk_tot[1] = is; blk_tot[2] = any_skew; }
    }
    __syncthreads();
    if (tid < SORT_CS) {                                        // publish this slice's totals to every CTA of the cluster
        uint32_t* ss = cluster.map_shared_rank(&slice_sum[0][0], tid);
        ss[q * 3 + 0] = blk_tot[0];
        ss[q * 3 + 1] = blk_tot[1];
        ss[q * 3 + 2] = blk_tot[2];
    }
    cluster.sync();
    uint32_t base_e = 0, base_s = 0, skewed = 0, all_e = 0, all_s = 0;
    for (uint32_t r = 0; r < SORT_CS; r++) {
        if (r < q) { base_e += slice_sum[r][0]; base_s += slice_sum[r][1]; }
        all_e += slice_sum[r][0]; all_s += slice_sum[r][1]; skewed |= slice_sum[r][2];
    }
    uint32_t run_e = base_e + warp_tot[0][wid] + inc_e - sum_e;
    uint32_t run_s = base_s + warp_tot[1][wid] + inc_s - sum_s;
    for (uint32_t k = 0; k < per; k++) {
        const uint32_t b = b0 + k;
        if (b < b_hi) {
            boff[b] = run_e;
            rnk[b] = run_s;
            uint32_t pos = run_e;
#pragma unroll
            for (uint32_t r = 0; r < SORT_CS; r++) {            // CTA r's first position inside bucket b
                const uint32_t h = rc[r][b];
                rc[r][b] = pos;
                pos += h;
            }
            run_s += (pos != run_e);
            run_e = pos;
        }
    }
    if (q == 0 && tid == 0) { boff[NB] = all_e; rnk[NB] = all_s; skew_out[col] = skewed ? 1u : 0u; }
    cluster.sync();                                             // every cursor is in place (and no DSMEM access after this)

    for (uint32_t i = i0 + tid; i < i1; i += nt)
        for_each_ref(sc, i, n, c, W, ps_base, [&](uint32_t ref, uint32_t b, bool negative) {
            uint32_t pos = atomicAdd(&cnt[b], 1u);
            out[pos] = ref | (negative ? 0x80000000u : 0u);
        });
}

// Thread t of a column sums references [t*SEG, (t+1)*SEG) of the bucket-sorted list with mixed XYZZ
// additions -- every lane of a warp does the same number of additions whatever the bucket sizes --
// and writes one partial sum per bucket it touches to slot rank[b] + t.  Slots are strictly
// increasing along the list, so they never collide; bucket b's partials are exactly the slots
// rank[b] + t for t in [boff[b]/SEG, (boff[b+1]-1)/SEG].
__global__ void __launch_bounds__(128) k_msm_accumulate(const g1_affine* __restrict__ table, uint32_t c, uint32_t SEG,
                                                        const uint32_t* __restrict__ bucket_off,
                                                        const uint32_t* __restrict__ rank_in,
                                                        const uint32_t* __restrict__ sorted, uint64_t sorted_stride,
                                                        g1_xyzz* partial, uint64_t partial_stride) {
    const uint32_t NB = 1u << (c - 1);
    const uint32_t col = blockIdx.y;
    const uint32_t* boff = bucket_off + (size_t)col * (NB + 1);
    const uint32_t* rnk = rank_in + (size_t)col * (NB + 1);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = boff[NB];
    const uint32_t k0 = t * SEG;
    if (k0 >= total) return;
    const uint32_t k1 = min(k0 + SEG, total);
    uint32_t lo = 0, hi = NB;               // largest b with boff[b] <= k0  (then boff[b+1] > k0)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (boff[mid] <= k0) lo = mid; else hi = mid;
    }
    uint32_t b = lo, next = boff[b + 1];
    const uint32_t* ref = sorted + (uint64_t)col * sorted_stride;
    g1_xyzz* out = partial + (uint64_t)col * partial_stride + t;
    g1_xyzz acc = xyzz_identity();
    for (uint32_t k = k0; k < k1; k++) {
        if (k == next) {
            xyzz_store(out + rnk[b], acc);
            acc = xyzz_identity();
            do { b++; next = boff[b + 1]; } while (next == k);
        }
        uint32_t e = ref[k];
        g1_affine pt = affine_load(table + (e & 0x7fffffffu));
        xyzz_madd(acc, pt, (e >> 31) != 0);
    }
    xyzz_store(out + rnk[b], acc);
}

// Out-of-line point addition for the reduction kernels: they are latency-bound and call it from
// several places, so keeping one copy keeps them inside the instruction cache.
__device__ __noinline__ void xyzz_add_ni(g1_xyzz& acc, const g1_xyzz& p) { xyzz_add(acc, p); }

// Reduction, level 0: bound the partial-sum list of every bucket.  Witness columns repeat a few
// small values thousands of times, so one bucket can own up to n/SEG consecutive partial sums; a
// serial walk over them would put a hundred dependent point additions on one thread.  Slices are
// grouped SUP at a time: thread s of a column owns level-2 slot s = rank[b] + u (u = super-slice
// index) and adds the <= SUP level-1 partials of bucket b inside super-slice u.  Same slot
// arithmetic as level 1 with SEG*SUP in place of SEG, so k_msm_fold reads the result unchanged.
static constexpr uint32_t SUP = 16;
__global__ void __launch_bounds__(128) k_msm_combine(uint32_t c, uint32_t SEG, const uint32_t* __restrict__ bucket_off,
                                                     const uint32_t* __restrict__ rank_in,
                                                     const g1_xyzz* __restrict__ partial, uint64_t partial_stride,
                                                     g1_xyzz* partial2, uint64_t partial2_stride,
                                                     const uint32_t* __restrict__ skew) {
    const uint32_t NB = 1u << (c - 1);
    const uint32_t col = blockIdx.y, s2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (!skew[col]) return;                               // short lists everywhere: k_msm_fold reads level 1 directly
    const uint32_t* boff = bucket_off + (size_t)col * (NB + 1);
    const uint32_t* rnk = rank_in + (size_t)col * (NB + 1);
    const uint32_t total = boff[NB];
    if (total == 0) return;
    const uint32_t SS = SEG * SUP;
    if (s2 > rnk[NB] + (total - 1) / SS) return;         // beyond the last used slot
    // largest b with key(b) = rank[b] + boff[b]/SS <= s2  (key is non-decreasing, key(0) = 0)
    uint32_t lo = 0, hi = NB;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (rnk[mid] + boff[mid] / SS <= s2) lo = mid; else hi = mid;
    }
    const uint32_t b = lo, e0 = boff[b], e1 = boff[b + 1];
    if (e1 == e0) return;
    const uint32_t u = s2 - rnk[b];
    if (u < e0 / SS || u > (e1 - 1) / SS) return;        // a gap slot: nothing maps here
    const uint32_t t0 = max(e0 / SEG, u * SUP), t1 = min((e1 - 1) / SEG, u * SUP + SUP - 1);
    const g1_xyzz* part = partial + (uint64_t)col * partial_stride + rnk[b];
    g1_xyzz acc = xyzz_load(part + t0);
    for (uint32_t t = t0 + 1; t <= t1; t++) xyzz_add_ni(acc, xyzz_load(part + t));
    xyzz_store(partial2 + (uint64_t)col * partial2_stride + s2, acc);
}

// Reduction, level 1: one thread per group of FOLD consecutive buckets.  Folds the partial sums of
// each bucket and runs the running-sum trick inside the group:
//   S_g = sum_b B_b,   A_g = sum_b (b - lo + 1) B_b      (so sum_b (b+1) B_b = A_g + lo * S_g)
__global__ void __launch_bounds__(128) k_msm_fold(uint32_t c, uint32_t fold, uint32_t SEG1, uint32_t SEG2, const uint32_t* __restrict__ bucket_off,
                                                  const uint32_t* __restrict__ rank_in,
                                                  const g1_xyzz* __restrict__ partial1, uint64_t partial1_stride,
                                                  const g1_xyzz* __restrict__ partial2, uint64_t partial2_stride,
                                                  const uint32_t* __restrict__ skew,
                                                  g1_xyzz* group_out /* [col][groups][2] */) {
    const uint32_t NB = 1u << (c - 1);
    const uint32_t groups = NB / fold;
    const uint32_t col = blockIdx.y, g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const uint32_t* boff = bucket_off + (size_t)col * (NB + 1);
    const uint32_t* rnk = rank_in + (size_t)col * (NB + 1);
    // skewed columns were combined SUP slices at a time, one or more times (SEG2 references per slot); the others
    // are read at level 1 (SEG1 references per slot)
    const bool sk = skew[col] != 0;
    const uint32_t SEG = sk ? SEG2 : SEG1;
    const g1_xyzz* part = sk ? partial2 + (uint64_t)col * partial2_stride : partial1 + (uint64_t)col * partial1_stride;
    const uint32_t lo = g * fold;
    g1_xyzz running = xyzz_identity(), acc = xyzz_identity();
    for (uint32_t b = lo + fold; b-- > lo;) {
        uint32_t e0 = boff[b], e1 = boff[b + 1];
        if (e1 > e0) {
            uint32_t r = rnk[b];
            for (uint32_t t = e0 / SEG; t <= (e1 - 1) / SEG; t++) xyzz_add_ni(running, xyzz_load(part + r + t));
        }
        xyzz_add_ni(acc, running);
    }
    g1_xyzz* o = group_out + ((size_t)col * groups + g) * 2;
    xyzz_store(o, running);
    xyzz_store(o + 1, acc);
}

// Reduction, level 1, latency variant for commits of a few columns (the h pieces, the two SHPLONK quotients, the
// phase-0 advice): one WARP per group of 32 buckets, lane = bucket.  The running sums become a shuffle suffix scan
// (R_l = sum_{m >= l} B_m, so S_g = R_0 and A_g = sum_l R_l by a shuffle tree): ~12 dependent point additions
// instead of 2 * 16 + the partial lists, at 5x the arithmetic -- irrelevant when the GPU is otherwise empty, wrong
// for the 100+ column commits, which stay on k_msm_fold.
__device__ __forceinline__ g1_xyzz xyzz_shfl_down_w(const g1_xyzz& p, uint32_t d) {
    g1_xyzz r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], d);
        r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], d);
        r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], d);
        r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], d);
    }
    return r;
}
__global__ void __launch_bounds__(128) k_msm_fold_warp(uint32_t c, uint32_t SEG1, uint32_t SEG2, const uint32_t* __restrict__ bucket_off,
                                                       const uint32_t* __restrict__ rank_in,
                                                       const g1_xyzz* __restrict__ partial1, uint64_t partial1_stride,
                                                       const g1_xyzz* __restrict__ partial2, uint64_t partial2_stride,
                                                       const uint32_t* __restrict__ skew,
                                                       g1_xyzz* group_out /* [col][NB/32][2] */) {
    const uint32_t NB = 1u << (c - 1), groups = NB >> 5;
    const uint32_t col = blockIdx.y, lane = threadIdx.x & 31, g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= groups) return;                              // whole warps only: blockDim is a multiple of 32
    const uint32_t* boff = bucket_off + (size_t)col * (NB + 1);
    const uint32_t* rnk = rank_in + (size_t)col * (NB + 1);
    const bool sk = skew[col] != 0;
    const uint32_t SEG = sk ? SEG2 : SEG1;
    const g1_xyzz* part = sk ? partial2 + (uint64_t)col * partial2_stride : partial1 + (uint64_t)col * partial1_stride;
    const uint32_t b = (g << 5) + lane;
    g1_xyzz R = xyzz_identity();
    const uint32_t e0 = boff[b], e1 = boff[b + 1];
    if (e1 > e0) {
        const uint32_t r = rnk[b];
        for (uint32_t t = e0 / SEG; t <= (e1 - 1) / SEG; t++) xyzz_add_ni(R, xyzz_load(part + r + t));
    }
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {               // R_l = sum_{m >= l} B_m
        g1_xyzz v = xyzz_shfl_down_w(R, d);
        if (lane + d < 32) xyzz_add_ni(R, v);
    }
    g1_xyzz A = R;
#pragma unroll 1
    for (uint32_t d = 16; d > 0; d >>= 1) {               // A_g = sum_l R_l
        g1_xyzz v = xyzz_shfl_down_w(A, d);
        if (lane < d) xyzz_add_ni(A, v);
    }
    if (lane == 0) {
        g1_xyzz* o = group_out + ((size_t)col * groups + g) * 2;
        xyzz_store(o, R);
        xyzz_store(o + 1, A);
    }
}

// Reduction, level 2: ONE WARP per column, lane l owning `per` consecutive groups.
//   total = sum_g A_g + fold * sum_g g * S_g
// Inside a lane (g0 = l * per): S_l = sum S_g, W_l = sum (g - g0) S_g (running sums), A_l = sum A_g.
// Across lanes: U_l = sum_{m >= l} S_m by a shuffle suffix scan, so sum_l l * S_l = sum_{l >= 1} U_l;
// every lane then forms V_l = A_l + fold * (W_l + per * [l >= 1] U_l) with doublings (fold and per
// are powers of two) and one shuffle tree adds the 32 V_l.  No shared memory, no barriers, ~1/3 of
// the point additions of a block-wide Hillis-Steele scan, and a column costs one warp instead of a
// whole SM; the single inversion per column is binary Euclid (inv_bin.cuh), off the IMAD pipe.
__device__ __forceinline__ g1_xyzz xyzz_shfl_down(const g1_xyzz& p, uint32_t d) {
    g1_xyzz r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], d);
        r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], d);
        r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], d);
        r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], d);
    }
    return r;
}
__device__ __noinline__ g1_xyzz xyzz_dbl_ni(const g1_xyzz p) { return xyzz_dbl(p); }

__global__ void __launch_bounds__(32) k_msm_final(uint32_t groups, uint32_t log_fold, const g1_xyzz* __restrict__ group_in,
                                                  g1_affine* out, const uint32_t* __restrict__ bucket_off, uint32_t NB,
                                                  unsigned long long* refs_total) {
    const uint32_t col = blockIdx.x, lane = threadIdx.x;
    if (lane == 0) atomicAdd(refs_total, (unsigned long long)bucket_off[(size_t)col * (NB + 1) + NB]);
    uint32_t log_per = 0;
    while ((32u << log_per) < groups) log_per++;                  // groups is a power of two
    const uint32_t per = 1u << log_per, g0 = lane << log_per;
    const g1_xyzz* in = group_in + (size_t)col * groups * 2;
    g1_xyzz S = xyzz_identity(), Wt = xyzz_identity(), A = xyzz_identity();
    if (g0 < groups) {
#pragma unroll 1
        for (uint32_t g = g0 + per; g-- > g0;) {
            xyzz_add_ni(S, xyzz_load(in + 2 * (size_t)g));
            if (g > g0) xyzz_add_ni(Wt, S);
            xyzz_add_ni(A, xyzz_load(in + 2 * (size_t)g + 1));
        }
    }
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {                       // inclusive suffix scan of S over the lanes
        g1_xyzz v = xyzz_shfl_down(S, d);
        if (lane + d < 32) xyzz_add_ni(S, v);
    }
    g1_xyzz V = lane ? S : xyzz_identity();
#pragma unroll 1
    for (uint32_t i = 0; i < log_per; i++) V = xyzz_dbl_ni(V);
    xyzz_add_ni(V, Wt);
#pragma unroll 1
    for (uint32_t i = 0; i < log_fold; i++) V = xyzz_dbl_ni(V);
    xyzz_add_ni(V, A);
#pragma unroll 1
    for (uint32_t d = 16; d > 0; d >>= 1) {
        g1_xyzz v = xyzz_shfl_down(V, d);
        if (lane < d) xyzz_add_ni(V, v);
    }
    if (lane == 0) affine_store(out + col, xyzz_to_affine(V));
}

// Reduction, level 2, CTA-wide: one THREAD per group instead of one lane per groups/32 groups.  The in-lane running
// sums of k_msm_final (3 * groups/32 dependent point additions before the warp scan even starts) become part of the
// scan: suffix sums U_g = sum_{m >= g} S_m by a shuffle scan inside each warp plus one shuffle scan of the warp totals,
// V_g = A_g + fold * [g >= 1] U_g, one shuffle tree per warp and one over the warp results.  ~20 dependent point
// operations + the inversion instead of ~35, on the same data layout; it is what every commit of a proof ends with
// (seven times ~250 us at 1.6 % warps active, profiles/r02_ncu_full_bench_reduce.txt).
__global__ void __launch_bounds__(512) k_msm_final_cta(uint32_t groups, uint32_t log_fold, const g1_xyzz* __restrict__ group_in,
                                                       g1_affine* out, const uint32_t* __restrict__ bucket_off, uint32_t NB,
                                                       unsigned long long* refs_total) {
    __shared__ g1_xyzz sh[16];                                    // one slot per warp (<= 512 threads)
    const uint32_t col = blockIdx.x, g = threadIdx.x, lane = g & 31, wid = g >> 5, nw = blockDim.x >> 5;
    if (g == 0) atomicAdd(refs_total, (unsigned long long)bucket_off[(size_t)col * (NB + 1) + NB]);
    const g1_xyzz* in = group_in + (size_t)col * groups * 2;
    g1_xyzz S = xyzz_load(in + 2 * (size_t)g);
    const g1_xyzz A = xyzz_load(in + 2 * (size_t)g + 1);
#pragma unroll 1
    for (uint32_t d = 1; d < 32; d <<= 1) {                       // inclusive suffix scan of S inside the warp
        g1_xyzz v = xyzz_shfl_down(S, d);
        if (lane + d < 32) xyzz_add_ni(S, v);
    }
    if (lane == 0) sh[wid] = S;                                   // this warp's total
    __syncthreads();
    if (wid == 0) {                                               // exclusive suffix sums of the warp totals
        g1_xyzz T = lane < nw ? sh[lane] : xyzz_identity();
#pragma unroll 1
        for (uint32_t d = 1; d < nw; d <<= 1) {
            g1_xyzz v = xyzz_shfl_down(T, d);
            if (lane + d < nw) xyzz_add_ni(T, v);
        }
        // T = sum_{m >= lane}; what warp `lane` must add is sum_{m > lane} = the next lane's inclusive sum
        g1_xyzz nxt = xyzz_shfl_down(T, 1);
        if (lane < nw) sh[lane] = lane + 1 < nw ? nxt : xyzz_identity();
    }
    __syncthreads();
    if (wid + 1 < nw) xyzz_add_ni(S, sh[wid]);                    // S = U_g
    g1_xyzz V = g ? S : xyzz_identity();
#pragma unroll 1
    for (uint32_t i = 0; i < log_fold; i++) V = xyzz_dbl_ni(V);
    xyzz_add_ni(V, A);
#pragma unroll 1
    for (uint32_t d = 16; d > 0; d >>= 1) {
        g1_xyzz v = xyzz_shfl_down(V, d);
        if (lane < d) xyzz_add_ni(V, v);
    }
    __syncthreads();                                              // everyone has read its suffix from sh
    if (lane == 0) sh[wid] = V;
    __syncthreads();
    if (wid == 0) {
        g1_xyzz W = lane < nw ? sh[lane] : xyzz_identity();
#pragma unroll 1
        for (uint32_t d = 16; d > 0; d >>= 1) {
            g1_xyzz v = xyzz_shfl_down(W, d);
            if (lane < d) xyzz_add_ni(W, v);
        }
        if (lane == 0) affine_store(out + col, xyzz_to_affine(W));
    }
}

// ---- test SRS (halo2 `ParamsKZG::setup` shape): g[i] = tau^i G, g_lagrange[i] = l_i(tau) G --------
__device__ __noinline__ g1_affine g1_generator_mul(const fr_t k_canon) {
    g1_affine g;
    g.x = fe_one<FQ>();
    g.y = add(fe_one<FQ>(), fe_one<FQ>());      // G = (1, 2)
    g1_xyzz acc = xyzz_identity();
    bool started = false;
#pragma unroll 1
    for (int i = 7; i >= 0; i--)
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            if (started) acc = xyzz_dbl(acc);
            if ((k_canon.v[i] >> bit) & 1) { xyzz_madd(acc, g, false); started = true; }
        }
    return xyzz_to_affine(acc);
}
__global__ void k_srs_setup(fr_t tau /*Montgomery*/, uint32_t log_n, g1_affine* g, g1_affine* gl) {
    const uint32_t n = 1u << log_n;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (g) affine_store(g + i, g1_generator_mul(from_mont(pow_u64(tau, i))));
    if (gl) {
        // l_i(tau) = w^i (tau^n - 1) / (n (tau - w^i))
        fr_t w = fr_t{ZKFHE_FR_ROOT_OF_UNITY_MONT};
        for (uint32_t s = log_n; s < 28; s++) w = sqr(w);
        fr_t wi = pow_u64(w, i);
        fr_t tn = sub(pow_u64(tau, n), fe_one<FR>());
        fr_t nn = fe_zero<FR>();
        nn.v[0] = n;
        fr_t den = mul(to_mont(nn), sub(tau, wi));
        fr_t li = mul(mul(wi, tn), inv(den));
        affine_store(gl + i, g1_generator_mul(from_mont(li)));
    }
}
__global__ void k_fr_convert(fr_t* data, uint64_t count, int to_montgomery) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fr_t v = fe_load(data + i);
    fe_store(data + i, to_montgomery ? to_mont(v) : from_mont(v));
}
int srs_setup(zkfhe_ctx* ctx, uint32_t log_n, const fr_t& tau_mont, g1_affine* d_g, g1_affine* d_gl) {
    k_srs_setup<<<((1u << log_n) + 63) / 64, 64, 0, ctx->stream>>>(tau_mont, log_n, d_g, d_gl);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}
__global__ void k_points_to_canonical(g1_affine* pts, uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fe_store(&pts[i].x, from_mont(fe_load(&pts[i].x)));
    fe_store(&pts[i].y, from_mont(fe_load(&pts[i].y)));
}
// In place: Montgomery affine points -> canonical coordinates (proof / transcript encoding).
int points_to_canonical(zkfhe_ctx* ctx, g1_affine* d_pts, uint32_t count) {
    if (!count) return ZKFHE_OK;
    k_points_to_canonical<<<(count + 127) / 128, 128, 0, ctx->stream>>>(d_pts, count);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}
int fr_convert(zkfhe_ctx* ctx, fr_t* d, uint64_t count, int to_montgomery) {
    if (!count) return ZKFHE_OK;
    k_fr_convert<<<(uint32_t)((count + 255) / 256), 256, 0, ctx->stream>>>(d, count, to_montgomery);
    ZK_CHECK_LAUNCH(ctx);
    return ZKFHE_OK;
}

static uint32_t pick_window(uint32_t log_n) {
    uint32_t c = log_n;                 // buckets ~ n/2: balances n*W additions against bucket reduction
    if (c < 4) c = 4;
    if (c > 15) c = 15;
    return c;
}

// window table (+ prefix sums behind it) of one expansion: table[w*n + i] = 2^(c*w) P_i
static int build_table(zkfhe_ctx* ctx, const g1_affine* d_bases, size_t n, uint32_t c, uint32_t W, bool prefix, g1_affine** out) {
    const size_t entries = n * W + (prefix ? (n + 1) * W : 0);
    if (entries >= (1ull << 31)) return fail(ctx, ZKFHE_ERR_ARG, "msm: %zu table entries do not fit 31 bits", entries);
    ZK_CUDA(ctx, cudaMalloc(out, entries * sizeof(g1_affine)));
    k_msm_precompute<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(d_bases, *out, (uint32_t)n, c, W);
    ZK_CHECK_LAUNCH(ctx);
    if (prefix) {
        const uint32_t chunks = (uint32_t)((n + PFX_CHUNK - 1) / PFX_CHUNK);
        g1_xyzz* totals;
        ZK_TRY(ws_get(ctx, "msm_pfx_totals", (size_t)chunks * W * sizeof(g1_xyzz), (void**)&totals));
        k_msm_prefix_totals<<<(chunks * W + 127) / 128, 128, 0, ctx->stream>>>(*out, (uint32_t)n, W, totals);
        ZK_CHECK_LAUNCH(ctx);
        k_msm_prefix_scan<<<(W + 31) / 32, 32, 0, ctx->stream>>>(totals, chunks, W);
        ZK_CHECK_LAUNCH(ctx);
        k_msm_prefix_fill<<<(chunks * W + 127) / 128, 128, 0, ctx->stream>>>(*out, (uint32_t)n, W, totals, *out + n * W);
        ZK_CHECK_LAUNCH(ctx);
    }
    return ZKFHE_OK;
}

int msm_load_basis(zkfhe_ctx* ctx, int which, const g1_affine* d_bases, uint32_t log_n, bool prefix) {
    MsmBasis& B = ctx->basis[which];
    if (B.table && !B.shared) {
        ZK_CUDA(ctx, cudaFree(B.table));
        if (B.table_s) ZK_CUDA(ctx, cudaFree(B.table_s));
    }
    B.table = B.table_s = nullptr; B.loaded = false; B.shared = false;
    B.log_n = log_n;
    B.c = pick_window(log_n);
    B.W = (255 + B.c - 1) / B.c;
    B.prefix = prefix && log_n >= 8;
    if (const char* e = getenv("ZKFHE_MSM_PREFIX")) B.prefix = B.prefix && atoi(e) != 0;
    size_t n = (size_t)1 << log_n;
    ZK_TRY(build_table(ctx, d_bases, n, B.c, B.W, B.prefix, &B.table));
    B.c_s = B.W_s = 0;
    if (log_n >= 11) {                   // narrow-window expansion for small-valued columns
        B.c_s = B.c - 3;
        B.W_s = (255 + B.c_s - 1) / B.c_s;
        ZK_TRY(build_table(ctx, d_bases, n, B.c_s, B.W_s, B.prefix, &B.table_s));
    }
    ZK_CUDA(ctx, zkfhe::stream_wait(ctx));
    B.loaded = true;
    return ZKFHE_OK;
}

int msm_run(zkfhe_ctx* ctx, const fr_t* d_scalars, uint64_t stride, uint32_t log_n, uint32_t batch, int which,
            g1_affine* d_out, int small_values) {
    if (which < 0 || which > 1) return fail(ctx, ZKFHE_ERR_ARG, "msm: basis must be 0 or 1");
    MsmBasis& B = ctx->basis[which];
    if (!B.loaded) return fail(ctx, ZKFHE_ERR_STATE, "msm: zkfhe_load_srs has not been called");
    if (log_n != B.log_n) return fail(ctx, ZKFHE_ERR_ARG, "msm: log_n=%u but SRS has k=%u", log_n, B.log_n);
    if (batch == 0) return ZKFHE_OK;
    // (measured and not kept: sending full-size few-column commits through the narrow table as well -- 8x fewer buckets
    // on the latency-bound reduction against 30 % more point additions -- made sort + reduce 4.68 -> 5.02 ms per proof)
    const bool narrow = small_values && B.table_s;
    const g1_affine* table = narrow ? B.table_s : B.table;
    const uint32_t n = 1u << log_n, c = narrow ? B.c_s : B.c, W = narrow ? B.W_s : B.W, NB = 1u << (c - 1);
    const uint64_t max_refs = (uint64_t)n * W;
    const uint32_t ps_base = B.prefix ? n * W : 0;          // runs of equal scalars are referenced through the prefix table
    const uint32_t SEG = pick_seg(batch, narrow);
    const uint64_t max_thr = (max_refs + SEG - 1) / SEG;     // accumulate threads per column
    const uint64_t max_segs = NB + max_thr;                  // partial slots: rank[b] + t
    // bound the workspace: process the batch in chunks
    const uint64_t per_col = max_refs * 4 + max_segs * sizeof(g1_xyzz) + 2ull * (NB + 1) * 4;
    uint32_t chunk = (uint32_t)((3ull << 30) / per_col);
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    if (chunk > 65535) chunk = 65535;
    uint32_t *boff, *soff, *sorted;
    g1_xyzz* partial;
    ZK_TRY(ws_get(ctx, "msm_boff", (size_t)chunk * (NB + 1) * 4, (void**)&boff));
    ZK_TRY(ws_get(ctx, "msm_soff", (size_t)chunk * (NB + 1) * 4, (void**)&soff));
    ZK_TRY(ws_get(ctx, "msm_sorted", (size_t)chunk * max_refs * 4, (void**)&sorted));
    ZK_TRY(ws_get(ctx, "msm_partial", (size_t)chunk * max_segs * sizeof(g1_xyzz), (void**)&partial));
    // combine levels for skewed columns: level l has one slot per (bucket, run of SEG * SUP^l references); enough
    // levels that even a bucket holding one reference per scalar ends with <= 32 partial sums for the fold chain
    // (one level at k = 13; a single level left 2048-long chains on the 0/1-valued witness columns at k = 19)
    uint32_t levels = 1;
    uint64_t seg_top = (uint64_t)SEG * SUP;
    while (levels < 3 && n / seg_top > 32) { levels++; seg_top *= SUP; }
    uint64_t lvl_slots[4] = {max_segs, 0, 0, 0};
    g1_xyzz* lvl_buf[4] = {partial, nullptr, nullptr, nullptr};
    {
        static const char* names[4] = {"", "msm_partial2", "msm_partial3", "msm_partial4"};
        uint64_t div = 1;
        for (uint32_t l = 1; l <= levels; l++) {
            div *= SUP;
            lvl_slots[l] = NB + (max_thr + div - 1) / div + 1;
            ZK_TRY(ws_get(ctx, names[l], (size_t)chunk * lvl_slots[l] * sizeof(g1_xyzz), (void**)&lvl_buf[l]));
        }
    }
    uint32_t* skew;
    ZK_TRY(ws_get(ctx, "msm_skew", (size_t)chunk * 4, (void**)&skew));
    // reduction shape: groups of 2^log_fold buckets; the final warp owns groups/32 groups per lane.
    // 16 buckets per fold thread and <= 512 groups balance the two dependent chains (2*fold and
    // 3*groups/32 point additions).
    uint32_t log_fold = 4;
    while (log_fold && (NB >> log_fold) == 0) log_fold--;
    // narrow-window commits have 8x fewer buckets: shorter fold chains keep a full GPU's worth of threads
    while (log_fold > 2 && (uint64_t)batch * (NB >> log_fold) < 32768) log_fold--;
    while ((NB >> log_fold) > 512u) log_fold++;
    const bool warp_fold = batch < 32 && NB >= 1024 && (NB >> 5) <= 512u;     // few columns: the latency variant
    if (warp_fold) log_fold = 5;
    const uint32_t groups = NB >> log_fold;
    g1_xyzz* grp;
    ZK_TRY(ws_get(ctx, "msm_groups", (size_t)chunk * groups * 2 * sizeof(g1_xyzz), (void**)&grp));
    // function attributes are process-wide: always raise them to the fixed maximum any call can need,
    // never to this call's size (several contexts may be launching from different host threads)
    size_t smem = (size_t)NB * 4;
    unsigned long long* refs;       // running count of point additions (zkfhe_timing_get category 5)
    {
        const bool fresh = ctx->ws.find("msm_refs") == ctx->ws.end();
        ZK_TRY(ws_get(ctx, "msm_refs", 8, (void**)&refs));
        if (fresh) ZK_CUDA(ctx, cudaMemsetAsync(refs, 0, 8, ctx->stream));
    }
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    ZK_CUDA(ctx, cudaFuncSetAttribute(k_msm_sort_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    // a cluster of 8 CTAs per column when there are few columns or they are long; one CTA per column otherwise
    bool cluster_sort = log_n >= 13 && NB >= 8 * SORT_CS && (batch < 32 || log_n >= 15);
    if (const char* e = getenv("ZKFHE_MSM_CLUSTER_SORT")) cluster_sort = cluster_sort && atoi(e) != 0;
    // the final reduction as one thread per group (a CTA per column) wherever a column has 64..512 groups
    bool cta_final = groups >= 64 && groups <= 512 && (groups & (groups - 1)) == 0;
    if (const char* e = getenv("ZKFHE_MSM_CTA_FINAL")) cta_final = cta_final && atoi(e) != 0;
    // a column goes through the combine levels when some bucket holds more references than this: 3 slices for the big
    // batches, 8 for few-column commits (16-reference slices).  Uniform scalars sit at ~40 references per bucket except for
    // the 64 buckets the short top window feeds (~170 each at k = 13), so they are still combined; measured: skipping the
    // combine level for them (limit 16 slices) lengthens the fold lanes of those buckets by more than the ~95 us launch
    // it saves (sort + reduce 4.39 -> 4.56 ms per proof).
    const uint32_t skew_limit = (batch < 32 ? 8 : 3) * SEG;
    timed_call_start(ctx);
    for (uint32_t done = 0; done < batch; done += chunk) {
        uint32_t nb = batch - done < chunk ? batch - done : chunk;
        const fr_t* sc = d_scalars + (uint64_t)done * stride;
        ZK_TRY(timed_begin(ctx, ZK_CAT_MSM_OTHER, 0));
        if (cluster_sort)
            k_msm_sort_cluster<<<nb * SORT_CS, 1024, smem, ctx->stream>>>(sc, stride, n, c, W, boff, soff, sorted, max_refs, skew_limit, skew, ps_base);
        else
            k_msm_sort<<<nb, 1024, smem, ctx->stream>>>(sc, stride, n, c, W, boff, soff, sorted, max_refs, skew_limit, skew, ps_base);
        ZK_CHECK_LAUNCH(ctx);
        ZK_TRY(timed_end(ctx));
        ZK_TRY(timed_begin(ctx, ZK_CAT_MSM_ACCUMULATE, (uint64_t)nb * n));
        dim3 grid((uint32_t)((max_thr + 127) / 128), nb);
        k_msm_accumulate<<<grid, 128, 0, ctx->stream>>>(table, c, SEG, boff, soff, sorted, max_refs, partial, max_segs);
        ZK_CHECK_LAUNCH(ctx);
        ZK_TRY(timed_end(ctx));
        ZK_TRY(timed_begin(ctx, ZK_CAT_MSM_FOLD, 0));
        {
            uint32_t seg_in = SEG;
            for (uint32_t l = 1; l <= levels; l++) {
                dim3 cgrid((uint32_t)((lvl_slots[l] + 127) / 128), nb);
                k_msm_combine<<<cgrid, 128, 0, ctx->stream>>>(c, seg_in, boff, soff, lvl_buf[l - 1], lvl_slots[l - 1], lvl_buf[l], lvl_slots[l], skew);
                ZK_CHECK_LAUNCH(ctx);
                seg_in *= SUP;
            }
        }
        if (warp_fold) {
            dim3 fgrid((groups * 32 + 127) / 128, nb);
            k_msm_fold_warp<<<fgrid, 128, 0, ctx->stream>>>(c, SEG, (uint32_t)seg_top, boff, soff, partial, max_segs, lvl_buf[levels], lvl_slots[levels], skew, grp);
        } else {
            dim3 fgrid((groups + 127) / 128, nb);
            k_msm_fold<<<fgrid, 128, 0, ctx->stream>>>(c, 1u << log_fold, SEG, (uint32_t)seg_top, boff, soff, partial, max_segs, lvl_buf[levels], lvl_slots[levels], skew, grp);
        }
        ZK_CHECK_LAUNCH(ctx);
        ZK_TRY(timed_end(ctx));
        ZK_TRY(timed_begin(ctx, ZK_CAT_MSM_FINAL, 0));
        if (cta_final) k_msm_final_cta<<<nb, groups, 0, ctx->stream>>>(groups, log_fold, grp, d_out + done, boff, NB, refs);
        else k_msm_final<<<nb, 32, 0, ctx->stream>>>(groups, log_fold, grp, d_out + done, boff, NB, refs);
        ZK_CHECK_LAUNCH(ctx);
        ZK_TRY(timed_end(ctx));
    }
    return ZKFHE_OK;
}

// sum_i s_i * P_i over ARBITRARY points (the verifier's commitment combination): the points are expanded into a
// throw-away fixed-base table and go through the same sort / accumulate / reduce pipeline.  Host buffers in, host
// point out; Montgomery coordinates, identity = zeros.
int msm_variable_base(zkfhe_ctx* ctx, const g1_affine* h_points, const fr_t* h_scalars, uint32_t count, g1_affine* h_out) {
    uint32_t log_n = 4;
    while ((1u << log_n) < count) log_n++;
    const size_t n = (size_t)1 << log_n;
    g1_affine* d_pts;
    fr_t* d_sc;
    g1_affine* d_out;
    ZK_TRY(ws_get(ctx, "vmsm_pts", n * sizeof(g1_affine), (void**)&d_pts));
    ZK_TRY(ws_get(ctx, "vmsm_sc", n * sizeof(fr_t), (void**)&d_sc));
    ZK_TRY(ws_get(ctx, "vmsm_out", sizeof(g1_affine), (void**)&d_out));
    ZK_CUDA(ctx, cudaMemsetAsync(d_pts, 0, n * sizeof(g1_affine), ctx->stream));
    ZK_CUDA(ctx, cudaMemsetAsync(d_sc, 0, n * sizeof(fr_t), ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_pts, h_points, count * sizeof(g1_affine), cudaMemcpyHostToDevice, ctx->stream));
    ZK_CUDA(ctx, cudaMemcpyAsync(d_sc, h_scalars, count * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    const MsmBasis saved = ctx->basis[0];
    ctx->basis[0] = MsmBasis{};
    int rc = msm_load_basis(ctx, 0, d_pts, log_n, false);
    if (rc == ZKFHE_OK) rc = msm_run(ctx, d_sc, n, log_n, 1, 0, d_out, 0);
    if (rc == ZKFHE_OK && cudaMemcpyAsync(h_out, d_out, sizeof(g1_affine), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        rc = fail(ctx, ZKFHE_ERR_CUDA, "msm_variable_base: copy back failed");
    zkfhe::stream_wait(ctx);
    if (ctx->basis[0].table) cudaFree(ctx->basis[0].table);
    if (ctx->basis[0].table_s) cudaFree(ctx->basis[0].table_s);
    ctx->basis[0] = saved;
    return rc;
}

}  // namespace zkfhe
