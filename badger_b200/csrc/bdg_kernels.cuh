// bdg_kernels.cuh -- hand-written CUDA kernels (sm_100a) for the barcode hot path.
//
// Kernel inventory (DESIGN.md has the roofline of each):
//   edges_kernel<MODE>   all-pairs edge construction over a sorted distinct-barcode array      INT-ALU bound
//   nearest_kernel       Q x W bounded plain edit distance, first minimum per query           INT-ALU bound
//   kmer_score_kernel    Q x W shared-6-mer product counts with per-position multiplicities   INT-ALU bound
//   member_kernel        sorted-whitelist membership (smem pivots + L2-resident search)       HBM/L2 bound
//   pack16_kernel        16 ASCII bases -> 2-bit packed uint32 + validity                     HBM bound
//   pipe_probe_kernel    LOP3 / IMAD / POPC issue-rate probe for the integer roofline
//
// No tensor cores on purpose: nothing here is a dense contraction (BASELINE.json north_star).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bdg_core.cuh"

namespace bdg {

constexpr int NT = 256;            // threads per CTA
constexpr int RA = 8;              // a-rows held in registers per thread
constexpr int ROW_TILE = NT * RA;  // 2048 == BDG_ROW_TILE

struct EdgeOut {
    uint32_t* a;
    uint32_t* b;
    uint8_t* d;
    unsigned long long* count;
    unsigned long long cap;
};

struct EdgeWork {
    const uint32_t* sorted;     // N strictly increasing barcodes
    uint32_t N;
    int t;                      // edit-distance threshold
    int T;                      // q-gram threshold T(t)
    const uint32_t* tile_ids;   // K row tiles owned by this part
    const uint32_t* item_start; // K+1 prefix sums of column chunks per owned tile
    uint32_t K;
    uint32_t n_items;
    uint32_t chunk_cols;        // columns per work item (multiple of 256)
    unsigned int* item_counter; // dynamic tile scheduler
    uint32_t one;               // == 1, opaque to the compiler: x*one+c keeps integer adds on the FMA pipe (IMAD)
};

// ---- output: warp-aggregated append (one atomic per warp that has anything to emit) ---------------
__device__ __forceinline__ void emit_warp(bool ok, uint32_t a, uint32_t b, int d, const EdgeOut& out)
{
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(out.count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ok) {
        const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < out.cap) {
            out.a[pos] = a;
            out.b[pos] = b;
            out.d[pos] = (uint8_t)d;
        }
    }
}

__device__ __forceinline__ void emit_single(uint32_t a, uint32_t b, int d, const EdgeOut& out)
{
    const unsigned long long pos = atomicAdd(out.count, 1ull);
    if (pos < out.cap) {
        out.a[pos] = a;
        out.b[pos] = b;
        out.d[pos] = (uint8_t)d;
    }
}

// exact stage: D (case analysis for t<=2, bit-vector pass otherwise), then S only for survivors
template <int MODE>
__device__ __forceinline__ int exact_edge(uint32_t a, uint32_t b, int t, int T)
{
    if (!(a < b)) return 0;   // rows are sorted and distinct: index order == value order; pads fail here
    const int d = (MODE == 3) ? dist3_min(a, b) : dist_small(a, b);
    if (d > t) return 0;
    return qgram_score(a, b) >= T ? d : 0;
}

__device__ __forceinline__ uint32_t pick4(const uint4& v, int k)
{
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}


// ---- stage-1 inner steps in PTX.  The integer ALU pipe (LOP3/IADD3/ISETP/SEL, 64 lanes/clk/SM) is the
// bound of this kernel, the FMA pipe (IMAD) issues beside it.  Written as PTX so that (a) the subtraction of
// the zero-field test is an IMAD (x*one + (-ones), `one` is a kernel argument the compiler cannot fold) and
// (b) the hit bit is set by ONE predicated LOP3 instead of ISETP+SEL+LOP3.
//   lop3 immLut: operands (A,B,C) = 0xF0,0xCC,0xAA;  A & ~B = 0x30;  C | (A & ~B) = 0xBA.
__device__ __forceinline__ void pair_t2(uint32_t a, uint32_t b0, uint32_t bR, uint32_t bL, uint32_t one, uint32_t neg_ones,
                                        uint32_t& hits, const uint32_t bit)
{
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 x0, x1, x2, y0, y1, y2, m;\n\t"
        "xor.b32 x0, %1, %2;\n\t"
        "xor.b32 x1, %1, %3;\n\t"
        "xor.b32 x2, %1, %4;\n\t"
        "mad.lo.u32 y0, x0, %5, %6;\n\t"
        "mad.lo.u32 y1, x1, %5, %6;\n\t"
        "mad.lo.u32 y2, x2, %5, %6;\n\t"
        "lop3.b32 m, y0, x0, 0, 0x30;\n\t"
        "lop3.b32 m, y1, x1, m, 0xBA;\n\t"
        "lop3.b32 m, y2, x2, m, 0xBA;\n\t"
        "and.b32 m, m, %7;\n\t"
        "setp.ne.u32 p, m, 0;\n\t"
        "@p or.b32 %0, %0, %8;\n\t"
        "}"
        : "+r"(hits)
        : "r"(a), "r"(b0), "r"(bR), "r"(bL), "r"(one), "r"(neg_ones), "n"(F10_HIGH), "r"(bit));
}

// t=1: both test words of one pair folded into the running accumulators of this column.
//   word 0 (two 16-bit halves, no spare bit): zero-field test  x = a^b, y = x-ones, accA |= y & ~x   (ALU, FMA, ALU)
//   word 1 (two 14-bit fields, each with a guard bit above it): equality as two guarded subtractions
//          d1 = (a2|G) - b1, d2 = (b1|G) - a2; the guard survives in both iff the fields are equal;
//          accB |= d1 & d2                                                                           (FMA, FMA, ALU)
// so a pair costs 3 ALU-pipe + 3 FMA-pipe instructions: the two pipes are loaded evenly.
//   lop3 immLut: C | (A & ~B) = 0xBA;  C | (A & B) = 0xEA.
constexpr uint32_t T1_GUARD = 0x40004000u;

__device__ __forceinline__ void pair_t1_acc(uint32_t a, uint32_t a2, uint32_t aG, uint32_t b0, uint32_t b1, uint32_t bG,
                                            uint32_t one, uint32_t mone, uint32_t neg_ones, uint32_t& accA, uint32_t& accB)
{
    asm("{\n\t"
        ".reg .b32 x0, y0, d1, d2;\n\t"
        "xor.b32 x0, %2, %5;\n\t"
        "mad.lo.u32 y0, x0, %8, %10;\n\t"
        "mad.lo.u32 d1, %6, %9, %4;\n\t"
        "mad.lo.u32 d2, %3, %9, %7;\n\t"
        "lop3.b32 %0, y0, x0, %0, 0xBA;\n\t"
        "lop3.b32 %1, d1, d2, %1, 0xEA;\n\t"
        "}"
        : "+r"(accA), "+r"(accB)
        : "r"(a), "r"(a2), "r"(aG), "r"(b0), "r"(b1), "r"(bG), "r"(one), "r"(mone), "r"(neg_ones));
}

// ---------------------------------------------------------------------------------------------------
// edges_kernel: persistent CTAs pull (row tile, column chunk) work items from an atomic counter.
//   * each thread keeps RA=8 barcodes a_r of the row tile in registers,
//   * the column chunk is staged through shared memory in sub-tiles of SB columns, pre-shifted once per
//     element, so the inner loop is pure LOP3/IADD on registers fed by broadcast 128-bit shared loads,
//   * stage 1 (prefilter, bdg_core.cuh) decides > 99 % of the pairs in a handful of integer instructions
//     and leaves ONE BIT per pair in a per-thread hit mask (one 32-bit word per 4 columns x 8 rows, kept in
//     shared memory): no queues, no atomics, no overflow case, no divergence in the hot loop,
//   * stage 2: every thread streams through its own hit bits, evaluates D and S exactly (bdg_core.cuh) and
//     appends the edges it finds.  Dense neighbourhoods next to the diagonal simply have more bits set.
// MODE 1: t = 1.  MODE 2: t = 2.  MODE 3: any t, no prefilter (every bit set, exact stage on every pair).
// ---------------------------------------------------------------------------------------------------
constexpr int SB = 128;            // columns per sub-tile
constexpr int NQ = SB / 4;         // hit-mask words per thread and sub-tile

template <int MODE>
__global__ void __launch_bounds__(NT, 3) edges_kernel(const EdgeWork w, const EdgeOut out)
{
    constexpr int NW = MODE == 3 ? 1 : 3;
    __shared__ __align__(16) uint32_t s_a[ROW_TILE];
    __shared__ __align__(16) uint32_t s_b[NW][SB];
    __shared__ uint32_t s_mask[NQ * NT];
    __shared__ uint32_t s_item[3];

    const int tid = threadIdx.x;
    const uint32_t one = w.one;
    const uint32_t mone = 0u - one;                                            // runtime -1: a*mone + c is c - a on the FMA pipe
    const uint32_t neg10 = 0u - F10_ONES * one, neg16 = 0u - F16_ONES * one;   // runtime values: not foldable

    for (;;) {
        __syncthreads();   // previous item fully drained before s_item / s_a are overwritten
        if (tid == 0) {
            const uint32_t item = atomicAdd(w.item_counter, 1u);
            uint32_t k = 0, j = 0;
            if (item < w.n_items) {
                uint32_t lo = 0, hi = w.K;   // largest k with item_start[k] <= item
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (__ldg(&w.item_start[mid]) <= item) lo = mid; else hi = mid;
                }
                k = lo;
                j = item - __ldg(&w.item_start[k]);
            }
            s_item[0] = item; s_item[1] = k; s_item[2] = j;
        }
        __syncthreads();
        if (s_item[0] >= w.n_items) break;
        const uint64_t row0 = (uint64_t)__ldg(&w.tile_ids[s_item[1]]) * ROW_TILE;
        const uint64_t col_lo = row0 + (uint64_t)s_item[2] * w.chunk_cols;
        const uint64_t col_hi = min((uint64_t)w.N, col_lo + w.chunk_cols);

        uint32_t a[RA];
        uint32_t a2[RA];   // MODE 1 only: second test word of a
#pragma unroll
        for (int r = 0; r < RA; r++) {
            const uint64_t idx = row0 + (uint64_t)r * NT + tid;
            a[r] = idx < w.N ? __ldg(&w.sorted[idx]) : 0xFFFFFFFFu;   // pad: never the smaller of a pair
            s_a[r * NT + tid] = a[r];
            a2[r] = t1_word_a(a[r]);
        }

        for (uint64_t sub = col_lo; sub < col_hi; sub += SB) {
            __syncthreads();   // everyone is done with the previous sub-tile's s_b
            if (tid < SB) {
                const uint64_t idx = sub + tid;
                const uint32_t b = idx < col_hi ? __ldg(&w.sorted[idx]) : 0u;   // pad: never the larger of a pair
                s_b[0][tid] = b;
                if constexpr (MODE == 1) { s_b[1][tid] = t1_word_b(b); s_b[2][tid] = t1_word_b(b) | T1_GUARD; }
                if constexpr (MODE == 2) { s_b[1][tid] = b >> 2; s_b[2][tid] = b << 2; }
            }
            __syncthreads();

            // ---------------- stage 1: one hit bit per pair, bit (k*8 + r) of word xq ----------------
            uint32_t any = 0;
            if constexpr (MODE == 1) {
#pragma unroll 1
                for (int xq = 0; xq < NQ; xq++) {
                    const uint4 B0 = *reinterpret_cast<const uint4*>(&s_b[0][xq * 4]);
                    const uint4 B1 = *reinterpret_cast<const uint4*>(&s_b[1][xq * 4]);
                    const uint4 BG = *reinterpret_cast<const uint4*>(&s_b[2][xq * 4]);
                    uint32_t hits = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t b0 = pick4(B0, k), b1 = pick4(B1, k), bG = pick4(BG, k);
                        uint32_t accA = 0, accB = 0;
#pragma unroll
                        for (int r = 0; r < RA; r++) pair_t1_acc(a[r], a2[r], a2[r] | T1_GUARD, b0, b1, bG, one, mone, neg16, accA, accB);
                        const uint32_t acc = (accA & F16_HIGH) | (accB & T1_GUARD);
                        if (acc) {   // ~1.4e-3 per thread and column on random data
#pragma unroll
                            for (int r = 0; r < RA; r++) {
                                const uint32_t x0 = a[r] ^ b0, x1 = a2[r] ^ b1;
                                if ((((x0 - F16_ONES) & ~x0) | ((x1 - F16_ONES) & ~x1)) & F16_HIGH) hits |= 1u << (k * 8 + r);
                            }
                        }
                    }
                    s_mask[xq * NT + tid] = hits;
                    any |= hits;
                }
            } else if constexpr (MODE == 2) {
#pragma unroll 1
                for (int xq = 0; xq < NQ; xq++) {
                    const uint4 B0 = *reinterpret_cast<const uint4*>(&s_b[0][xq * 4]);
                    const uint4 BR = *reinterpret_cast<const uint4*>(&s_b[1][xq * 4]);
                    const uint4 BL = *reinterpret_cast<const uint4*>(&s_b[2][xq * 4]);
                    uint32_t hits = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t b0 = pick4(B0, k), bR = pick4(BR, k), bL = pick4(BL, k);
#pragma unroll
                        for (int r = 0; r < RA; r++)   // ~0.9 % of random pairs set their bit
                            pair_t2(a[r], b0, bR, bL, one, neg10, hits, 1u << (k * 8 + r));
                    }
                    s_mask[xq * NT + tid] = hits;
                    any |= hits;
                }
            } else {
                for (int xq = 0; xq < NQ; xq++) s_mask[xq * NT + tid] = 0xFFFFFFFFu;
                any = 1;
            }

            // ---------------- stage 2: exact D and S on this thread's own hit bits ----------------
            if (any) {
                int xq = 0;
                uint32_t m = s_mask[tid];
                for (;;) {
                    while (m == 0 && ++xq < NQ) m = s_mask[xq * NT + tid];
                    if (m == 0) break;
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t av = s_a[(j & 7) * NT + tid];
                    const uint32_t bv = s_b[0][xq * 4 + (j >> 3)];
                    const int d = exact_edge<MODE>(av, bv, w.t, w.T);
                    if (d > 0) emit_single(av, bv, d, out);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// nearest_kernel (a-7): every thread owns RA queries, the targets stream through shared memory in order,
// the t=2 prefilter guards the exact plain distance, and the running (distance, index) minimum lives in a
// register as one packed key so that "first strict minimum in caller order" is a plain unsigned min.
// gridDim.y splits the target list; the per-query keys are merged with atomicMin.
// ---------------------------------------------------------------------------------------------------
constexpr uint32_t NEAR_IDX_BITS = 28;
constexpr int NEAR_TB = 512;

template <bool SMALL>
__global__ void __launch_bounds__(NT, 2) nearest_kernel(const uint32_t* __restrict__ q, uint32_t Q,
                                                        const uint32_t* __restrict__ tg, uint32_t W, int max_d,
                                                        uint32_t w_per_block, uint32_t* __restrict__ keys)
{
    __shared__ __align__(16) uint32_t s_t[NEAR_TB];
    const int tid = threadIdx.x;
    uint32_t a[RA], best[RA];
#pragma unroll
    for (int r = 0; r < RA; r++) {
        const uint64_t idx = ((uint64_t)blockIdx.x * RA + r) * NT + tid;
        a[r] = idx < Q ? __ldg(&q[idx]) : 0u;
        best[r] = 0xFFFFFFFFu;
    }
    const uint32_t w_lo = blockIdx.y * w_per_block;
    const uint32_t w_hi = min(W, w_lo + w_per_block);
    for (uint32_t base = w_lo; base < w_hi; base += NEAR_TB) {
        __syncthreads();
        for (int x = tid; x < NEAR_TB; x += NT) s_t[x] = (base + x) < w_hi ? __ldg(&tg[base + x]) : 0u;
        __syncthreads();
        const int n = min((uint32_t)NEAR_TB, w_hi - base);
        for (int x = 0; x < n; x++) {
            const uint32_t b = s_t[x];
            const uint32_t bR = b >> 2, bL = b << 2;
#pragma unroll
            for (int r = 0; r < RA; r++) {
                int d;
                if (SMALL) {
                    const uint32_t x0 = a[r] ^ b, x1 = a[r] ^ bR, x2 = a[r] ^ bL;
                    const uint32_t m = ((x0 - F10_ONES) & ~x0) | ((x1 - F10_ONES) & ~x1) | ((x2 - F10_ONES) & ~x2);
                    if (!(m & F10_HIGH)) continue;
                    d = dist_small(a[r], b, true);
                } else {
                    d = myers3(a[r], b).full;
                }
                if (d <= max_d) best[r] = min(best[r], ((uint32_t)d << NEAR_IDX_BITS) | (base + x));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RA; r++) {
        const uint64_t idx = ((uint64_t)blockIdx.x * RA + r) * NT + tid;
        if (idx < Q && best[r] != 0xFFFFFFFFu) atomicMin(&keys[idx], best[r]);
    }
}

__global__ void nearest_finish_kernel(const uint32_t* __restrict__ keys, uint32_t Q, int32_t* __restrict__ argmin,
                                      uint8_t* __restrict__ dist)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    const uint32_t k = keys[i];
    if (k == 0xFFFFFFFFu) { argmin[i] = -1; dist[i] = 255; }
    else { argmin[i] = (int32_t)(k & ((1u << NEAR_IDX_BITS) - 1u)); dist[i] = (uint8_t)(k >> NEAR_IDX_BITS); }
}

// ---------------------------------------------------------------------------------------------------
// kmer_score_kernel (a-5): thread per whitelist entry, queries broadcast from shared memory; exact S with
// per-position multiplicities; hits appended through a warp-aggregated cursor.
// ---------------------------------------------------------------------------------------------------
constexpr int KS_QB = 256;

__global__ void __launch_bounds__(NT) kmer_score_kernel(const uint32_t* __restrict__ q, uint32_t Q,
                                                        const uint32_t* __restrict__ wl, uint32_t W, int min_kmers,
                                                        unsigned long long cap, uint32_t* __restrict__ hit_q,
                                                        uint32_t* __restrict__ hit_w, uint8_t* __restrict__ cnt,
                                                        unsigned long long* __restrict__ mult,
                                                        unsigned long long* __restrict__ total)
{
    __shared__ uint32_t s_q[KS_QB];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const uint64_t wi = (uint64_t)blockIdx.x * NT + tid;
    const uint32_t b = wi < W ? __ldg(&wl[wi]) : 0u;
    const uint32_t q_lo = blockIdx.y * KS_QB;
    const int nq = min((uint32_t)KS_QB, Q - q_lo);
    for (int x = tid; x < nq; x += NT) s_q[x] = __ldg(&q[q_lo + x]);
    __syncthreads();
    for (int x = 0; x < nq; x++) {
        const uint32_t a = s_q[x];
        const int s = wi < W ? qgram_score(a, b) : 0;
        const bool ok = wi < W && s >= min_kmers && s > 0;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m == 0) continue;
        unsigned long long base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(total, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ok) {
            const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) {
                uint64_t mu;
                qgram_score(a, b, &mu);
                hit_q[pos] = q_lo + x;
                hit_w[pos] = (uint32_t)wi;
                cnt[pos] = (uint8_t)s;
                mult[pos] = mu;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// member_kernel (a-6): 1024 evenly spaced pivots of the sorted whitelist in shared memory give the first
// ten levels of the search; the remaining levels touch a window of W/1024 entries (L2-resident: the 3 M
// whitelist is 12 MB).  4 B in + 1 B out per query.
// ---------------------------------------------------------------------------------------------------
constexpr int MEM_PIV = 1024;

__global__ void __launch_bounds__(NT) member_kernel(const uint32_t* __restrict__ wl, uint32_t W,
                                                    const uint32_t* __restrict__ q, uint32_t Q, uint8_t* __restrict__ hit)
{
    __shared__ uint32_t s_p[MEM_PIV];
    const uint64_t step = ((uint64_t)W + MEM_PIV - 1) / MEM_PIV;   // pivot p = wl[min(W-1, p*step)]
    for (int p = threadIdx.x; p < MEM_PIV; p += NT) {
        const uint64_t i = (uint64_t)p * step;
        s_p[p] = __ldg(&wl[i < W ? i : W - 1]);
    }
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * NT + threadIdx.x; i < Q; i += (uint64_t)gridDim.x * NT) {
        const uint32_t v = __ldg(&q[i]);
        int lo = 0, hi = MEM_PIV;   // largest pivot index with s_p[p] <= v (or 0)
#pragma unroll
        for (int it = 0; it < 10; it++) {
            const int mid = (lo + hi) >> 1;
            if (s_p[mid] <= v) lo = mid; else hi = mid;
        }
        uint64_t l = (uint64_t)lo * step, h = min((uint64_t)W, l + step);   // candidates in [l, h)
        if (l >= W) { l = W - 1; h = W; }
        while (l < h) {
            const uint64_t mid = (l + h) >> 1;
            if (__ldg(&wl[mid]) < v) l = mid + 1; else h = mid;
        }
        hit[i] = (l < W && __ldg(&wl[l]) == v) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// pack16_kernel (a-1): one 128-bit load per read, 16 B in + 5 B out.  code = ((c>>1) ^ (c>>2)) & 3 maps
// A,C,G,T -> 0,1,2,3; validity is an exact match against the four upper-case letters.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack4(uint32_t wrd, int base_pos, uint32_t& r, bool& ok)
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t c = (wrd >> (8 * i)) & 0xFFu;
        ok = ok && (c == 'A' || c == 'C' || c == 'G' || c == 'T');
        r |= (((c >> 1) ^ (c >> 2)) & 3u) << (2 * (base_pos + i));
    }
}

__global__ void __launch_bounds__(NT) pack16_kernel(const uint4* __restrict__ seqs, uint64_t R, uint32_t* __restrict__ out,
                                                    uint8_t* __restrict__ valid)
{
    for (uint64_t i = (uint64_t)blockIdx.x * NT + threadIdx.x; i < R; i += (uint64_t)gridDim.x * NT) {
        const uint4 v = __ldg(&seqs[i]);
        uint32_t r = 0;
        bool ok = true;
        pack4(v.x, 0, r, ok); pack4(v.y, 4, r, ok); pack4(v.z, 8, r, ok); pack4(v.w, 12, r, ok);
        out[i] = r;
        valid[i] = ok ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------
// pipe_probe_kernel: 64 independent instructions per loop trip (8 chains x 8), inline PTX so that ptxas
// keeps the opcode.  kind 0 LOP3, 1 IMAD, 2 alternating LOP3/IMAD, 3 POPC(+LOP3 to keep chains alive).
// ---------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(NT) pipe_probe_kernel(int iters, uint32_t* __restrict__ sink)
{
    uint32_t v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 2654435761u + i * 40503u + blockIdx.x;
    uint32_t c = sink[0] | 0x9E3779B9u, m = sink[1] | 5u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (KIND == 0 || (KIND == 2 && (u & 1) == 0))
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(c), "r"(m));
                else if (KIND == 1 || KIND == 2)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(m), "r"(c));
                else
                    asm volatile("popc.b32 %0, %0;" : "+r"(v[i]));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= v[i];
    if (s == 0x12345678u) sink[2] = s;   // practically never; keeps the chains live
}

}  // namespace bdg
