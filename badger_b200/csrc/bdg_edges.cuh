// bdg_edges.cuh -- the edge-construction kernel (sm_100a): reference index.py:77-93 (candidate filter) +
// barcode_graph.py:224-249 (3-way edit-distance verify, emit) over every unordered pair of a SORTED array of
// distinct packed barcodes.
//
// Shape of the computation
//   * worker = one WARP.  Workers pull (row group, column chunk) items from an atomic counter: a row group is
//     256 consecutive rows (lane l holds rows row0 + r*32 + l, r = 0..7, in registers), a chunk is a run of
//     columns right of the group's first row.  No block-level barrier anywhere: a warp stages its own column
//     sub-tiles in its own slice of shared memory (__syncwarp only), so warps never wait for one another.
//   * per column sub-tile (SB columns) the warp first asks whether ANY pair of the tile can meet one of the
//     "top" conditions of the prefilter (bdg_core.cuh: interval test on the first/last row and column - the
//     array is sorted, so the high bits of a tile's rows and columns barely move).  ~99 % of the tiles cannot:
//     they run the LIGHT loop, which evaluates only the remaining conditions, pair by pair, in 1.25 (t=1) or
//     6 (t=2) integer instructions per pair, split over the ALU pipe (LOP3) and the FMA pipe (IMAD).  The rest
//     (tiles next to the diagonal and the few whose top fields line up) run the FULL prefilter.
//   * stage 1 leaves one hit bit per pair in registers (32 pairs per word: 4 columns x 8 rows); every
//     16 columns the warp votes, and lanes with hits append (row, column) codes to the warp's candidate queue
//     in shared memory.
//   * stage 2 is warp-cooperative: whenever the queue holds >= 32 candidates, each lane takes one, evaluates
//     D and S exactly (bdg_core.cuh) and the warp appends the edges it found with one atomic.  The rare,
//     expensive exact test therefore runs with 32 busy lanes instead of one.
// MODE 1: t = 1.  MODE 2: t = 2.  MODE 3: any t: the generic quick test (quick_pass_any) in front of the bit-vector distance.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bdg_core.cuh"

namespace bdg {

constexpr int EW = 8;              // warps per CTA
constexpr int ENT = EW * 32;       // threads per CTA
constexpr int RA = 8;              // rows per lane
constexpr int GROUP = 32 * RA;     // rows per work item (one warp)
constexpr int ROW_TILE = 2048;     // == BDG_ROW_TILE: rows are dealt to parts in tiles of 8 groups
constexpr int QCAP = 256;          // candidate queue entries per warp
constexpr unsigned FULL = 0xffffffffu;

template <int MODE> struct EdgeCfg { static constexpr int SB = (MODE == 1) ? 256 : 128; };   // columns per sub-tile
constexpr int SB_MAX = 256;

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
    const uint32_t* group_ids;  // K row groups (GROUP rows each) owned by this part
    const uint32_t* item_start; // K+1 prefix sums of column chunks per owned group
    uint32_t K;
    uint32_t n_items;
    uint32_t chunk_cols;        // columns per work item (multiple of SB_MAX, <= 2^17)
    unsigned int* item_counter; // dynamic scheduler
    unsigned long long* stats;  // optional [8]: sub-tiles visited, sub-tiles scored pair by pair, pairs scored, candidates,
                                //               sum / max over warps of the warp's busy time in ns, (unused), pairs that reached S
    uint32_t one;               // == 1, opaque to the compiler: x*(-one)+c keeps the subtraction on the FMA pipe (IMAD)
    uint32_t mone;              // == 0xFFFFFFFF, a SEPARATE opaque value: u*one + mone is u - 1 as one IMAD (derived from `one`
                                // the compiler would rewrite it as (u-1)*one, an add on the ALU pipe plus a multiply)
    int pass;                   // sparse kernel: pass index p (bdg_core.cuh pass_pred); `sorted` holds rotl(key, rot) sorted
    int rot;
    // bipartite form (queries x targets, bdg_nearest_bounded): rows come from `sorted` (N queries), columns from `cols`
    const uint32_t* cols;       // NC target keys sorted by the same rotated key
    uint32_t NC;
    const uint32_t* row_pay;    // original index of every sorted query / target
    const uint32_t* col_pay;
    uint32_t* near_keys;        // per query: min over targets of (plain distance << 28 | target index)
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// warp exit bookkeeping for the load-balance statistics: sum and max over the warps of the time each one was busy
__device__ __forceinline__ void warp_exit_stats(unsigned long long* stats, unsigned long long t_start)
{
    const unsigned long long dt = global_ns() - t_start;
    atomicAdd(&stats[4], dt);
    atomicMax(&stats[5], dt);
}

// Next work item of the dynamic scheduler: (k, j) = owned group index, column chunk.  The position of the item in
// the prefix sums is found by a 32-ary search (one probe per lane and step: 3 round trips for 2k groups, 4 for 200k).
__device__ __forceinline__ bool fetch_item(const EdgeWork& w, int lane, uint32_t& k, uint32_t& j)
{
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(w.item_counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= w.n_items) return false;
    uint32_t lo = 0, hi = w.K;                        // item_start[lo] <= item < item_start[hi]
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 32) / 33;
        const uint32_t pos = lo + (uint32_t)(lane + 1) * step;
        const bool le = pos < hi && __ldg(&w.item_start[pos]) <= item;
        const uint32_t c = (uint32_t)__popc(__ballot_sync(FULL, le));   // item_start is increasing: a prefix of the lanes
        hi = min(hi, lo + (c + 1) * step);
        lo = lo + c * step;
    }
    k = lo;
    j = item - __ldg(&w.item_start[lo]);
    return true;
}

// ---- output: warp-aggregated append (one atomic per warp that has anything to emit) ---------------
__device__ __forceinline__ void emit_warp(bool ok, uint32_t a, uint32_t b, int d, const EdgeOut& out)
{
    const unsigned m = __ballot_sync(FULL, ok);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(out.count, (unsigned long long)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (ok) {
        const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < out.cap) {
            out.a[pos] = a;
            out.b[pos] = b;
            out.d[pos] = (uint8_t)d;
        }
    }
}

// exact stage: D (case analysis for t<=2, bit-vector pass otherwise), then S only for survivors
template <int MODE>
__device__ __forceinline__ int exact_edge(uint32_t a, uint32_t b, int t, int T)
{
    if (!(a < b)) return 0;   // rows are sorted and distinct: index order == value order
    const int d = (MODE == 3) ? dist3_min(a, b) : dist_small(a, b);
    if (d > t) return 0;
    return qgram_score(a, b) >= T ? d : 0;
}

__device__ __forceinline__ uint32_t pick4(const uint4& v, int k)
{
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// ---- stage-1 inner steps in PTX (so that the subtraction is an IMAD and the hit bit ONE predicated LOP3) ----
// t=2 light, one pair: 2 XOR + 2 IMAD + LOP3(->predicate) + predicated OR.  lop3 0xA8 = (A | B) & C.
__device__ __forceinline__ void pair_t2_light(uint32_t aA, uint32_t aB, uint32_t bA, uint32_t bB, uint32_t mone, uint32_t guard,
                                              uint32_t& hits, const uint32_t bit)
{
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 xA, xB, yA, yB, m;\n\t"
        "xor.b32 xA, %1, %3;\n\t"
        "xor.b32 xB, %2, %4;\n\t"
        "mad.lo.u32 yA, xA, %5, %6;\n\t"
        "mad.lo.u32 yB, xB, %5, %6;\n\t"
        "lop3.b32 m, yA, yB, %7, 0xA8;\n\t"
        "setp.ne.u32 p, m, 0;\n\t"
        "@p or.b32 %0, %0, %8;\n\t"
        "}"
        : "+r"(hits)
        : "r"(aA), "r"(aB), "r"(bA), "r"(bB), "r"(mone), "r"(guard), "n"(T2_GUARD), "r"(bit));
}

// t=1 light, one row against two packed words (= 4 columns): 2 XOR + 2 IMAD + 1 LOP3.  lop3 0xFE = A | B | C.
constexpr uint32_t T1L_GUARD = 0x80008000u;
__device__ __forceinline__ void quad_t1_light(uint32_t aw, uint32_t p0, uint32_t p1, uint32_t mone, uint32_t guard, uint32_t& acc)
{
    asm("{\n\t"
        ".reg .b32 x0, x1, y0, y1;\n\t"
        "xor.b32 x0, %1, %2;\n\t"
        "xor.b32 x1, %1, %3;\n\t"
        "mad.lo.u32 y0, x0, %4, %5;\n\t"
        "mad.lo.u32 y1, x1, %4, %5;\n\t"
        "lop3.b32 %0, y0, y1, %0, 0xFE;\n\t"
        "}"
        : "+r"(acc)
        : "r"(aw), "r"(p0), "r"(p1), "r"(mone), "r"(guard));
}

// ---- per-warp context -------------------------------------------------------------------------------
struct WarpCtx {
    uint32_t* q;           // candidate queue (shared)
    int* qn;               // its fill count (shared)
    const uint32_t* sorted;
    uint32_t N;
    uint64_t row0, col_lo, col_hi;
    int t, T, lane;
};

// stage 2 on full batches of 32 candidates (all == true: also the partial rest)
template <int MODE>
__device__ __forceinline__ void drain(const WarpCtx& c, const EdgeOut& out, bool all)
{
    __syncwarp();
    int n = *(volatile int*)c.qn;
    n = n < QCAP ? n : QCAP;
    __syncwarp();
    while (n >= 32 || (all && n > 0)) {
        const int take = n >= 32 ? 32 : n;
        n -= take;
        bool ok = false;
        uint32_t a = 0, b = 0;
        int d = 0;
        if (c.lane < take) {
            const uint32_t e = c.q[n + c.lane];
            const uint64_t row = c.row0 + (e & 255u), col = c.col_lo + (e >> 8);
            if (row < c.N && col < c.col_hi) {
                a = __ldg(&c.sorted[row]);
                b = __ldg(&c.sorted[col]);
                d = exact_edge<MODE>(a, b, c.t, c.T);
                ok = d > 0;
            }
        }
        emit_warp(ok, a, b, d, out);
    }
    if (c.lane == 0) *c.qn = n;
    __syncwarp();
}

// append the set bits of h[0..3] (bit k*8+r of h[s] = column colrel + 4s + k, row r*32+lane) to the queue
template <int MODE>
__device__ __forceinline__ void push_hits(uint32_t (&h)[4], uint32_t colrel, const WarpCtx& c, const EdgeOut& out)
{
    do {
        bool room = true;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            while (room && h[s]) {
                const int j = __ffs(h[s]) - 1;
                const int slot = atomicAdd(c.qn, 1);
                if (slot >= QCAP) { room = false; break; }
                c.q[slot] = ((colrel + 4 * s + (j >> 3)) << 8) | (uint32_t)((j & 7) * 32 + c.lane);
                h[s] &= h[s] - 1;
            }
        }
        drain<MODE>(c, out, false);
    } while (__any_sync(FULL, (h[0] | h[1] | h[2] | h[3]) != 0));
}

template <int MODE>
__global__ void __launch_bounds__(ENT, 3) edges_kernel(const EdgeWork w, const EdgeOut out)
{
    constexpr int SB = EdgeCfg<MODE>::SB;
    constexpr int NB = SB / 32;                        // staged columns per lane
    __shared__ __align__(16) uint32_t s_raw[EW][SB];                          // the sub-tile's barcodes
    __shared__ __align__(16) uint32_t s_w0[EW][MODE == 1 ? SB / 2 : (MODE == 2 ? SB : 4)];   // t1: packed low-15 pairs; t2: word A
    __shared__ __align__(16) uint32_t s_w1[EW][MODE == 2 ? SB : 4];          // t2: word B
    __shared__ uint32_t s_q[EW][QCAP];
    __shared__ int s_qn[EW];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* const raw = s_raw[wid];
    uint32_t* const w0 = s_w0[wid];
    uint32_t* const w1 = s_w1[wid];
    const uint32_t mone = 0u - w.one;                                         // runtime -1
    const uint32_t g2 = T2_GUARD * w.one, g1 = T1L_GUARD * w.one;             // runtime guards: keep IMAD from folding
    (void)g1; (void)g2; (void)w0; (void)w1;

    WarpCtx c;
    c.q = s_q[wid]; c.qn = &s_qn[wid]; c.sorted = w.sorted; c.N = w.N; c.t = w.t; c.T = w.T; c.lane = lane;
    if (lane == 0) s_qn[wid] = 0;
    __syncwarp();
    const unsigned long long t_start = global_ns();
    unsigned long long n_sub = 0, n_full = 0, n_cand = 0;

    for (;;) {
        uint32_t k = 0, j = 0;
        if (!fetch_item(w, lane, k, j)) break;
        c.row0 = (uint64_t)__ldg(&w.group_ids[k]) * GROUP;
        c.col_lo = c.row0 + (uint64_t)j * w.chunk_cols;
        c.col_hi = min((uint64_t)w.N, c.col_lo + w.chunk_cols);

        uint32_t a[RA];
        uint32_t aw0[RA], aw1[RA];   // light-loop row words (t1: aw0 only)
#pragma unroll
        for (int r = 0; r < RA; r++) {
            const uint64_t idx = c.row0 + (uint64_t)r * 32 + lane;
            a[r] = idx < w.N ? __ldg(&w.sorted[idx]) : 0xFFFFFFFFu;
            if constexpr (MODE == 1) { aw0[r] = (a[r] & 0x7FFFu) * 0x00010001u; aw1[r] = 0; }
            else if constexpr (MODE == 2) { aw0[r] = t2_word_aA(a[r]); aw1[r] = t2_word_aB(a[r]); }
            else { aw0[r] = aw1[r] = 0; }
        }
        const uint32_t a_lo = __ldg(&w.sorted[c.row0]);
        const uint32_t a_hi = __ldg(&w.sorted[min((uint64_t)w.N, c.row0 + GROUP) - 1]);

        uint32_t nb[NB];             // prefetched columns of the next sub-tile
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const uint64_t idx = c.col_lo + (uint64_t)i * 32 + lane;
            nb[i] = idx < c.col_hi ? __ldg(&w.sorted[idx]) : 0u;
        }

        for (uint64_t sub = c.col_lo; sub < c.col_hi; sub += SB) {
            __syncwarp();            // every lane is done with the previous sub-tile's words
#pragma unroll
            for (int i = 0; i < NB; i++) {
                raw[i * 32 + lane] = nb[i];
                if constexpr (MODE == 2) { w0[i * 32 + lane] = t2_word_bA(nb[i]); w1[i * 32 + lane] = t2_word_bB(nb[i]); }
            }
            __syncwarp();
            if constexpr (MODE == 1) {
#pragma unroll
                for (int i = 0; i < NB / 2; i++) {
                    const uint2 p = *reinterpret_cast<const uint2*>(&raw[2 * (i * 32 + lane)]);
                    w0[i * 32 + lane] = (p.x & 0x7FFFu) | ((p.y & 0x7FFFu) << 16);
                }
                __syncwarp();
            }
            if (sub + SB < c.col_hi) {
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    const uint64_t idx = sub + SB + (uint64_t)i * 32 + lane;
                    nb[i] = idx < c.col_hi ? __ldg(&w.sorted[idx]) : 0u;
                }
            }
            const int ncols = (int)min((uint64_t)SB, c.col_hi - sub);
            const uint32_t colrel0 = (uint32_t)(sub - c.col_lo);
            bool full = true;
            if constexpr (MODE == 1) full = t1_top_possible(a_lo, a_hi, raw[0], raw[ncols - 1]);
            if constexpr (MODE == 2) full = t2_top_possible(a_lo, a_hi, raw[0], raw[ncols - 1]);
            n_sub++; n_full += full ? 1 : 0;

            if (!full) {
                // ------------------------------ light loop: 16 columns per vote ------------------------------
#pragma unroll 1
                for (int cb = 0; cb < SB; cb += 16) {
                    uint32_t h[4] = {0u, 0u, 0u, 0u};
                    if constexpr (MODE == 1) {
                        const uint4 P0 = *reinterpret_cast<const uint4*>(&w0[cb / 2]);
                        const uint4 P1 = *reinterpret_cast<const uint4*>(&w0[cb / 2 + 4]);
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const uint32_t p0 = s < 2 ? pick4(P0, 2 * s) : pick4(P1, 2 * s - 4);
                            const uint32_t p1 = s < 2 ? pick4(P0, 2 * s + 1) : pick4(P1, 2 * s - 3);
                            uint32_t acc = 0;
#pragma unroll
                            for (int r = 0; r < RA; r++) quad_t1_light(aw0[r], p0, p1, mone, g1, acc);
                            if (acc & T1L_GUARD) {      // ~1e-3 per lane and step on random data
                                const uint4 B = *reinterpret_cast<const uint4*>(&raw[cb + 4 * s]);
#pragma unroll
                                for (int kk = 0; kk < 4; kk++)
#pragma unroll
                                    for (int r = 0; r < RA; r++)
                                        if (t1_light(a[r], pick4(B, kk))) h[s] |= 1u << (kk * 8 + r);
                            }
                        }
                    } else if constexpr (MODE == 2) {
#pragma unroll
                        for (int s = 0; s < 4; s++) {
                            const uint4 A4 = *reinterpret_cast<const uint4*>(&w0[cb + 4 * s]);
                            const uint4 B4 = *reinterpret_cast<const uint4*>(&w1[cb + 4 * s]);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++)
#pragma unroll
                                for (int r = 0; r < RA; r++)
                                    pair_t2_light(aw0[r], aw1[r], pick4(A4, kk), pick4(B4, kk), mone, g2, h[s], 1u << (kk * 8 + r));
                        }
                    }
                    if (__any_sync(FULL, (h[0] | h[1] | h[2] | h[3]) != 0)) {
                        n_cand += __popc(h[0]) + __popc(h[1]) + __popc(h[2]) + __popc(h[3]);
                        push_hits<MODE>(h, colrel0 + cb, c, out);
                    }
                }
            } else {
                // ------------------------------ full prefilter (rare tiles; every pair for MODE 3) -----------
#pragma unroll 1
                for (int cb = 0; cb < SB; cb += 4) {
                    const uint4 B = *reinterpret_cast<const uint4*>(&raw[cb]);
                    uint32_t h[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const uint32_t b = pick4(B, kk);
#pragma unroll
                        for (int r = 0; r < RA; r++) {
                            const bool hit = MODE == 1 ? prefilter_t1(a[r], b) : (MODE == 2 ? prefilter_t2(a[r], b) : quick_pass_any(a[r], b, w.t));
                            h[0] |= (hit ? 1u : 0u) << (kk * 8 + r);
                        }
                    }
                    if (__any_sync(FULL, h[0] != 0)) { n_cand += __popc(h[0]); push_hits<MODE>(h, colrel0 + cb, c, out); }
                }
            }
        }
        drain<MODE>(c, out, true);   // the queue's codes are relative to this item: empty it before the next one
    }
    if (w.stats) {
        for (int o = 16; o; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);           // per-lane counts
        if (lane == 0) {                                                                  // n_sub / n_full are uniform per warp
            atomicAdd(&w.stats[0], n_sub); atomicAdd(&w.stats[1], n_full);
            atomicAdd(&w.stats[2], n_sub * (unsigned long long)(SB * GROUP)); atomicAdd(&w.stats[3], n_cand);
            warp_exit_stats(w.stats, t_start);
        }
    }
}


// =====================================================================================================
// Sparse multi-pass form (bdg_core.cuh "Multi-pass").  `sorted` is the array sorted by the pass's rotated key.
// Three small kernels per pass:
//   tile_bounds_kernel   first / last key of every 128-column sub-tile (8 B per sub-tile, stays in L2)
//   sparse_scan_kernel   LEVEL 1: one interval test (pass_possible) per (row group of 256 rows, sub-tile right of
//                        it); the (group, sub-tile) pairs that can hold a candidate are appended to a compact
//                        tile list.  Pure streaming, every pair of the matrix is decided here or handed on.
//   sparse_tile_kernel   persistent warps pull tiles from the list (small uniform units: no load imbalance):
//                        LEVEL 2 - lane l tests (32-row slab l&7, 32-column quarter l>>3) -> mask of 32x32 blocks;
//                        inside those every pair gets the quick test of bdg_core.cuh (<= t columns mismatching
//                        on all three diagonals: ~9 ALU-pipe + 5 FMA-pipe instructions), hit bits -> the warp's
//                        candidate queue (ballot/popc slots) -> warp-cooperative exact stage whenever 32 wait.
// The exact stage keeps a candidate only if THIS pass's predicate holds and no EARLIER pass's does, so the
// passes' outputs are disjoint and simply share one output cursor.
// =====================================================================================================
constexpr int SSB = 128;           // columns per sub-tile
constexpr int SQCAP = 128;         // entries per warp of the candidate queue and of the second (scoring) queue

struct TileList {
    uint2* tiles;                  // (group, sub-tile)
    unsigned long long* count;     // tiles appended (may exceed cap: the host then grows the list and rescans)
    unsigned long long cap;
};

__global__ void tile_bounds_kernel(const uint32_t* __restrict__ sorted, uint32_t N, uint2* __restrict__ bnd, uint32_t NS)
{
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < NS; s += gridDim.x * blockDim.x) {
        const uint64_t c0 = (uint64_t)s * SSB, c1 = min((uint64_t)N, c0 + SSB) - 1;
        bnd[s] = make_uint2(__ldg(&sorted[c0]), __ldg(&sorted[c1]));
    }
}

template <int T_, int P_, bool BIP>
__global__ void __launch_bounds__(256) sparse_scan_kernel(const uint32_t* __restrict__ sorted, uint32_t N,
                                                          const uint32_t* __restrict__ group_ids, uint32_t K,
                                                          const uint2* __restrict__ bnd, uint32_t NS, TileList list)
{
    const int lane = threadIdx.x & 31;
    for (uint32_t k = blockIdx.x; k < K; k += gridDim.x) {
        const uint32_t g = group_ids ? __ldg(&group_ids[k]) : k;
        const uint64_t row0 = (uint64_t)g * GROUP;
        const uint32_t a_lo = __ldg(&sorted[row0]);
        const uint32_t a_hi = __ldg(&sorted[min((uint64_t)N, row0 + GROUP) - 1]);
        const uint32_t s0 = BIP ? 0u : g * (GROUP / SSB);         // triangular: first sub-tile that reaches past the group's first row
        for (uint32_t sb = s0; sb < NS; sb += blockDim.x) {        // whole warps stay in the loop together
            const uint32_t sidx = sb + threadIdx.x;
            bool poss = false;
            if (sidx < NS) {
                const uint2 b = __ldg(&bnd[sidx]);
                poss = pass_possible(T_, P_, a_lo, a_hi, b.x, b.y);
            }
            const unsigned m = __ballot_sync(FULL, poss);
            if (m) {
                unsigned long long base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(list.count, (unsigned long long)__popc(m));
                base = __shfl_sync(FULL, base, leader);
                const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
                if (poss && pos < list.cap) list.tiles[pos] = make_uint2(g, sidx);
            }
        }
    }
}

struct SparseCtx {
    uint2* q;                      // candidate queue: (row, column) indices into `sorted`
    uint2* q2;                     // second queue: (x, y) values with D <= t that still need S
    uint8_t* q2d;                  // their distances
    const uint32_t* sorted;
    uint32_t N;
    int t, T, rot, lane;
    const uint32_t* cols;          // bipartite form only
    uint32_t NC;
    const uint32_t* row_pay;
    const uint32_t* col_pay;
    uint32_t* near_keys;
    unsigned long long* n_score;   // this warp's count of pairs that reached stage 3 (statistics)
};

// stage 3 on a batch: S = shared 6-mer count (the dearest test: 21 diagonals) for pairs that already have D <= t
__device__ __forceinline__ void sparse_score(const SparseCtx& c, const EdgeOut& out, int n)
{
    bool ok = false;
    uint32_t x = 0, y = 0;
    int d = 0;
    if (c.lane < n) {
        const uint2 e = c.q2[c.lane];
        x = e.x; y = e.y; d = c.q2d[c.lane];
        ok = qgram_score(x, y) >= c.T;
    }
    emit_warp(ok, x, y, d, out);
}

// stage 2 on a batch of candidates: pass predicates (this pass yes, earlier passes no) and exact D; survivors
// wait in the second queue so that stage 3 also runs with full warps
// bipartite stage 2: plain edit distance of (query, target); the best (distance, target index) per query is kept with
// an atomicMin, which also absorbs pairs that several passes find
__device__ __forceinline__ void near_process(const SparseCtx& c, uint2 e, bool active)
{
    if (active && e.x < c.N && e.y < c.NC) {
        const uint32_t x = rotr32(__ldg(&c.sorted[e.x]), c.rot);
        const uint32_t y = rotr32(__ldg(&c.cols[e.y]), c.rot);
        const int d = dist_small(x, y, true);
        if (d <= c.t) atomicMin(&c.near_keys[__ldg(&c.row_pay[e.x])], ((uint32_t)d << 28) | __ldg(&c.col_pay[e.y]));
    }
}

template <int T_, int P_, bool BIP>
__device__ __forceinline__ void sparse_process(const SparseCtx& c, const EdgeOut& out, uint2 e, bool active, int& q2n)
{
    if constexpr (BIP) { near_process(c, e, active); return; }
    bool ok = false;
    uint32_t x = 0, y = 0;
    int d = 0;
    if (active && e.x < e.y && e.y < c.N) {             // each unordered pair once: row index < column index
        const uint32_t xr = __ldg(&c.sorted[e.x]), yr = __ldg(&c.sorted[e.y]);
        x = rotr32(xr, c.rot);
        y = rotr32(yr, c.rot);
        if (x > y) { const uint32_t tmp = x; x = y; y = tmp; }
        bool mine = pass_pred_rot(T_, P_, xr, yr);      // found by this pass (predicates are orientation-free) and by no earlier one
#pragma unroll
        for (int q = 0; q < P_; q++) mine = mine && !pass_pred(T_, q, x, y);
        if (mine) {
            d = dist_small(x, y);
            ok = d <= c.t;
        }
    }
    const unsigned m = __ballot_sync(FULL, ok);
    if (m == 0) return;
    *c.n_score += __popc(m);
    // fewer than 32 entries wait on entry (full batches are scored right below), so 32 more always fit
    if (ok) {
        const int slot = q2n + __popc(m & ((1u << c.lane) - 1u));
        c.q2[slot] = make_uint2(x, y);
        c.q2d[slot] = (uint8_t)d;
    }
    q2n += __popc(m);
    __syncwarp();
    while (q2n >= 32) {                                 // stage 3 on full batches, taken from the top
        q2n -= 32;
        const uint2 mv = c.q2[q2n + c.lane];
        const uint8_t md = c.q2d[q2n + c.lane];
        const bool good = qgram_score(mv.x, mv.y) >= c.T;
        emit_warp(good, mv.x, mv.y, md, out);
        __syncwarp();
    }
}

// append the set bits of h (bit k*8+r = column col0+k, row row0 + r*32 + lane); run stage 2 whenever 32 candidates wait
template <int T_, int P_, bool BIP>
__device__ __forceinline__ void sparse_push(uint32_t h, uint32_t row0, uint32_t col0, int& qn, int& q2n, const SparseCtx& c, const EdgeOut& out)
{
    // common case: every lane writes all its hits at once at offsets from a warp prefix sum of the hit counts
    const int cnt = __popc(h);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (c.lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (qn + total <= SQCAP) {
        int slot = qn + incl - cnt;
        while (h) {
            const int j = __ffs(h) - 1;
            h &= h - 1;
            c.q[slot++] = make_uint2(row0 + (uint32_t)((j & 7) * 32 + c.lane), col0 + (uint32_t)(j >> 3));
        }
        qn += total;
        __syncwarp();
        while (qn >= 32) {
            qn -= 32;
            const uint2 e = c.q[qn + c.lane];
            sparse_process<T_, P_, BIP>(c, out, e, true, q2n);
            __syncwarp();
        }
        return;
    }
    // burst larger than the queue: one hit per lane and round
    for (;;) {
        const bool has = h != 0;
        const unsigned m = __ballot_sync(FULL, has);
        if (m == 0) break;
        if (has) {
            const int j = __ffs(h) - 1;
            h &= h - 1;
            c.q[qn + __popc(m & ((1u << c.lane) - 1u))] = make_uint2(row0 + (uint32_t)((j & 7) * 32 + c.lane), col0 + (uint32_t)(j >> 3));
        }
        qn += __popc(m);
        __syncwarp();
        while (qn >= 32) {
            qn -= 32;
            const uint2 e = c.q[qn + c.lane];
            sparse_process<T_, P_, BIP>(c, out, e, true, q2n);
            __syncwarp();
        }
    }
}

template <int T_, int P_, bool BIP>
__global__ void __launch_bounds__(ENT, 4) sparse_tile_kernel(const EdgeWork w, const EdgeOut out, const TileList list)
{
    __shared__ __align__(16) uint32_t s_b[EW][3][SSB];   // unrotated b, b >> 2, b << 2 of the staged sub-tile
    __shared__ uint2 s_q[EW][SQCAP];
    __shared__ uint2 s_q2[EW][SQCAP];
    __shared__ uint8_t s_q2d[EW][SQCAP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* const b0s = s_b[wid][0];
    uint32_t* const bPs = s_b[wid][1];
    uint32_t* const bMs = s_b[wid][2];
    SparseCtx c;
    c.q = s_q[wid]; c.q2 = s_q2[wid]; c.q2d = s_q2d[wid];
    c.sorted = w.sorted; c.N = w.N; c.t = w.t; c.T = w.T; c.rot = w.rot; c.lane = lane;
    c.cols = BIP ? w.cols : w.sorted; c.NC = BIP ? w.NC : w.N; c.row_pay = w.row_pay; c.col_pay = w.col_pay; c.near_keys = w.near_keys;
    int qn = 0, q2n = 0;                          // queue fills, uniform across the warp
    const uint32_t mone = w.mone;                 // runtime -1: u*one + mone is u - 1 on the FMA pipe
    const unsigned long long t_start = global_ns();
    unsigned long long n_combo = 0, n_cand = 0, n_score = 0;
    c.n_score = &n_score;
    const unsigned long long n_listed = *list.count;
    const unsigned long long n_tiles = min(n_listed, list.cap);
    if (n_listed > list.cap && !BIP && blockIdx.x == 0 && threadIdx.x == 0)
        atomicOr(out.count, 1ull << 63);          // the tile list was too small: poison the edge count, the host grows the list and reruns
    uint32_t g_cur = 0xFFFFFFFFu;
    uint32_t a[RA];                               // UNROTATED rows of the current group
    uint32_t my_alo = 0, my_ahi = 0;              // rotated end points of this lane's slab (lane & 7)

    for (;;) {
        unsigned long long ti = 0;
        if (lane == 0) ti = atomicAdd((unsigned long long*)w.item_counter, 1ull);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= n_tiles) break;
        const uint2 tile = list.tiles[ti];
        const uint32_t row0 = tile.x * GROUP;
        if (tile.x != g_cur) {                    // consecutive list entries mostly share the group
            g_cur = tile.x;
#pragma unroll
            for (int r = 0; r < RA; r++) {
                const uint32_t idx = row0 + (uint32_t)r * 32 + lane;
                const uint32_t ar = idx < w.N ? __ldg(&w.sorted[idx]) : 0xFFFFFFFFu;    // pad rows only widen the interval;
                a[r] = rotr32(ar, w.rot);                                               // they are dropped in sparse_process
                const uint32_t lo = __shfl_sync(FULL, ar, 0), hi = __shfl_sync(FULL, ar, 31);
                if ((lane & 7) == r) { my_alo = lo; my_ahi = hi; }
            }
        }
        const uint32_t sub = tile.y * SSB;
        const int ncols = (int)min((uint32_t)SSB, c.NC - sub);
        __syncwarp();
        uint32_t my_blo = 0, my_bhi = 0;
#pragma unroll
        for (int i = 0; i < SSB / 32; i++) {                     // quarter i = columns sub + 32i .. +31
            const uint32_t idx = sub + (uint32_t)i * 32 + lane;
            const uint32_t br = idx < c.NC ? __ldg(&c.cols[idx]) : 0xFFFFFFFFu;    // pads only widen the interval
            const uint32_t b = idx < c.NC ? rotr32(br, w.rot) : 0u;
            b0s[i * 32 + lane] = b;
            bPs[i * 32 + lane] = b >> 2;
            bMs[i * 32 + lane] = b << 2;
            const uint32_t lo = __shfl_sync(FULL, br, 0), hi = __shfl_sync(FULL, br, 31);
            if ((lane >> 3) == i) { my_blo = lo; my_bhi = hi; }
        }
        __syncwarp();
        const bool mine = (32 * (lane >> 3) < ncols) && pass_possible(T_, P_, my_alo, my_ahi, my_blo, my_bhi);
        const unsigned cm = __ballot_sync(FULL, mine);             // bit 8c + r: slab r x quarter c can hold a candidate
        n_combo += __popc(cm);
#pragma unroll 1
        for (int q4 = 0; q4 < SSB / 32; q4++) {
            const unsigned sm = (cm >> (8 * q4)) & 0xFFu;
            if (sm == 0) continue;
            const int cend = min(ncols, 32 * q4 + 32);
#pragma unroll 1
            for (int cb = 32 * q4; cb < cend; cb += 4) {
                const uint4 B0 = *reinterpret_cast<const uint4*>(&b0s[cb]);
                const uint4 BP = *reinterpret_cast<const uint4*>(&bPs[cb]);
                const uint4 BM = *reinterpret_cast<const uint4*>(&bMs[cb]);
                uint32_t h = 0;
#pragma unroll
                for (int r = 0; r < RA; r++) {
                    if (sm & (1u << r)) {
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) {
                            uint32_t u = quick_marks(a[r], pick4(B0, kk), pick4(BP, kk), pick4(BM, kk));
                            u &= u * w.one + mone;               // drop the lowest mark (IMAD keeps the -1 off the ALU pipe)
                            if (T_ == 2) u &= u * w.one + mone;
                            h |= (u == 0 ? 1u : 0u) << (kk * 8 + r);
                        }
                    }
                }
                if (cb + 4 > ncols) h &= (1u << (8 * (ncols - cb))) - 1u;      // columns past the end of the array
                if (__any_sync(FULL, h != 0)) { n_cand += __popc(h); sparse_push<T_, P_, BIP>(h, row0, sub + cb, qn, q2n, c, out); }
            }
        }
    }
    __syncwarp();
    while (qn > 0) {                               // what is left in the queues: partial batches
        const int take = min(qn, 32);
        qn -= take;
        const uint2 e = lane < take ? c.q[qn + lane] : make_uint2(0u, 0u);
        sparse_process<T_, P_, BIP>(c, out, e, lane < take, q2n);
        __syncwarp();
    }
    if (q2n > 0) sparse_score(c, out, q2n);         // fewer than 32 by construction
    if (w.stats) {
        for (int o = 16; o; o >>= 1) n_cand += __shfl_down_sync(FULL, n_cand, o);           // per-lane counts
        if (lane == 0) {                                                                  // n_combo is uniform per warp
            atomicAdd(&w.stats[2], n_combo * 1024ull); atomicAdd(&w.stats[3], n_cand); atomicAdd(&w.stats[7], n_score);
            warp_exit_stats(w.stats, t_start);
        }
    }
}

// first index i with v[i] <= v[i-1] (STRICT: the edge construction needs a strictly increasing array) or v[i] < v[i-1] (the
// whitelist only has to be sorted); *first_bad primed with ~0
template <bool STRICT>
__global__ void sorted_check_kernel(const uint32_t* __restrict__ v, uint32_t n, unsigned long long* __restrict__ first_bad)
{
    unsigned long long bad = ~0ull;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t cur = __ldg(&v[i]), prev = __ldg(&v[i - 1]);
        if (STRICT ? cur <= prev : cur < prev) { bad = i; break; }
    }
    if (bad != ~0ull) atomicMin(first_bad, bad);
}

// rotl(key, rot) and the identity payload for a whole array (input of the per-pass radix sort of the bipartite form)
__global__ void rotate_keys_iota_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ pay, uint32_t n, int rot)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { out[i] = rotl32(__ldg(&in[i]), rot); pay[i] = i; }
}

// rotl(key, rot) for a whole array (input of the per-pass radix sort)
__global__ void rotate_keys_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, int rot)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = rotl32(__ldg(&in[i]), rot);
}

}  // namespace bdg
