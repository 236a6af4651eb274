// bdg_api.cu -- C ABI of libbadger_b200.so (see include/badger_b200.h for the contract and the
// reference call sites each entry point replaces).  Host side only: contexts, buffers, work plans,
// launches.  All arithmetic lives in bdg_kernels.cuh / bdg_core.cuh.  There is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/badger_b200.h"
#include "bdg_kernels.cuh"
#include "bdg_join.cuh"
#include "bdg_tsv.hpp"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "%s failed: %s", \
                        #expr, cudaGetErrorString(e_));                                                \
    } while (0)

// grow-only device buffer: the host-buffer entry points reuse their workspaces from call to call
// (cudaMalloc / cudaFree per call cost far more than the kernels at BASELINE config-2 sizes)
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return BDG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return (int)e; }
        cap = want;
        return BDG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// the work plan of one part (built by build_plan below)
struct Plan {
    std::vector<uint32_t> group_ids;
    std::vector<uint32_t> item_start;
    uint32_t chunk_cols = 256;
    uint64_t pairs = 0;
};

struct DevCtx {
    int dev = -1;
    cudaStream_t stream = nullptr;
    int sms = 0;
    Buf sorted, ea, eb, ed, count, plan;   // edge construction
    // sparse passes run side by side on their own streams, so each has its own rotated keys, sorted copy, radix-sort scratch,
    // sub-tile end keys and list of tiles that survive level 1
    Buf rot_in[bdg::MAX_PASSES], rot_sorted[bdg::MAX_PASSES], sort_tmp[bdg::MAX_PASSES], tile_bnd[bdg::MAX_PASSES], tile_list[bdg::MAX_PASSES];
    cudaStream_t aux[bdg::MAX_PASSES] = {nullptr, nullptr, nullptr};   // aux[p]: stream of pass p (pass 0 stays on the caller's stream)
    cudaEvent_t ev_start = nullptr, ev_done[bdg::MAX_PASSES] = {nullptr, nullptr, nullptr};
    unsigned long long host_stats[bdg::MAX_PASSES][2] = {};   // interval tests done / tiles listed per pass of the last launch
    unsigned long long generation = 0;                        // bumped by every host-buffer edge build on this device
    Buf dd[10];                                               // dedup: keys, idx, sorted keys/idx, heads, scan, run arrays, cub scratch
    Buf cl[6];                                                // clustering: centres, centre index, level, claim min / max, flags
    unsigned long long cl_token = 0;                          // read-map token the clustering result resident in cl[1] / cl[2] belongs to (0: none)
    Buf ce[9];                                                // centre selection: positions above the cutoff, sort keys / values, results, scratch
    Buf as[8];                                                // masked dedup + per-read gather: all ranks, valid bytes, their scan, scratch, centre idx / value, result, counter
    unsigned long long map_token = 0, map_serial = 0;         // read map left on the device by bdg_dedup_reads (0: none)
    size_t map_rows = 0, map_reads = 0, map_distinct = 0;
    bool map_masked = false;
    Buf gather_a, gather_b;                                   // node indices of the edge ends: this device's own, and (first device) those of ALL devices, gathered for cluster()
    Buf top16;                                                // first node of every high half of the barcode (the node of an edge end is searched in that run)
    Buf io[4];                                                // host-buffer calls of pack16 / membership: inputs, outputs, check flag
    Buf nn[9];                                                // sparse nearest: rotated keys + payload (in/out) of queries and targets, scratch
    // join form of the t = 2 edge construction (bdg_join.cuh): barcodes in the key order of every condition's row side / column
    // side, first column of every key, scratch keys (in / sorted) and radix-sort scratch, units per slab and their prefix sums
    // (one set per stream: the conditions of a block set run back to back on one stream)
    Buf jn_rows[2], jn_cols[2], jn_tab[2], jn_tabc[2], jn_hist[2], jn_cub[2], jn_counts[2], jn_offs[2], jn_rank[2], jn_lut, jn_weigh;
    int scheme_serial = 0;                                    // which scheme sits in this device's constant memory (0: none)
    cudaEvent_t ev_cond[bdg::SEED_MAX_CONDS] = {};            // bdg_edges_build_into: "condition c has appended its edges"
    cudaEvent_t ev_pre[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};   // streaming form: scratch set k is bucketed / its join has finished
    unsigned long long* snap_host = nullptr;                  // mapped page-locked: the edge count after every condition
    unsigned long long* snap_dev = nullptr;
};
std::vector<DevCtx> g_ctx;

int owner_of_tile(uint64_t I, int P)
{
    const uint64_t m = I % (2ull * P);
    return (int)(m < (uint64_t)P ? m : 2ull * P - 1 - m);
}

int grid_for(const void* kernel, int* out, int threads = bdg::NT)
{
    int dev = 0, sms = 0, occ = 0;
    CU_TRY(cudaGetDevice(&dev));
    CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0));
    if (occ < 1) occ = 1;
    *out = sms * occ;
    return BDG_OK;
}

// ---- the work plan of one part: owned row groups (256 rows, one warp each), column chunks per group ----
// Ownership is dealt in tiles of BDG_ROW_TILE rows (8 groups), boustrophedon over the parts.
void build_plan(size_t N, int part, int nparts, int workers, Plan& p, uint64_t col_quantum = bdg::SB_MAX)
{
    const uint64_t tiles = (N + bdg::ROW_TILE - 1) / bdg::ROW_TILE;
    const uint64_t groups = (N + bdg::GROUP - 1) / bdg::GROUP;
    constexpr uint64_t GPT = bdg::ROW_TILE / bdg::GROUP;
    p.group_ids.clear();
    p.pairs = 0;
    for (uint64_t I = 0; I < tiles; I++) {
        if (owner_of_tile(I, nparts) != part) continue;
        const uint64_t r0 = I * bdg::ROW_TILE, r1 = std::min<uint64_t>(N, r0 + bdg::ROW_TILE);
        const uint64_t n = r1 - r0;                       // rows i in [r0,r1): N-1-i partners each
        p.pairs += n * (N - 1) - (r0 + r1 - 1) * n / 2;
        for (uint64_t g = I * GPT; g < std::min(groups, (I + 1) * GPT); g++) p.group_ids.push_back((uint32_t)g);
    }
    // aim at >= 16 work items per resident warp so that the dynamic scheduler can level the load
    double per_worker = 16.0;
    if (const char* e = getenv("BDG_EDGE_ITEMS")) per_worker = std::max(1.0, atof(e));
    const double want_items = per_worker * std::max(workers, 1);
    const double cols = ((double)p.pairs / bdg::GROUP) / want_items;
    uint64_t cc = (uint64_t)cols / col_quantum * col_quantum;
    cc = std::min<uint64_t>(std::max<uint64_t>(cc, std::max<uint64_t>(col_quantum, 4 * bdg::SB_MAX)), 1u << 17);
    p.chunk_cols = (uint32_t)cc;
    p.item_start.assign(p.group_ids.size() + 1, 0);
    uint64_t acc = 0;
    for (size_t k = 0; k < p.group_ids.size(); k++) {
        p.item_start[k] = (uint32_t)acc;
        const uint64_t ncols = N - (uint64_t)p.group_ids[k] * bdg::GROUP;
        acc += (ncols + cc - 1) / cc;
    }
    p.item_start[p.group_ids.size()] = (uint32_t)acc;
}

constexpr size_t PLAN_HDR = 128;  // per launch: [work cursor u64 | tile-list count u64 | 8 x u64 statistics | pad]
constexpr size_t HDR_LIST = 8, HDR_STATS = 16;
int g_edge_mode = -1;             // -1: BDG_EDGE_MODE or default (by threshold and size); 0 dense; 1 sparse; 2 join

// 0 dense, 1 sparse, 2 join.  The join form exists for t = 2 only and pays off once the key buckets fill.
int edge_mode_for(int t, size_t N)
{
    if (t != 1 && t != 2) return 0;
    int m = g_edge_mode;
    if (m < 0) {
        const char* e = getenv("BDG_EDGE_MODE");
        if (e && !strcmp(e, "dense")) m = 0;
        else if (e && !strcmp(e, "sparse")) m = 1;
        else if (e && !strcmp(e, "join")) m = 2;
        else {
            size_t min_n = 150000;
            if (const char* j = getenv("BDG_JOIN_MIN_N")) min_n = (size_t)std::max(0ll, atoll(j));
            m = (t == 2 && N >= min_n) ? 2 : 1;
        }
    }
    if (m == 2 && (t != 2 || N >= (1ull << 28))) m = 1;
    return m;
}

// ---- the seed scheme of the join form: built once on the host (bdg_seed.cuh), copied to every device that uses it ----
bdg::SeedScheme g_scheme;
std::vector<uint8_t> g_scheme_lut;
int g_scheme_serial = 0;

std::mutex g_scheme_mutex;        // the devices of one build call arrive on their own host threads

int scheme_ready()
{
    std::lock_guard<std::mutex> lock(g_scheme_mutex);
    if (g_scheme_serial) return BDG_OK;
    int bases[bdg::SEED_MAX_BLOCKS] = {3, 3, 3, 3, 3};
    int nb = 5;
    if (const char* e = getenv("BDG_JOIN_BLOCKS")) {          // e.g. "4,4,4,3": development aid
        nb = 0;
        for (const char* p = e; *p && nb < bdg::SEED_MAX_BLOCKS;) {
            bases[nb++] = atoi(p);
            while (*p && *p != ',') p++;
            if (*p == ',') p++;
        }
    }
    if (!bdg::seed_scheme_build(g_scheme, bases, nb)) return fail(BDG_ERR_ARG, "BDG_JOIN_BLOCKS is not a usable block layout (3..6 blocks, 15 bases in all)");
    g_scheme_lut.assign((size_t)1 << g_scheme.nflags, 0);
    bdg::seed_lut_build(g_scheme, g_scheme_lut.data());
    g_scheme_serial = 1;
    return BDG_OK;
}

template <int T_, int P_>
void launch_scan(int blocks, cudaStream_t st, const bdg::EdgeWork& w, const uint2* bnd, uint32_t NS, const bdg::TileList& l)
{
    bdg::sparse_scan_kernel<T_, P_, false><<<blocks, 256, 0, st>>>(w.sorted, w.N, w.group_ids, w.K, bnd, NS, l);
}
template <int T_, int P_>
void launch_tiles(int blocks, cudaStream_t st, const bdg::EdgeWork& w, const bdg::EdgeOut& o, const bdg::TileList& l)
{
    bdg::sparse_tile_kernel<T_, P_, false><<<blocks, bdg::ENT, 0, st>>>(w, o, l);
}
template <int T_, int P_>
void launch_scan_bip(int blocks, cudaStream_t st, const bdg::EdgeWork& w, const uint2* bnd, uint32_t NS, const bdg::TileList& l)
{
    bdg::sparse_scan_kernel<T_, P_, true><<<blocks, 256, 0, st>>>(w.sorted, w.N, nullptr, w.K, bnd, NS, l);
}
template <int T_, int P_>
void launch_tiles_bip(int blocks, cudaStream_t st, const bdg::EdgeWork& w, const bdg::EdgeOut& o, const bdg::TileList& l)
{
    bdg::sparse_tile_kernel<T_, P_, true><<<blocks, bdg::ENT, 0, st>>>(w, o, l);
}
#define BDG_PASS_DISPATCH(FN, ...)                                                    \
    do {                                                                              \
        if (t == 1) { if (p == 0) FN<1, 0>(__VA_ARGS__); else FN<1, 1>(__VA_ARGS__); } \
        else if (p == 0) FN<2, 0>(__VA_ARGS__);                                       \
        else if (p == 1) FN<2, 1>(__VA_ARGS__);                                       \
        else FN<2, 2>(__VA_ARGS__);                                                   \
    } while (0)

double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

// Join form (t = 2), bdg_join.cuh.  The seed conditions are laid on a line by weight (their estimated work, see below) and the
// line is cut into nparts equal pieces: a part sorts and
// joins only the conditions its piece touches, a condition on a cut is shared by row range (cut at a bucket boundary).  Per condition: counting sort by the
// key (rows once per block set, columns for the shifted conditions; its prefix sums are colstart) -> units per slab -> prefix sums ->
// one persistent join launch.  The conditions of a block set follow one another and share the row order.  Consecutive block sets alternate between the caller's stream and an auxiliary one, so that one
// condition's sorts and the tail of its join overlap the neighbour's join.
// Host buffers the finished edges are copied into WHILE later conditions still run (bdg_edges_build_into): the conditions
// then run on one stream, a snapshot of the edge count follows every join launch, and the ranges between two snapshots -
// final, because their launches have ended - leave on the copy stream.
struct JoinStream {
    uint32_t* h_a;
    uint32_t* h_b;
    uint8_t* h_d;
    size_t h_cap;
};

int launch_edges_join(const uint32_t* d_sorted, size_t N, int part, int nparts, uint32_t* d_a, uint32_t* d_b, uint8_t* d_d, size_t cap,
                      unsigned long long* d_count, cudaStream_t caller, DevCtx* ws, const JoinStream* js = nullptr)
{
    if (int rc = scheme_ready()) return rc;
    const bdg::SeedScheme& S = g_scheme;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    if (ws->scheme_serial != g_scheme_serial) {
        if (int e = ensure(ws->jn_lut, g_scheme_lut.size())) return e;
        CU_TRY(cudaMemcpyToSymbolAsync(bdg::c_scheme, &g_scheme, sizeof(g_scheme), 0, cudaMemcpyHostToDevice, caller));
        CU_TRY(cudaMemcpyAsync(ws->jn_lut.p, g_scheme_lut.data(), g_scheme_lut.size(), cudaMemcpyHostToDevice, caller));
        CU_TRY(cudaStreamSynchronize(caller));              // the host copies above are read asynchronously
        ws->scheme_serial = g_scheme_serial;
    }
    // rows per sub-slab: 8 when the key buckets are short (fewer rows then span fewer foreign buckets), else 32
    int rs = (N >> S.ka[0].key_bits) < 48 ? 8 : 32;
    if (const char* e = getenv("BDG_JOIN_RS")) rs = atoi(e) == 8 ? 8 : 32;
    const uint32_t n_slabs = (uint32_t)((N + bdg::JROWS - 1) / bdg::JROWS);
    // Weights of the conditions for the deal to the parts: one part takes everything; several parts estimate every condition's
    // pairs from bucket sizes over a sample (identical integer arithmetic on every part) and add the bucketing of its sides.
    long long weight[bdg::SEED_MAX_CONDS];
    for (int c = 0; c < S.nconds; c++) weight[c] = 1;
    const double t_weigh = getenv("BDG_TRACE") ? (cudaStreamSynchronize(caller), now_ms()) : 0.0;
    if (nparts > 1 || getenv("BDG_TRACE")) {                 // (traced runs print the estimate beside the measured time of every condition)
        uint32_t maxtab = 4;
        for (int c = 0; c < S.nconds; c++) maxtab = std::max(maxtab, 1u << S.ka[c].key_bits);
        bdg::WeighSlots slots{};
        for (int c = 0; c < S.nconds; c++) {
            slots.row[c] = S.cond[c].row_sort == c ? (uint8_t)slots.n++ : slots.row[S.cond[c].row_sort];
            slots.col[c] = S.cond[c].self ? slots.row[c] : (uint8_t)slots.n++;
        }
        const uint32_t stride = (uint32_t)std::max<size_t>(1, (N + (1u << 17) - 1) >> 17);
        const size_t hist_bytes = (size_t)slots.n * maxtab * 4;
        if (int e = ensure(ws->jn_weigh, hist_bytes + 8 * bdg::SEED_MAX_CONDS)) return e;
        uint32_t* d_hist = (uint32_t*)ws->jn_weigh.p;
        unsigned long long* d_pairs = (unsigned long long*)((char*)ws->jn_weigh.p + hist_bytes);
        CU_TRY(cudaMemsetAsync(ws->jn_weigh.p, 0, hist_bytes + 8 * bdg::SEED_MAX_CONDS, caller));
        const uint32_t m = (uint32_t)((N + stride - 1) / stride);
        bdg::join_weigh_hist_kernel<<<(int)std::min<uint32_t>((m + 255) / 256, (uint32_t)ws->sms * 8), 256, 0, caller>>>(d_sorted, (uint32_t)N, stride, S.nconds, maxtab, slots, d_hist);
        bdg::join_weigh_sum_kernel<<<dim3((maxtab / 4 + 1023) / 1024, S.nconds), 256, 0, caller>>>(d_hist, maxtab, slots, d_pairs);
        g_launches += 2;
        unsigned long long pairs[bdg::SEED_MAX_CONDS];
        CU_TRY(cudaMemcpyAsync(pairs, d_pairs, 8 * S.nconds, cudaMemcpyDeviceToHost, caller));
        CU_TRY(cudaStreamSynchronize(caller));
        // Cost model fitted to the per-condition times of a traced C4 run (BDG_TRACE; 25 conditions, rms error 8 %):
        //   time ~ pairs x (1.45 for a symmetric condition, 1 for a shifted one) + 2.4 x N.
        // A symmetric condition's pairs are counted once (n (n - 1) / 2) and hold most of the really close pairs of barcode data -
        // the ones that cost an exact distance, a hand-over check and a score sit on diagonal 0; the N term is the bucketing and
        // the per-slab bookkeeping.  Scaled to <= 4096 so that the cut arithmetic stays small.
        unsigned long long cost[bdg::SEED_MAX_CONDS], top = 1;
        for (int c = 0; c < S.nconds; c++) {
            cost[c] = pairs[c] * stride * stride * (S.cond[c].self ? 29 : 20) + 48ull * N;
            top = std::max(top, cost[c]);
        }
        for (int c = 0; c < S.nconds; c++) weight[c] = (long long)std::max<unsigned long long>(1, cost[c] * 4096 / top);
        if (getenv("BDG_TRACE")) {
            fprintf(stderr, "[bdg] join work estimate: %.3f ms\n", now_ms() - t_weigh);
            fprintf(stderr, "[bdg] join weights (part %d of %d):", part, nparts);
            for (int c = 0; c < S.nconds; c++) fprintf(stderr, " %lld", weight[c]);
            fprintf(stderr, "\n[bdg] join estimated pairs:");
            for (int c = 0; c < S.nconds; c++) fprintf(stderr, " %llu", pairs[c] * stride * stride);
            fprintf(stderr, "\n");
        }
    }
    // Order of the conditions along the line: block set by block set (a set shares its row order), the sets taken alternately
    // from the front and from the back of the table.  A pair that meets several conditions is emitted by the first one in TABLE
    // order, so the early sets own most of the edges; alternating spreads the edges (their scores, stores and PCIe bytes) over
    // the parts.
    int order[bdg::SEED_MAX_CONDS], n_order = 0;
    {
        int set_first[bdg::SEED_MAX_CONDS], set_last[bdg::SEED_MAX_CONDS], n_sets = 0;
        for (int c = 0; c < S.nconds; c++) {
            if (S.cond[c].row_sort == c) { set_first[n_sets] = c; set_last[n_sets] = c; n_sets++; }
            else set_last[n_sets - 1] = c;
        }
        for (int i = 0, lo = 0, hi = n_sets - 1; lo <= hi; i++) {
            const int k = (i & 1) ? hi-- : lo++;
            for (int c = set_first[k]; c <= set_last[k]; c++) order[n_order++] = c;
        }
    }
    long long W = 0;
    for (int c = 0; c < S.nconds; c++) W += weight[c];
    const long long piece_lo = (long long)part * W, piece_hi = (long long)(part + 1) * W;      // cond c covers [start_c * nparts, (start_c + w_c) * nparts)
    if (int e = ensure(ws->plan, PLAN_HDR * bdg::MAX_PASSES + 8 * bdg::SEED_MAX_CONDS)) return e;
    char* d_plan = (char*)ws->plan.p;
    CU_TRY(cudaMemsetAsync(d_plan, 0, PLAN_HDR * bdg::MAX_PASSES + 8 * bdg::SEED_MAX_CONDS, caller));
    for (int p = 0; p < bdg::MAX_PASSES; p++) ws->host_stats[p][0] = ws->host_stats[p][1] = 0;
    unsigned long long* d_cursors = (unsigned long long*)(d_plan + PLAN_HDR * bdg::MAX_PASSES);
    int max_bits = 0;
    for (int c = 0; c < S.nconds; c++) max_bits = std::max(max_bits, (int)S.ka[c].key_bits);
    const size_t max_keys = (size_t)1 << max_bits;
    size_t tmp_a = 0, tmp_b = 0;
    CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_a, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(max_keys + 1), caller));
    CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n_slabs + 1), caller));
    const size_t tmp_bytes = std::max(tmp_a, tmp_b);
    bool ranked = true;                                     // counting sort with one atomic per barcode (its place in the bucket is kept)
    if (const char* e = getenv("BDG_JOIN_RANK")) ranked = atoi(e) != 0;
    for (int k = 0; k < 2; k++) {                           // one scratch set per stream
        if (int e = ensure(ws->jn_hist[k], (max_keys + 1) * 4 * 2)) return e;       // bucket sizes | fill cursors
        if (int e = ensure(ws->jn_tab[k], (max_keys + 1) * 4)) return e;
        if (int e = ensure(ws->jn_tabc[k], (max_keys + 1) * 4)) return e;
        if (int e = ensure(ws->jn_rows[k], N * 4)) return e;
        if (int e = ensure(ws->jn_cols[k], N * 4)) return e;
        if (ranked) { if (int e = ensure(ws->jn_rank[k], N * 4)) return e; }
        if (int e = ensure(ws->jn_cub[k], tmp_bytes)) return e;
        if (int e = ensure(ws->jn_counts[k], ((size_t)n_slabs + 1) * 4)) return e;
        if (int e = ensure(ws->jn_offs[k], ((size_t)n_slabs + 1) * 4)) return e;
    }
    int occ = 6;                                            // resident CTAs per SM the kernel is compiled for (registers per thread follow)
    if (const char* e = getenv("BDG_JOIN_OCC")) occ = atoi(e) == 4 ? 4 : (atoi(e) == 5 ? 5 : 6);
    const void* kern = rs == 8 ? (occ == 4 ? (const void*)bdg::join_kernel<8, 4> : occ == 5 ? (const void*)bdg::join_kernel<8, 5> : (const void*)bdg::join_kernel<8, 6>)
                               : (occ == 4 ? (const void*)bdg::join_kernel<32, 4> : occ == 5 ? (const void*)bdg::join_kernel<32, 5> : (const void*)bdg::join_kernel<32, 6>);
    int grid = 0;
    if (int rc = grid_for(kern, &grid, bdg::ENT)) return rc;
    if (const char* e = getenv("BDG_JOIN_CTAS")) grid = std::min(grid, ws->sms * std::max(1, atoi(e)));   // CTAs per SM of a join launch
    const int gb = (int)std::min<size_t>((N + 255) / 256, (size_t)ws->sms * 8);
    const int bb = (int)std::min<size_t>(((size_t)n_slabs + 256) / 256, (size_t)ws->sms * 8);
    // Two ways to keep the bucketing of one condition out of the way of the joins.  The resident form forks: its two streams
    // carry whole conditions (bucketing + join) and overlap each other.  The streaming form (js) must be able to say "edges
    // [0, n) are complete" after every condition, so its joins stay on ONE stream in order, and only the bucketing of the next
    // condition runs ahead on a second stream, into the other scratch set.
    const bool fork = !js && !getenv("BDG_EDGE_SERIAL");
    const bool ahead = js && !getenv("BDG_EDGE_SERIAL");
    if (fork || ahead) CU_TRY(cudaEventRecord(ws->ev_start, caller));
    bool set_joined[2] = {false, false};                     // ahead: a join on scratch set k is in flight / done (ev_join[k] recorded)
    bool aux_used = false;
    int streamed[bdg::SEED_MAX_CONDS], n_streamed = 0;
    if (js && !ws->snap_host) {
        CU_TRY(cudaHostAlloc((void**)&ws->snap_host, sizeof(unsigned long long) * bdg::SEED_MAX_CONDS, cudaHostAllocMapped));
        CU_TRY(cudaHostGetDevicePointer((void**)&ws->snap_dev, ws->snap_host, 0));
        for (int c = 0; c < bdg::SEED_MAX_CONDS; c++) CU_TRY(cudaEventCreateWithFlags(&ws->ev_cond[c], cudaEventDisableTiming));
        for (int q = 0; q < 2; q++) { CU_TRY(cudaEventCreateWithFlags(&ws->ev_pre[q], cudaEventDisableTiming)); CU_TRY(cudaEventCreateWithFlags(&ws->ev_join[q], cudaEventDisableTiming)); }
    }
    bdg::EdgeOut o{d_a, d_b, d_d, d_count, (unsigned long long)cap};
    long long start = 0;
    int cur_set = -1, k = 1;                                 // k: scratch set / stream of the current condition
    int row_set[2] = {-1, -1};                               // block set whose row order the scratch set holds
    const bool trace = getenv("BDG_TRACE") != nullptr;
    double t_prev = 0;
    if (trace) { cudaStreamSynchronize(caller); t_prev = now_ms(); }
    for (int oi = 0; oi < n_order; oi++) {
        const int c = order[oi];
        const long long w = weight[c];
        const long long c_lo = start * nparts, c_hi = (start + w) * nparts;
        start += w;
        const long long lo = std::max(c_lo, piece_lo), hi = std::min(c_hi, piece_hi);
        if (lo >= hi) continue;
        // The whole job alternates its two streams block set by block set (a set sorts its rows once).  A part of several has few
        // conditions, often of one set: it alternates condition by condition, so that the bucketing of one condition runs beside
        // the join of the other, and sorts the rows again when the stream's row order is of another set.
        const int set = S.cond[c].row_sort;
        if (nparts > 1 || ahead || set != cur_set) k ^= 1;
        cur_set = set;
        const bool new_set = row_set[k] != set;
        row_set[k] = set;
        cudaStream_t st = (fork && k == 1) ? ws->aux[1] : caller;           // the join's stream
        cudaStream_t sp = ahead ? ws->aux[1] : st;                           // the bucketing's stream
        if (fork && k == 1 && !aux_used) { CU_TRY(cudaStreamWaitEvent(st, ws->ev_start, 0)); aux_used = true; }
        if (ahead) {
            if (!aux_used) { CU_TRY(cudaStreamWaitEvent(sp, ws->ev_start, 0)); aux_used = true; }
            if (set_joined[k]) CU_TRY(cudaStreamWaitEvent(sp, ws->ev_join[k], 0));      // the last join that read this scratch set
        }
        // counting sort of the barcodes by one side's key: bucket sizes -> first position of every key (kept: colstart) -> scatter
        auto bucket_side = [&](const bdg::SeedKey& key, Buf& dst, Buf& table) -> int {
            const size_t nkeys = (size_t)1 << key.key_bits;
            uint32_t* hist = (uint32_t*)ws->jn_hist[k].p;
            uint32_t* fill = hist + (nkeys + 1);
            size_t bytes = ws->jn_cub[k].cap;
            if (ranked) {
                uint32_t* rank = (uint32_t*)ws->jn_rank[k].p;
                CU_TRY(cudaMemsetAsync(hist, 0, (nkeys + 1) * 4, sp));
                bdg::join_hist_rank_kernel<<<gb, 256, 0, sp>>>(d_sorted, (uint32_t)N, key, hist, rank);
                CU_TRY(cub::DeviceScan::ExclusiveSum(ws->jn_cub[k].p, bytes, (const uint32_t*)hist, (uint32_t*)table.p, (int)(nkeys + 1), sp));
                bdg::join_scatter_rank_kernel<<<gb, 256, 0, sp>>>(d_sorted, (uint32_t)N, key, (const uint32_t*)table.p, rank, (uint32_t*)dst.p);
            } else {
                CU_TRY(cudaMemsetAsync(hist, 0, (nkeys + 1) * 4 * 2, sp));
                bdg::join_hist_kernel<<<gb, 256, 0, sp>>>(d_sorted, (uint32_t)N, key, hist);
                CU_TRY(cub::DeviceScan::ExclusiveSum(ws->jn_cub[k].p, bytes, (const uint32_t*)hist, (uint32_t*)table.p, (int)(nkeys + 1), sp));
                bdg::join_scatter_kernel<<<gb, 256, 0, sp>>>(d_sorted, (uint32_t)N, key, (const uint32_t*)table.p, fill, (uint32_t*)dst.p);
            }
            g_launches += 2;
            return BDG_OK;
        };
        if (new_set) { if (int e = bucket_side(S.ka[c], ws->jn_rows[k], ws->jn_tab[k])) return e; }      // the stream's row order: this block set
        bdg::JoinArgs A{};
        A.rows = (const uint32_t*)ws->jn_rows[k].p;
        A.rowstart = (const uint32_t*)ws->jn_tab[k].p;
        A.nkeys = 1u << S.ka[c].key_bits;
        if (S.cond[c].self) {
            A.cols = A.rows;
            A.colstart = (const uint32_t*)ws->jn_tab[k].p;
        } else {
            if (int e = bucket_side(S.kb[c], ws->jn_cols[k], ws->jn_tabc[k])) return e;
            A.cols = (const uint32_t*)ws->jn_cols[k].p;
            A.colstart = (const uint32_t*)ws->jn_tabc[k].p;
        }
        A.offs = (const uint32_t*)ws->jn_offs[k].p;
        A.lut = (const uint8_t*)ws->jn_lut.p;
        A.cursor = d_cursors + c;
        A.stats = (unsigned long long*)(d_plan + HDR_STATS);
        A.N = (uint32_t)N;
        A.n_slabs = n_slabs;
        A.cond = c;
        A.self = S.cond[c].self;
        A.f0 = (uint32_t)(lo - c_lo); A.f1 = (uint32_t)(hi - c_lo); A.fden = (uint32_t)(c_hi - c_lo);
        A.T = bdg::qgram_threshold(2);
        A.one = 1u;
        A.mone = 0xFFFFFFFFu;
        bdg::join_band_kernel<<<bb, 256, 0, sp>>>(A, (uint32_t*)ws->jn_counts[k].p);
        size_t bytes = ws->jn_cub[k].cap;
        CU_TRY(cub::DeviceScan::ExclusiveSum(ws->jn_cub[k].p, bytes, (const uint32_t*)ws->jn_counts[k].p, (uint32_t*)ws->jn_offs[k].p, (int)(n_slabs + 1), sp));
        if (ahead) {
            CU_TRY(cudaEventRecord(ws->ev_pre[k], sp));
            CU_TRY(cudaStreamWaitEvent(st, ws->ev_pre[k], 0));
        }
        if (rs == 8) {
            if (occ == 4) bdg::join_kernel<8, 4><<<grid, bdg::ENT, 0, st>>>(A, o);
            else if (occ == 5) bdg::join_kernel<8, 5><<<grid, bdg::ENT, 0, st>>>(A, o);
            else bdg::join_kernel<8, 6><<<grid, bdg::ENT, 0, st>>>(A, o);
        } else {
            if (occ == 4) bdg::join_kernel<32, 4><<<grid, bdg::ENT, 0, st>>>(A, o);
            else if (occ == 5) bdg::join_kernel<32, 5><<<grid, bdg::ENT, 0, st>>>(A, o);
            else bdg::join_kernel<32, 6><<<grid, bdg::ENT, 0, st>>>(A, o);
        }
        g_launches += 3;
        CU_TRY(cudaGetLastError());
        if (ahead) { CU_TRY(cudaEventRecord(ws->ev_join[k], st)); set_joined[k] = true; }
        if (trace) {                                        // development aid: per-condition wall time (serialises the streams)
            CU_TRY(cudaStreamSynchronize(st));
            const double now = now_ms();
            fprintf(stderr, "[bdg] join cond %2d (%s, rows %u/%u..%u/%u, weight %lld): %.3f ms\n", c, S.cond[c].self ? "sym" : "shf", A.f0, A.fden, A.f1, A.fden,
                    weight[c], now - t_prev);
            t_prev = now;
        }
        if (js) {
            bdg::join_snapshot_kernel<<<1, 1, 0, st>>>(d_count, ws->snap_dev + c);
            g_launches++;
            CU_TRY(cudaEventRecord(ws->ev_cond[c], st));
            streamed[n_streamed++] = c;
        }
    }
    if (aux_used) {
        CU_TRY(cudaEventRecord(ws->ev_done[1], ws->aux[1]));
        CU_TRY(cudaStreamWaitEvent(caller, ws->ev_done[1], 0));
    }
    if (trace) { cudaStreamSynchronize(caller); fprintf(stderr, "[bdg] join part %d of %d: %.3f ms from entry to the last condition\n", part, nparts, now_ms() - t_weigh); }
    if (js) {                                               // everything is enqueued: follow the conditions and copy what they have finished
        cudaStream_t cp = ws->aux[2];
        size_t done = 0;
        for (int k = 0; k < n_streamed; k++) {
            CU_TRY(cudaEventSynchronize(ws->ev_cond[streamed[k]]));
            const size_t upto = (size_t)std::min<unsigned long long>(ws->snap_host[streamed[k]], js->h_cap);
            if (upto > done) {
                CU_TRY(cudaMemcpyAsync(js->h_a + done, d_a + done, (upto - done) * 4, cudaMemcpyDeviceToHost, cp));
                CU_TRY(cudaMemcpyAsync(js->h_b + done, d_b + done, (upto - done) * 4, cudaMemcpyDeviceToHost, cp));
                CU_TRY(cudaMemcpyAsync(js->h_d + done, d_d + done, upto - done, cudaMemcpyDeviceToHost, cp));
                done = upto;
            }
        }
        CU_TRY(cudaStreamSynchronize(cp));
    }
    return BDG_OK;
}

// Launch the edge construction of one part on the current device / stream.  d_count is zeroed on the stream.
// ws: the device's grow-only workspaces (plan, rotated keys, sort scratch); one in-flight call per device.
int launch_edges(const uint32_t* d_sorted, size_t N, int t, int part, int nparts, uint32_t* d_a, uint32_t* d_b,
                 uint8_t* d_d, size_t cap, unsigned long long* d_count, cudaStream_t st, DevCtx* ws, const JoinStream* js = nullptr)
{
    if (nparts < 1 || part < 0 || part >= nparts) return fail(BDG_ERR_ARG, "part %d of %d is not a valid part", part, nparts);
    if (N > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "N = %zu exceeds the 2^32 distinct 16-mers", N);
    if (!ws) return fail(BDG_ERR_NODEVICE, "bdg_init has not claimed the current CUDA device");
    CU_TRY(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
    if (t <= 0 || N < 2) return BDG_OK;   // D >= 1 for distinct barcodes: no edges (barcode_graph.py:245)
    const int mode = edge_mode_for(t, N);
    if (mode == 2) return launch_edges_join(d_sorted, N, part, nparts, d_a, d_b, d_d, cap, d_count, st, ws, js);
    const bool sparse = mode == 1;
    const int passes = sparse ? bdg::n_passes(t) : 1;
    const void* kern = sparse ? (t == 1 ? (const void*)bdg::sparse_tile_kernel<1, 0, false> : (const void*)bdg::sparse_tile_kernel<2, 0, false>)
                              : (t == 1 ? (const void*)bdg::edges_kernel<1> : t == 2 ? (const void*)bdg::edges_kernel<2>
                                                                                     : (const void*)bdg::edges_kernel<3>);
    int grid = 0;
    if (int rc = grid_for(kern, &grid, bdg::ENT)) return rc;
    if (const char* e = getenv("BDG_JOIN_CTAS")) grid = std::min(grid, ws->sms * std::max(1, atoi(e)));   // CTAs per SM of a join launch
    Plan plan;
    build_plan(N, part, nparts, grid * bdg::EW, plan, bdg::SB_MAX);
    const uint32_t n_items = plan.item_start.empty() ? 0 : plan.item_start.back();
    if (n_items == 0) return BDG_OK;
    const size_t nb_groups = plan.group_ids.size() * sizeof(uint32_t), nb_items = plan.item_start.size() * sizeof(uint32_t);
    const size_t hdr = PLAN_HDR * bdg::MAX_PASSES;
    if (cudaError_t e = (cudaError_t)ws->plan.ensure(hdr + nb_groups + nb_items))
        return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "plan workspace: %s", cudaGetErrorString(e));
    char* d_plan = (char*)ws->plan.p;   // [headers | group_ids | item_start]
    CU_TRY(cudaMemsetAsync(d_plan, 0, hdr, st));
    for (int p = 0; p < bdg::MAX_PASSES; p++) ws->host_stats[p][0] = ws->host_stats[p][1] = 0;
    CU_TRY(cudaMemcpyAsync(d_plan + hdr, plan.group_ids.data(), nb_groups, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_plan + hdr + nb_groups, plan.item_start.data(), nb_items, cudaMemcpyHostToDevice, st));
    bdg::EdgeWork w;
    w.N = (uint32_t)N;
    w.t = t;
    w.T = bdg::qgram_threshold(t);
    w.group_ids = (const uint32_t*)(d_plan + hdr);
    w.item_start = (const uint32_t*)(d_plan + hdr + nb_groups);
    w.K = (uint32_t)plan.group_ids.size();
    w.n_items = n_items;
    w.chunk_cols = plan.chunk_cols;
    w.one = 1u;
    w.mone = 0xFFFFFFFFu;
    bdg::EdgeOut o{d_a, d_b, d_d, d_count, (unsigned long long)cap};
    const int blocks = (int)std::min<uint64_t>((uint64_t)grid, (n_items + bdg::EW - 1) / bdg::EW);
    // sparse passes are independent (disjoint outputs through one atomic cursor): pass 0 runs on the caller's stream, the others
    // on the device's auxiliary streams, so that one pass's sort / scan and its tail overlap the neighbour's tile kernel
    const bool fork = sparse && passes > 1 && !getenv("BDG_EDGE_SERIAL");
    cudaStream_t caller = st;
    if (fork) CU_TRY(cudaEventRecord(ws->ev_start, caller));
    for (int p = 0; p < passes; p++) {
        st = (fork && p > 0) ? ws->aux[p] : caller;
        if (fork && p > 0) CU_TRY(cudaStreamWaitEvent(st, ws->ev_start, 0));
        w.item_counter = (unsigned int*)(d_plan + PLAN_HDR * p);
        w.stats = (unsigned long long*)(d_plan + PLAN_HDR * p + HDR_STATS);
        w.pass = p;
        w.rot = sparse ? bdg::pass_rot(t, p) : 0;
        w.sorted = d_sorted;
        auto ensure = [&](Buf& b, size_t bytes) -> int {
            if (cudaError_t e = (cudaError_t)b.ensure(bytes))
                return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
            return BDG_OK;
        };
        if (w.rot != 0) {   // this pass scans the array in the order of rotl(key, rot): rotate, radix-sort
            if (int e = ensure(ws->rot_in[p], N * 4)) return e;
            if (int e = ensure(ws->rot_sorted[p], N * 4)) return e;
            size_t tmp_bytes = 0;
            CU_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const uint32_t*)ws->rot_in[p].p, (uint32_t*)ws->rot_sorted[p].p, (int)N, 0, 32, st));
            if (int e = ensure(ws->sort_tmp[p], tmp_bytes)) return e;
            const int rb = (int)std::min<size_t>((N + 255) / 256, (size_t)ws->sms * 8);
            bdg::rotate_keys_kernel<<<rb, 256, 0, st>>>(d_sorted, (uint32_t*)ws->rot_in[p].p, (uint32_t)N, w.rot);
            g_launches++;
            CU_TRY(cub::DeviceRadixSort::SortKeys(ws->sort_tmp[p].p, tmp_bytes, (const uint32_t*)ws->rot_in[p].p, (uint32_t*)ws->rot_sorted[p].p, (int)N, 0, 32, st));
            w.sorted = (const uint32_t*)ws->rot_sorted[p].p;
        }
        if (!sparse) {
            if (t == 1) bdg::edges_kernel<1><<<blocks, bdg::ENT, 0, st>>>(w, o);
            else if (t == 2) bdg::edges_kernel<2><<<blocks, bdg::ENT, 0, st>>>(w, o);
            else bdg::edges_kernel<3><<<blocks, bdg::ENT, 0, st>>>(w, o);
        } else {
            // level 1: interval scan of every (owned row group, sub-tile right of it) -> compact tile list
            const uint32_t NS = (uint32_t)((N + bdg::SSB - 1) / bdg::SSB);
            if (int e = ensure(ws->tile_bnd[p], (size_t)NS * sizeof(uint2))) return e;
            if (ws->tile_list[p].cap == 0) {
                size_t want = std::max<size_t>((size_t)1 << 22, 4 * N);
                if (const char* e = getenv("BDG_TILE_LIST_CAP")) want = (size_t)std::max(1ll, atoll(e));   // tests: force the regrow path
                if (int e = ensure(ws->tile_list[p], want * sizeof(uint2))) return e;
            }
            bdg::tile_bounds_kernel<<<std::min<uint32_t>((NS + 255) / 256, (uint32_t)ws->sms * 8), 256, 0, st>>>(w.sorted, w.N, (uint2*)ws->tile_bnd[p].p, NS);
            g_launches++;
            // no read-back between the scan and the tile kernel: the tile kernel reads the list length on the device and, if
            // the list was too small, poisons the edge count (bit 63), which every caller inspects before using the edges
            bdg::TileList l{(uint2*)ws->tile_list[p].p, (unsigned long long*)(d_plan + PLAN_HDR * p + HDR_LIST), ws->tile_list[p].cap / sizeof(uint2)};
            const int sblocks = (int)std::min<uint64_t>(w.K, (uint64_t)ws->sms * 16);
            BDG_PASS_DISPATCH(launch_scan, sblocks, st, w, (const uint2*)ws->tile_bnd[p].p, NS, l);
            g_launches++;
            CU_TRY(cudaGetLastError());
            uint64_t tests = 0;
            for (uint32_t g : plan.group_ids) tests += NS - std::min<uint64_t>(NS, (uint64_t)g * (bdg::GROUP / bdg::SSB));
            ws->host_stats[p][0] = tests;
            BDG_PASS_DISPATCH(launch_tiles, grid, st, w, o, l);
        }
        g_launches++;
        CU_TRY(cudaGetLastError());
        if (fork && p > 0) CU_TRY(cudaEventRecord(ws->ev_done[p], st));
    }
    if (fork)
        for (int p = 1; p < passes; p++) CU_TRY(cudaStreamWaitEvent(caller, ws->ev_done[p], 0));
    return BDG_OK;
}


// Q x W scoring through the sparse machinery (max_d <= 2): queries and targets are each sorted by the pass's rotated
// key (original indices ride along), tiles of (256 queries x 128 targets) whose key intervals cannot meet are decided
// by the scan, the rest get the quick test, survivors the exact plain distance and an atomicMin on the query's key.
int launch_nearest_sparse(const uint32_t* d_q, size_t Q, const uint32_t* d_t, size_t W, int max_d, uint32_t* d_keys, cudaStream_t st, DevCtx* ws)
{
    const int T_ = max_d <= 1 ? 1 : 2;
    const int passes = bdg::n_passes(T_);
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "workspace of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    // nn[0..3]: query keys in / payload in / keys sorted / payload sorted; nn[4..7]: the same for the targets; nn[8]: sort scratch
    for (int k = 0; k < 4; k++) if (int e = ensure(ws->nn[k], Q * 4)) return e;
    for (int k = 4; k < 8; k++) if (int e = ensure(ws->nn[k], W * 4)) return e;
    uint32_t *qk = (uint32_t*)ws->nn[0].p, *qp = (uint32_t*)ws->nn[1].p, *qks = (uint32_t*)ws->nn[2].p, *qps = (uint32_t*)ws->nn[3].p;
    uint32_t *tk = (uint32_t*)ws->nn[4].p, *tp = (uint32_t*)ws->nn[5].p, *tks = (uint32_t*)ws->nn[6].p, *tps = (uint32_t*)ws->nn[7].p;
    size_t tmp_q = 0, tmp_t = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_q, qk, qks, qp, qps, (int)Q, 0, 32, st));
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_t, tk, tks, tp, tps, (int)W, 0, 32, st));
    if (int e = ensure(ws->nn[8], std::max(tmp_q, tmp_t))) return e;
    const uint32_t NS = (uint32_t)((W + bdg::SSB - 1) / bdg::SSB);
    const uint32_t K = (uint32_t)((Q + bdg::GROUP - 1) / bdg::GROUP);
    if (int e = ensure(ws->tile_bnd[0], (size_t)NS * sizeof(uint2))) return e;
    if (ws->tile_list[0].cap == 0)
        if (int e = ensure(ws->tile_list[0], ((size_t)1 << 22) * sizeof(uint2))) return e;
    if (int e = ensure(ws->plan, PLAN_HDR * bdg::MAX_PASSES)) return e;
    char* d_plan = (char*)ws->plan.p;   // only the headers are used here; a cached edge plan behind them stays valid
    CU_TRY(cudaMemsetAsync(d_plan, 0, PLAN_HDR * bdg::MAX_PASSES, st));
    int grid = 0;
    if (int rc = grid_for(T_ == 1 ? (const void*)bdg::sparse_tile_kernel<1, 0, true> : (const void*)bdg::sparse_tile_kernel<2, 0, true>, &grid, bdg::ENT)) return rc;
    const int t = T_;   // BDG_PASS_DISPATCH reads `t` and `p`
    for (int p = 0; p < passes; p++) {
        const int rot = bdg::pass_rot(T_, p);
        const int qb = (int)std::min<size_t>((Q + 255) / 256, (size_t)ws->sms * 8), tb = (int)std::min<size_t>((W + 255) / 256, (size_t)ws->sms * 8);
        bdg::rotate_keys_iota_kernel<<<qb, 256, 0, st>>>(d_q, qk, qp, (uint32_t)Q, rot);
        bdg::rotate_keys_iota_kernel<<<tb, 256, 0, st>>>(d_t, tk, tp, (uint32_t)W, rot);
        CU_TRY(cub::DeviceRadixSort::SortPairs(ws->nn[8].p, tmp_q, qk, qks, qp, qps, (int)Q, 0, 32, st));
        CU_TRY(cub::DeviceRadixSort::SortPairs(ws->nn[8].p, tmp_t, tk, tks, tp, tps, (int)W, 0, 32, st));
        bdg::tile_bounds_kernel<<<std::min<uint32_t>((NS + 255) / 256, (uint32_t)ws->sms * 8), 256, 0, st>>>(tks, (uint32_t)W, (uint2*)ws->tile_bnd[0].p, NS);
        g_launches += 3;
        bdg::EdgeWork w{};
        w.sorted = qks; w.N = (uint32_t)Q; w.t = max_d; w.T = 0; w.group_ids = nullptr; w.item_start = nullptr; w.K = K;
        w.item_counter = (unsigned int*)(d_plan + PLAN_HDR * p);
        w.stats = nullptr; w.one = 1u; w.mone = 0xFFFFFFFFu; w.pass = p; w.rot = rot;
        w.cols = tks; w.NC = (uint32_t)W; w.row_pay = qps; w.col_pay = tps; w.near_keys = d_keys;
        unsigned long long n_tiles = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            bdg::TileList l{(uint2*)ws->tile_list[0].p, (unsigned long long*)(d_plan + PLAN_HDR * p + HDR_LIST), ws->tile_list[0].cap / sizeof(uint2)};
            CU_TRY(cudaMemsetAsync(l.count, 0, 8, st));
            const int sblocks = (int)std::min<uint64_t>(K, (uint64_t)ws->sms * 16);
            BDG_PASS_DISPATCH(launch_scan_bip, sblocks, st, w, (const uint2*)ws->tile_bnd[0].p, NS, l);
            g_launches++;
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(&n_tiles, l.count, 8, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            if (n_tiles <= l.cap) break;
            if (attempt == 1) return fail(BDG_ERR_CUDA, "tile list count changed between identical scans");
            if (int e = ensure(ws->tile_list[0], (size_t)n_tiles * sizeof(uint2))) return e;
        }
        if (n_tiles == 0) continue;
        bdg::TileList l{(uint2*)ws->tile_list[0].p, (unsigned long long*)(d_plan + PLAN_HDR * p + HDR_LIST), ws->tile_list[0].cap / sizeof(uint2)};
        const int tblocks = (int)std::min<uint64_t>((uint64_t)grid, (n_tiles + bdg::EW - 1) / bdg::EW);
        bdg::EdgeOut o{nullptr, nullptr, nullptr, nullptr, 0};
        BDG_PASS_DISPATCH(launch_tiles_bip, tblocks, st, w, o, l);
        g_launches++;
        CU_TRY(cudaGetLastError());
    }
    return BDG_OK;
}

DevCtx* ctx_of_current_device();

DevCtx* ctx_of_current_device()
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (auto& c : g_ctx) if (c.dev == dev) return &c;
    return nullptr;
}

int need_ctx()
{
    if (g_ctx.empty()) {
        int rc = bdg_init(nullptr, 0);
        if (rc) return rc;
    }
    return BDG_OK;
}

int check_sorted(const uint32_t* v, size_t N)
{
    for (size_t i = 1; i < N; i++)
        if (v[i] <= v[i - 1]) return fail(BDG_ERR_ARG, "input not strictly increasing at index %zu", i);
    return BDG_OK;
}

size_t edge_cap_guess(size_t N, int t, int nparts)
{
    // a too small guess costs a whole second run (the kernel only counts what it cannot store); 9 bytes per slot are cheap
    size_t per_row = t <= 1 ? 16 : 64;
    if (const char* e = getenv("BDG_EDGE_CAP_PER_ROW")) per_row = (size_t)std::max(1ll, atoll(e));
    return std::max<size_t>(1u << 16, per_row * N / (size_t)nparts + 1024);
}

}  // namespace

// Handle of one edge construction: the edges stay in the devices' grow-only workspaces until the caller copies them
// out (bdg_edges_copy: ONE device-to-host copy per array, straight into the caller's buffers).  A later build on the
// same device reuses those workspaces, which makes older handles stale (checked through the generation numbers).
struct bdg_edges {
    std::vector<int> ctx;                 // indices into g_ctx
    std::vector<size_t> count;            // edges held by each of them
    std::vector<unsigned long long> gen;  // workspace generation at build time
};

// Known strings of a KmerIndexer / QGramIndex, resident on one device, plus grow-only query / result buffers.
struct bdg_kmer_index {
    int dev = -1;
    size_t W = 0;
    Buf wl;
    Buf outb[6];
    Buf kstart, post, scratch;        // posting lists: first posting of every 6-mer (4097 words), string ids grouped by 6-mer
    bool posted = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};   // around the last query's kernel
    bool timed = false;
};

extern "C" {

const char* bdg_version(void) { return "badger_b200 0.1 (sm_100a)"; }
const char* bdg_last_error(void) { return g_err; }
int bdg_device_count(void) { return (int)g_ctx.size(); }
unsigned long long bdg_launch_count(void) { return g_launches.load(); }

int bdg_init(const int* device_ids, int n_devices)
{
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible == 0)
        return fail(BDG_ERR_NODEVICE, "no CUDA device: %s (libbadger_b200 has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) for (int i = 0; i < visible; i++) ids.push_back(i);
    else ids.assign(device_ids, device_ids + n_devices);
    bool same = ids.size() == g_ctx.size();
    for (size_t i = 0; same && i < ids.size(); i++) same = g_ctx[i].dev == ids[i];
    if (same) return BDG_OK;
    bdg_shutdown();
    for (int id : ids) {
        if (id < 0 || id >= visible) { bdg_shutdown(); return fail(BDG_ERR_ARG, "device id %d out of range (%d visible)", id, visible); }
        DevCtx c;
        c.dev = id;
        CU_TRY(cudaSetDevice(id));
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, id));
        if (prop.major < 10) { bdg_shutdown(); return fail(BDG_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", id, prop.major, prop.minor); }
        c.sms = prop.multiProcessorCount;
        CU_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CU_TRY(cudaEventCreateWithFlags(&c.ev_start, cudaEventDisableTiming));
        for (int p = 1; p < bdg::MAX_PASSES; p++) {
            CU_TRY(cudaStreamCreateWithFlags(&c.aux[p], cudaStreamNonBlocking));
            CU_TRY(cudaEventCreateWithFlags(&c.ev_done[p], cudaEventDisableTiming));
        }
        cudaMemPool_t pool;   // keep stream-ordered allocations cached instead of returning them at every sync
        if (cudaDeviceGetDefaultMemPool(&pool, id) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        g_ctx.push_back(c);
    }
    // direct NVLink copies into the first device (bdg_cluster_levels_from_edges gathers the other devices' edge lists there);
    // without peer access cudaMemcpyPeerAsync stages through host memory
    // (both directions: the edge lists travel to the first device, the distinct-barcode array fans out from it)
    for (size_t g = 1; g < g_ctx.size(); g++) {
        for (int dir = 0; dir < 2; dir++) {
            const int from = dir ? g_ctx[g].dev : g_ctx[0].dev, to = dir ? g_ctx[0].dev : g_ctx[g].dev;
            int can = 0;
            CU_TRY(cudaSetDevice(from));
            if (cudaDeviceCanAccessPeer(&can, from, to) == cudaSuccess && can) {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(to, 0);
                if (pe != cudaSuccess) (void)cudaGetLastError();      // already enabled (e.g. by the caller's framework): fine
            }
        }
    }
    CU_TRY(cudaSetDevice(g_ctx[0].dev));
    return BDG_OK;
}

void bdg_shutdown(void)
{
    for (auto& c : g_ctx) {
        if (c.stream) { cudaSetDevice(c.dev); cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); }
        c.sorted.release(); c.ea.release(); c.eb.release(); c.ed.release(); c.count.release(); c.plan.release();
        for (int p = 0; p < bdg::MAX_PASSES; p++) {
            c.rot_in[p].release(); c.sort_tmp[p].release(); c.tile_bnd[p].release(); c.tile_list[p].release();
            if (c.aux[p]) cudaStreamDestroy(c.aux[p]);
            if (c.ev_done[p]) cudaEventDestroy(c.ev_done[p]);
        }
        if (c.ev_start) cudaEventDestroy(c.ev_start);
        for (auto& b : c.dd) b.release();
        for (auto& b : c.nn) b.release();
        for (auto& b : c.io) b.release();
        for (auto& b : c.cl) b.release();
        c.gather_a.release(); c.gather_b.release(); c.top16.release();
        for (auto& b : c.as) b.release();
        for (auto& b : c.ce) b.release();
        for (auto& b : c.rot_sorted) b.release();
        for (int k = 0; k < 2; k++) {
            c.jn_rows[k].release(); c.jn_cols[k].release(); c.jn_tab[k].release(); c.jn_tabc[k].release(); c.jn_hist[k].release();
            c.jn_cub[k].release(); c.jn_counts[k].release(); c.jn_offs[k].release();
        }
        c.jn_lut.release(); c.jn_weigh.release(); for (auto& b : c.jn_rank) b.release();
        if (c.snap_host) { cudaFreeHost(c.snap_host); c.snap_host = nullptr; for (auto& e : c.ev_cond) if (e) cudaEventDestroy(e);
                           for (int q = 0; q < 2; q++) { if (c.ev_pre[q]) cudaEventDestroy(c.ev_pre[q]); if (c.ev_join[q]) cudaEventDestroy(c.ev_join[q]); } }
    }
    g_ctx.clear();
}

unsigned long long bdg_part_pairs(size_t N, int part, int nparts)
{
    if (nparts < 1 || part < 0 || part >= nparts) return 0;
    Plan p;
    build_plan(N, part, nparts, 1, p);
    return p.pairs;
}

// ---------------------------------------------------------------- device-resident entry points
int bdg_dev_edges_build(const uint32_t* d_sorted, size_t N, int t, int part, int nparts, uint32_t* d_a, uint32_t* d_b,
                        uint8_t* d_d, size_t cap, unsigned long long* d_count, void* stream)
{
    if (!d_count || (N && !d_sorted) || (cap && (!d_a || !d_b || !d_d))) return fail(BDG_ERR_ARG, "NULL pointer argument");
    DevCtx* c = ctx_of_current_device();
    return launch_edges(d_sorted, N, t, part, nparts, d_a, d_b, d_d, cap, d_count, (cudaStream_t)stream, c);
}

// Page-locked host memory for callers that want device-to-host copies at full PCIe speed (badger_b200.ops keeps a small
// pool of such blocks behind the numpy arrays it returns).
int bdg_host_alloc(size_t bytes, void** out)
{
    if (!out) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    if (int rc = need_ctx()) return rc;
    cudaError_t e = cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable);
    if (e != cudaSuccess) { *out = nullptr; return fail(BDG_ERR_OOM, "cudaHostAlloc of %zu bytes: %s", bytes, cudaGetErrorString(e)); }
    return BDG_OK;
}

void bdg_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int bdg_set_edge_mode(int mode)
{
    if (mode < -1 || mode > 2) return fail(BDG_ERR_ARG, "edge mode must be -1 (default), 0 (dense), 1 (sparse) or 2 (join)");
    g_edge_mode = mode;
    return BDG_OK;
}

int bdg_dev_edges_stats(unsigned long long* out5, void* stream)
{
    DevCtx* c = ctx_of_current_device();
    if (!c || !c->plan.p || !out5) return fail(BDG_ERR_ARG, "no edge launch on this device yet");
    unsigned long long v[(PLAN_HDR / 8) * bdg::MAX_PASSES];
    CU_TRY(cudaMemcpyAsync(v, c->plan.p, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    for (int k = 0; k < 5; k++) {
        out5[k] = 0;
        const int slot = k < 4 ? k : 7;
        for (int p = 0; p < bdg::MAX_PASSES; p++) {
            out5[k] += v[(PLAN_HDR / 8) * p + HDR_STATS / 8 + slot];
            if (k == 0) out5[k] += c->host_stats[p][0];                       // interval tests of the sparse scans (counted on the host)
            if (k == 1) out5[k] += v[(PLAN_HDR / 8) * p + HDR_LIST / 8];      // tiles the scans listed
        }
    }
    return BDG_OK;
}

// The kernels' eight raw counters summed over the passes (join form: [0] work units, [2] pairs tested, [3] candidates,
// [6] candidates with D <= 2, [7] pairs whose 6-mer score was computed; [4] / [5] sum / max of the warps' busy time in ns).
int bdg_dev_edges_stats_raw(unsigned long long* out8, void* stream)
{
    DevCtx* c = ctx_of_current_device();
    if (!c || !c->plan.p || !out8) return fail(BDG_ERR_ARG, "no edge launch on this device yet");
    unsigned long long v[(PLAN_HDR / 8) * bdg::MAX_PASSES];
    CU_TRY(cudaMemcpyAsync(v, c->plan.p, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    for (int k = 0; k < 8; k++) {
        out8[k] = 0;
        for (int p = 0; p < bdg::MAX_PASSES; p++) {
            const unsigned long long x = v[(PLAN_HDR / 8) * p + HDR_STATS / 8 + k];
            out8[k] = k == 5 ? std::max(out8[k], x) : out8[k] + x;
        }
    }
    return BDG_OK;
}

// Development aid (tools/sweep_edges.py): per pass, sum and max over warps of the warp's exit time in ns
// measured from the first warp's start - the load-balance picture of the last edge launch.
int bdg_dev_edges_balance(unsigned long long* out /* 2 * MAX_PASSES */, void* stream)
{
    DevCtx* c = ctx_of_current_device();
    if (!c || !c->plan.p || !out) return fail(BDG_ERR_ARG, "no edge launch on this device yet");
    unsigned long long v[(PLAN_HDR / 8) * bdg::MAX_PASSES];
    CU_TRY(cudaMemcpyAsync(v, c->plan.p, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    for (int p = 0; p < bdg::MAX_PASSES; p++) {
        out[2 * p] = v[(PLAN_HDR / 8) * p + HDR_STATS / 8 + 4];
        out[2 * p + 1] = v[(PLAN_HDR / 8) * p + HDR_STATS / 8 + 5];
    }
    return BDG_OK;
}

int bdg_dev_pack16(const char* d_seqs, size_t R, uint32_t* d_out, uint8_t* d_valid, void* stream)
{
    if (R == 0) return BDG_OK;
    if (!d_seqs || !d_out || !d_valid) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if ((uintptr_t)d_seqs % 16) return fail(BDG_ERR_ARG, "sequence buffer must be 16-byte aligned");
    int grid = 0;
    if (int rc = grid_for((const void*)bdg::pack16_kernel, &grid)) return rc;
    const int blocks = (int)std::min<uint64_t>((uint64_t)grid, (R + 4 * bdg::NT - 1) / (4 * bdg::NT));
    bdg::pack16_kernel<<<blocks, bdg::NT, 0, (cudaStream_t)stream>>>((const uint4*)d_seqs, R, d_out, d_valid);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return BDG_OK;
}

int bdg_dev_member_sorted(const uint32_t* d_wl, size_t W, const uint32_t* d_q, size_t Q, uint8_t* d_hit, void* stream)
{
    if (Q == 0) return BDG_OK;
    if (!d_q || !d_hit || (W && !d_wl)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (W > 0xFFFFFFFFull || Q > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "size exceeds 2^32");
    if (W == 0) { CU_TRY(cudaMemsetAsync(d_hit, 0, Q, (cudaStream_t)stream)); return BDG_OK; }
    int grid = 0;
    if (int rc = grid_for((const void*)bdg::member_kernel, &grid)) return rc;
    const int blocks = (int)std::min<uint64_t>((uint64_t)grid, (Q + bdg::NT - 1) / bdg::NT);
    bdg::member_kernel<<<blocks, bdg::NT, 0, (cudaStream_t)stream>>>(d_wl, (uint32_t)W, d_q, (uint32_t)Q, d_hit);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return BDG_OK;
}

int bdg_dev_nearest_bounded(const uint32_t* d_q, size_t Q, const uint32_t* d_t, size_t W, int max_d, uint32_t* d_keys,
                            int32_t* d_argmin, uint8_t* d_dist, void* stream)
{
    if (Q == 0) return BDG_OK;
    if (!d_q || !d_keys || !d_argmin || !d_dist || (W && !d_t)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (W >= (1ull << bdg::NEAR_IDX_BITS) || Q > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "W must be < 2^28 and Q < 2^32");
    if (max_d > 15) return fail(BDG_ERR_ARG, "max_d must be <= 15 (a distance of 16 does not fit the packed (distance, index) key)");
    cudaStream_t st = (cudaStream_t)stream;
    CU_TRY(cudaMemsetAsync(d_keys, 0xFF, Q * sizeof(uint32_t), st));
    DevCtx* ws = ctx_of_current_device();
    const bool sparse = ws && W > 0 && max_d >= 0 && max_d <= 2 && (unsigned long long)Q * W >= (1ull << 24) && !getenv("BDG_NEAREST_DENSE");
    if (sparse) {
        if (int rc = launch_nearest_sparse(d_q, Q, d_t, W, max_d, d_keys, st, ws)) return rc;
    } else if (W > 0 && max_d >= 0) {
        int sms = 0, dev = 0;
        CU_TRY(cudaGetDevice(&dev));
        CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const uint32_t gx = (uint32_t)((Q + bdg::ROW_TILE - 1) / bdg::ROW_TILE);
        uint32_t gy = std::max<uint32_t>(1, (uint32_t)(4 * sms) / gx);          // split targets when there are few query tiles
        gy = std::min<uint32_t>(gy, (uint32_t)((W + bdg::NEAR_TB - 1) / bdg::NEAR_TB));
        gy = std::min<uint32_t>(gy, 65535u);
        uint32_t wpb = (uint32_t)((W + gy - 1) / gy);
        wpb = (wpb + bdg::NEAR_TB - 1) / bdg::NEAR_TB * bdg::NEAR_TB;
        gy = (uint32_t)((W + wpb - 1) / wpb);
        dim3 grid(gx, gy);
        if (max_d <= 2) bdg::nearest_kernel<true><<<grid, bdg::NT, 0, st>>>(d_q, (uint32_t)Q, d_t, (uint32_t)W, max_d, wpb, d_keys);
        else bdg::nearest_kernel<false><<<grid, bdg::NT, 0, st>>>(d_q, (uint32_t)Q, d_t, (uint32_t)W, max_d, wpb, d_keys);
        g_launches++;
        CU_TRY(cudaGetLastError());
    }
    bdg::nearest_finish_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, st>>>(d_keys, (uint32_t)Q, d_argmin, d_dist);
    g_launches++;
    CU_TRY(cudaGetLastError());
    return BDG_OK;
}

int bdg_dev_pipe_probe(int kind, int blocks, int iters, uint32_t* d_sink, unsigned long long* ops_per_thread, void* stream)
{
    if (!d_sink || blocks < 1 || iters < 1) return fail(BDG_ERR_ARG, "bad probe arguments");
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
    case 0: bdg::pipe_probe_kernel<0><<<blocks, bdg::NT, 0, st>>>(iters, d_sink); break;
    case 1: bdg::pipe_probe_kernel<1><<<blocks, bdg::NT, 0, st>>>(iters, d_sink); break;
    case 2: bdg::pipe_probe_kernel<2><<<blocks, bdg::NT, 0, st>>>(iters, d_sink); break;
    case 3: bdg::pipe_probe_kernel<3><<<blocks, bdg::NT, 0, st>>>(iters, d_sink); break;
    default: return fail(BDG_ERR_ARG, "unknown probe kind %d", kind);
    }
    g_launches++;
    CU_TRY(cudaGetLastError());
    if (ops_per_thread) *ops_per_thread = 64ull * (unsigned long long)iters;
    return BDG_OK;
}

// ---------------------------------------------------------------- host-buffer entry points
// ---- a-2  dedup + count in first-seen order (barcode_graph.py:192-204) ------------------------------------
// The keys of the n reads sit in dd[0] on entry (put there by the caller).  keep_map: compute the read -> first-seen
// position map and leave it in dd[7] (download it too when read_to_distinct != NULL).
static int dedup_core(DevCtx& c, uint32_t n, uint32_t* distinct, uint32_t* counts, uint32_t* read_to_distinct, uint32_t* sorted_pos,
                      bool keep_map, size_t* n_distinct, uint32_t* sorted_distinct = nullptr)
{
    cudaStream_t st = c.stream;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    const size_t R = n;
    // dd[0] keys, [1] idx, [2] sorted keys, [3] sorted idx, [4] heads, [5] inclusive scan, [6] run key, [7] run first, [8] run start, [9] scratch
    for (int k = 1; k < 9; k++) if (int e = ensure(c.dd[k], R * 4)) return e;
    uint32_t *d_k = (uint32_t*)c.dd[0].p, *d_i = (uint32_t*)c.dd[1].p, *d_sk = (uint32_t*)c.dd[2].p, *d_si = (uint32_t*)c.dd[3].p;
    uint32_t *d_head = (uint32_t*)c.dd[4].p, *d_scan = (uint32_t*)c.dd[5].p, *d_rk = (uint32_t*)c.dd[6].p, *d_rf = (uint32_t*)c.dd[7].p, *d_rs = (uint32_t*)c.dd[8].p;
    size_t tmp1 = 0, tmp2 = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp1, d_k, d_sk, d_i, d_si, (int)n, 0, 32, st));
    CU_TRY(cub::DeviceScan::InclusiveSum(nullptr, tmp2, d_head, d_scan, (int)n, st));
    if (int e = ensure(c.dd[9], std::max(tmp1, tmp2))) return e;
    const int blocks = (int)std::min<size_t>((R + 255) / 256, (size_t)c.sms * 8);
    bdg::iota_kernel<<<blocks, 256, 0, st>>>(d_i, n);
    CU_TRY(cub::DeviceRadixSort::SortPairs(c.dd[9].p, tmp1, d_k, d_sk, d_i, d_si, (int)n, 0, 32, st));
    bdg::dedup_heads_kernel<<<blocks, 256, 0, st>>>(d_sk, n, d_head);
    CU_TRY(cub::DeviceScan::InclusiveSum(c.dd[9].p, tmp2, d_head, d_scan, (int)n, st));
    bdg::dedup_runs_kernel<<<blocks, 256, 0, st>>>(d_sk, d_si, d_head, d_scan, n, d_rk, d_rf, d_rs);
    uint32_t n_runs = 0;
    CU_TRY(cudaMemcpyAsync(&n_runs, d_scan + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    g_launches += 3;
    // order the runs by their first read index: reuse keys/idx buffers (d_k := sorted first indices, d_i := order), heads := iota
    bdg::iota_kernel<<<blocks, 256, 0, st>>>(d_head, n_runs);
    CU_TRY(cub::DeviceRadixSort::SortPairs(c.dd[9].p, tmp1, d_rf, d_k, d_head, d_i, (int)n_runs, 0, 32, st));
    // d_sk is free again after dedup_runs: distinct / counts / pos_of_run live in d_sk, d_k, d_head
    uint32_t *d_distinct = d_sk, *d_counts = d_k, *d_pos = d_head;
    bdg::dedup_finish_kernel<<<blocks, 256, 0, st>>>(d_i, d_rk, d_rs, n_runs, n, d_distinct, d_counts, d_pos);
    g_launches += 2;
    if (distinct) CU_TRY(cudaMemcpyAsync(distinct, d_distinct, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, st));
    if (counts) CU_TRY(cudaMemcpyAsync(counts, d_counts, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, st));
    if (sorted_pos) CU_TRY(cudaMemcpyAsync(sorted_pos, d_i, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, st));   // order[pos] = run = ascending position
    if (sorted_distinct) CU_TRY(cudaMemcpyAsync(sorted_distinct, d_rk, (size_t)n_runs * 4, cudaMemcpyDeviceToHost, st));   // runs are numbered in key order
    if (keep_map || read_to_distinct) {
        bdg::dedup_scatter_kernel<<<blocks, 256, 0, st>>>(d_si, d_scan, d_pos, n, d_rf);   // d_rf is free after the run sort
        g_launches++;
        if (read_to_distinct) CU_TRY(cudaMemcpyAsync(read_to_distinct, d_rf, R * 4, cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(st));
    *n_distinct = n_runs;
    return BDG_OK;
}

int bdg_dedup_first_seen(const uint32_t* ranks, size_t R, uint32_t* distinct, uint32_t* counts, uint32_t* read_to_distinct,
                         uint32_t* sorted_pos, size_t* n_distinct)
{
    if (!n_distinct) return fail(BDG_ERR_ARG, "NULL n_distinct pointer");
    *n_distinct = 0;
    if (R == 0) return BDG_OK;
    if (!ranks || !distinct || !counts) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (R > 0x7FFFFFFFull) return fail(BDG_ERR_ARG, "more than 2^31 reads in one call");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    CU_TRY(cudaSetDevice(c.dev));
    c.map_token = 0;                                  // the workspaces are about to be reused: an older read map dies here
    if (cudaError_t e = (cudaError_t)c.dd[0].ensure(R * 4))
        return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", R * 4, cudaGetErrorString(e));
    CU_TRY(cudaMemcpyAsync(c.dd[0].p, ranks, R * 4, cudaMemcpyHostToDevice, c.stream));
    return dedup_core(c, (uint32_t)R, distinct, counts, read_to_distinct, sorted_pos, false, n_distinct);
}

// ---- f-1 + f-2 on the device: masked dedup that keeps the read map resident, and the per-read gather that uses it ------
int bdg_dedup_reads(const uint32_t* ranks, const uint8_t* valid, size_t R_all, uint32_t* distinct, uint32_t* counts, uint32_t* sorted_pos,
                    uint32_t* sorted_distinct, size_t* n_distinct, size_t* n_valid, unsigned long long* token)
{
    if (!n_distinct || !n_valid || !token) return fail(BDG_ERR_ARG, "NULL result pointer");
    *n_distinct = 0; *n_valid = 0; *token = 0;
    if (R_all == 0) return BDG_OK;
    if (!ranks) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (R_all > 0x7FFFFFFFull) return fail(BDG_ERR_ARG, "more than 2^31 reads in one call");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    CU_TRY(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    c.map_token = 0;
    c.cl_token = 0;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    if (int e = ensure(c.dd[0], R_all * 4)) return e;
    size_t n_reads = R_all;
    c.map_masked = valid != nullptr;
    if (valid) {
        // as[0] all ranks, as[1] valid bytes, as[2] exclusive scan of valid (the index of a valid row among the valid rows), as[3] scratch
        if (int e = ensure(c.as[0], R_all * 4)) return e;
        if (int e = ensure(c.as[1], R_all)) return e;
        if (int e = ensure(c.as[2], (R_all + 1) * 4)) return e;
        const uint8_t* d_valid = (const uint8_t*)c.as[1].p;
        uint32_t* d_excl = (uint32_t*)c.as[2].p;
        auto flags = thrust::make_transform_iterator(d_valid, bdg::NonZero());
        size_t tmp = 0;
        CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, flags, d_excl, (int)R_all, st));
        if (int e = ensure(c.as[3], tmp)) return e;
        CU_TRY(cudaMemcpyAsync(c.as[0].p, ranks, R_all * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemcpyAsync(c.as[1].p, valid, R_all, cudaMemcpyHostToDevice, st));
        CU_TRY(cub::DeviceScan::ExclusiveSum(c.as[3].p, tmp, flags, d_excl, (int)R_all, st));
        const int blocks = (int)std::min<size_t>((R_all + 255) / 256, (size_t)c.sms * 8);
        bdg::compact_valid_kernel<<<blocks, 256, 0, st>>>((const uint32_t*)c.as[0].p, d_valid, d_excl, (uint32_t)R_all, (uint32_t*)c.dd[0].p);
        g_launches++;
        uint32_t last_excl = 0;
        uint8_t last_valid = 0;
        CU_TRY(cudaMemcpyAsync(&last_excl, d_excl + (R_all - 1), 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&last_valid, d_valid + (R_all - 1), 1, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        n_reads = (size_t)last_excl + (last_valid ? 1 : 0);
    } else {
        CU_TRY(cudaMemcpyAsync(c.dd[0].p, ranks, R_all * 4, cudaMemcpyHostToDevice, st));
    }
    *n_valid = n_reads;
    c.map_rows = R_all; c.map_reads = n_reads; c.map_distinct = 0;
    if (n_reads == 0) return BDG_OK;
    if (int rc = dedup_core(c, (uint32_t)n_reads, distinct, counts, nullptr, sorted_pos, true, n_distinct, sorted_distinct)) return rc;
    c.map_distinct = *n_distinct;
    c.map_token = ++c.map_serial;
    *token = c.map_token;
    return BDG_OK;
}

// What bdg_dedup_reads left on the device, on demand (any pointer may be NULL): the arrays of its output list.
int bdg_dedup_fetch(unsigned long long token, uint32_t* distinct, uint32_t* counts, uint32_t* sorted_pos, uint32_t* sorted_distinct)
{
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    if (token == 0 || token != c.map_token) return fail(BDG_ERR_ARG, "stale read-map token: a later dedup call has reused the workspaces");
    CU_TRY(cudaSetDevice(c.dev));
    const size_t n = c.map_distinct;
    if (distinct) CU_TRY(cudaMemcpyAsync(distinct, c.dd[2].p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    if (counts) CU_TRY(cudaMemcpyAsync(counts, c.dd[0].p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    if (sorted_pos) CU_TRY(cudaMemcpyAsync(sorted_pos, c.dd[1].p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    if (sorted_distinct) CU_TRY(cudaMemcpyAsync(sorted_distinct, c.dd[6].p, n * 4, cudaMemcpyDeviceToHost, c.stream));
    CU_TRY(cudaStreamSynchronize(c.stream));
    return BDG_OK;
}

// ---- a-6 / centre selection, barcode_graph.py:252-267, over the distinct barcodes a dedup call left on the device ----
// cutoff = max(mean(counts of the first n_cells barcodes in first-seen order) / 5, 5) (:255-256); the barcodes with
// count > cutoff in count-descending order, ties in first-seen order (the head of `bc_by_counts`, :253), with their counts and
// - when a whitelist is given - their membership (:264).  The short walk over that head stays with the caller.
int bdg_centres_above(unsigned long long token, size_t n_cells, const uint32_t* sorted_wl, size_t W, uint32_t* top_ranks, uint32_t* top_counts,
                      uint8_t* top_hits, size_t cap, size_t* n_above, double* cutoff)
{
    if (!n_above || !cutoff) return fail(BDG_ERR_ARG, "NULL result pointer");
    *n_above = 0; *cutoff = 0.0;
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    if (token == 0 || token != c.map_token) return fail(BDG_ERR_ARG, "stale read-map token: a later dedup call has reused the workspaces");
    const size_t N = c.map_distinct;
    const size_t first = std::min(n_cells, N);
    if (first == 0) return fail(BDG_ERR_ARG, "mean requires at least one data point");
    CU_TRY(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    const uint32_t* d_distinct = (const uint32_t*)c.dd[2].p;
    const uint32_t* d_counts = (const uint32_t*)c.dd[0].p;
    // ce[0] selected positions, [1] keys, [2] sorted keys, [3] sorted positions, [4] top ranks, [5] top counts, [6] hits, [7] scalars + cub scratch
    if (int e = ensure(c.ce[7], 64)) return e;
    unsigned long long* d_sum = (unsigned long long*)c.ce[7].p;
    uint32_t* d_nsel = (uint32_t*)((char*)c.ce[7].p + 8);
    CU_TRY(cudaMemsetAsync(d_sum, 0, 16, st));
    bdg::sum_first_kernel<<<(int)std::min<size_t>((first + 255) / 256, (size_t)c.sms * 4), 256, 0, st>>>(d_counts, (uint32_t)first, d_sum);
    g_launches++;
    unsigned long long sum = 0;
    CU_TRY(cudaMemcpyAsync(&sum, d_sum, 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const double cut = std::max(((double)sum / (double)first) / 5.0, 5.0);
    *cutoff = cut;
    const uint32_t thr = (uint32_t)std::floor(cut);          // counts are integers: count > cutoff <=> count > floor(cutoff)
    if (int e = ensure(c.ce[0], N * 4)) return e;
    thrust::counting_iterator<uint32_t> all(0);
    bdg::CountAbove pred{d_counts, thr};
    size_t tmp = 0;
    CU_TRY(cub::DeviceSelect::If(nullptr, tmp, all, (uint32_t*)c.ce[0].p, d_nsel, (int)N, pred, st));
    if (int e = ensure(c.ce[8], tmp)) return e;
    CU_TRY(cub::DeviceSelect::If(c.ce[8].p, tmp, all, (uint32_t*)c.ce[0].p, d_nsel, (int)N, pred, st));
    uint32_t nsel = 0;
    CU_TRY(cudaMemcpyAsync(&nsel, d_nsel, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    *n_above = nsel;
    if (nsel == 0) return BDG_OK;
    if (nsel > cap) return fail(BDG_ERR_CAPACITY, "%u barcodes above the cutoff but room for %zu", nsel, cap);
    if (!top_ranks || !top_counts) return fail(BDG_ERR_ARG, "NULL pointer argument");
    for (int k = 1; k <= 5; k++) if (int e = ensure(c.ce[k], (size_t)nsel * 4)) return e;
    if (int e = ensure(c.ce[6], nsel)) return e;
    const int nb = (int)std::min<size_t>(((size_t)nsel + 255) / 256, (size_t)c.sms * 4);
    bdg::centres_keys_kernel<<<nb, 256, 0, st>>>((const uint32_t*)c.ce[0].p, d_counts, nsel, (uint32_t*)c.ce[1].p);
    size_t tmp2 = 0;
    CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, (const uint32_t*)c.ce[1].p, (uint32_t*)c.ce[2].p, (const uint32_t*)c.ce[0].p, (uint32_t*)c.ce[3].p, (int)nsel, 0, 32, st));
    if (int e = ensure(c.ce[8], tmp2)) return e;
    CU_TRY(cub::DeviceRadixSort::SortPairs(c.ce[8].p, tmp2, (const uint32_t*)c.ce[1].p, (uint32_t*)c.ce[2].p, (const uint32_t*)c.ce[0].p, (uint32_t*)c.ce[3].p, (int)nsel, 0, 32, st));
    bdg::centres_gather_kernel<<<nb, 256, 0, st>>>((const uint32_t*)c.ce[3].p, d_distinct, d_counts, nsel, (uint32_t*)c.ce[4].p, (uint32_t*)c.ce[5].p);
    g_launches += 2;
    CU_TRY(cudaGetLastError());
    if (top_hits) {
        if (W > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "size exceeds 2^32");
        if (W && !sorted_wl) return fail(BDG_ERR_ARG, "NULL pointer argument");
        if (int e = ensure(c.io[0], std::max<size_t>(W, 1) * 4)) return e;
        if (int e = ensure(c.io[3], 8)) return e;
        unsigned long long* d_bad = (unsigned long long*)c.io[3].p;
        CU_TRY(cudaMemcpyAsync(c.io[0].p, sorted_wl, W * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemsetAsync(d_bad, 0xFF, 8, st));
        if (W > 1) {
            bdg::sorted_check_kernel<false><<<(int)std::min<size_t>((W + 255) / 256, (size_t)c.sms * 8), 256, 0, st>>>((const uint32_t*)c.io[0].p, (uint32_t)W, d_bad);
            g_launches++;
        }
        if (int rc = bdg_dev_member_sorted((const uint32_t*)c.io[0].p, W, (const uint32_t*)c.ce[4].p, nsel, (uint8_t*)c.ce[6].p, st)) return rc;
        unsigned long long first_bad = ~0ull;
        CU_TRY(cudaMemcpyAsync(top_hits, c.ce[6].p, nsel, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(&first_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(top_ranks, c.ce[4].p, (size_t)nsel * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaMemcpyAsync(top_counts, c.ce[5].p, (size_t)nsel * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (first_bad != ~0ull) return fail(BDG_ERR_ARG, "whitelist not sorted at index %llu", first_bad);
        return BDG_OK;
    }
    CU_TRY(cudaMemcpyAsync(top_ranks, c.ce[4].p, (size_t)nsel * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(top_counts, c.ce[5].p, (size_t)nsel * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return BDG_OK;
}

// The stretch of `bc_by_counts` right behind the head bdg_centres_above returns (barcode_graph.py:273-276 runs into it when too few
// centres were found): the first `need` barcodes with count <= floor(cutoff) in count-descending order, ties in first-seen
// order.  One stable stream compaction per count value, from floor(cutoff) downwards, until enough are found.
int bdg_centres_rest(unsigned long long token, double cutoff, size_t need, uint32_t* out_ranks, size_t* n_out)
{
    if (!n_out) return fail(BDG_ERR_ARG, "NULL result pointer");
    *n_out = 0;
    if (need == 0) return BDG_OK;
    if (!out_ranks) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    if (token == 0 || token != c.map_token) return fail(BDG_ERR_ARG, "stale read-map token: a later dedup call has reused the workspaces");
    const size_t N = c.map_distinct;
    CU_TRY(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    const uint32_t* d_distinct = (const uint32_t*)c.dd[2].p;
    const uint32_t* d_counts = (const uint32_t*)c.dd[0].p;
    if (int e = ensure(c.ce[0], N * 4)) return e;
    if (int e = ensure(c.ce[4], std::min(need, N) * 4)) return e;
    if (int e = ensure(c.ce[5], std::min(need, N) * 4)) return e;
    if (int e = ensure(c.ce[7], 64)) return e;
    uint32_t* d_nsel = (uint32_t*)((char*)c.ce[7].p + 8);
    thrust::counting_iterator<uint32_t> all(0);
    size_t tmp = 0;
    CU_TRY(cub::DeviceSelect::If(nullptr, tmp, all, (uint32_t*)c.ce[0].p, d_nsel, (int)N, bdg::CountIs{d_counts, 0u}, st));
    if (int e = ensure(c.ce[8], tmp)) return e;
    size_t got = 0;
    for (long long v = (long long)std::floor(cutoff); v >= 1 && got < need; v--) {
        CU_TRY(cub::DeviceSelect::If(c.ce[8].p, tmp, all, (uint32_t*)c.ce[0].p, d_nsel, (int)N, bdg::CountIs{d_counts, (uint32_t)v}, st));
        uint32_t nsel = 0;
        CU_TRY(cudaMemcpyAsync(&nsel, d_nsel, 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        const size_t take = std::min<size_t>(nsel, need - got);
        if (take == 0) continue;
        const int nb = (int)std::min<size_t>((take + 255) / 256, (size_t)c.sms * 4);
        bdg::centres_gather_kernel<<<nb, 256, 0, st>>>((const uint32_t*)c.ce[0].p, d_distinct, d_counts, (uint32_t)take, (uint32_t*)c.ce[4].p, (uint32_t*)c.ce[5].p);
        g_launches++;
        CU_TRY(cudaMemcpyAsync(out_ranks + got, c.ce[4].p, take * 4, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        got += take;
    }
    *n_out = got;
    return BDG_OK;
}

// bdg_assign_reads with a 5-byte result per row (centre barcode + "has a centre" byte).  centre_idx == NULL: the clustering a
// bdg_cluster_resident call left on the device is used in place (nothing is uploaded).
int bdg_assign_reads32(unsigned long long token, const int32_t* centre_idx, size_t N, uint32_t* centre_per_row, uint8_t* has_centre, size_t R_all,
                       size_t* n_assigned)
{
    if (n_assigned) *n_assigned = 0;
    if (R_all == 0) return BDG_OK;
    if (!centre_per_row || !has_centre) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    if (token == 0 || token != c.map_token) return fail(BDG_ERR_ARG, "stale read-map token: a later dedup call has reused the workspaces");
    if (N != c.map_distinct || R_all != c.map_rows) return fail(BDG_ERR_ARG, "sizes do not match the dedup call the token came from");
    if (!centre_idx && N && c.cl_token != token) return fail(BDG_ERR_ARG, "no clustering result of this dedup call is resident on the device (bdg_cluster_resident)");
    CU_TRY(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    // as[4] centre index per node (upload), as[5] centre value + has byte per distinct barcode, as[6] result per row, as[7] counter
    if (int e = ensure(c.as[5], std::max<size_t>(N, 1) * 5)) return e;
    if (int e = ensure(c.as[6], R_all * 5)) return e;
    if (int e = ensure(c.as[7], 8)) return e;
    const int32_t* d_ci = (const int32_t*)c.cl[1].p;
    if (centre_idx) {
        if (int e = ensure(c.as[4], std::max<size_t>(N, 1) * 4)) return e;
        CU_TRY(cudaMemcpyAsync(c.as[4].p, centre_idx, N * 4, cudaMemcpyHostToDevice, st));
        d_ci = (const int32_t*)c.as[4].p;
    }
    CU_TRY(cudaMemsetAsync(c.as[7].p, 0, 8, st));
    uint32_t* d_cv = (uint32_t*)c.as[5].p;
    uint8_t* d_ch = (uint8_t*)c.as[5].p + std::max<size_t>(N, 1) * 4;
    uint32_t* d_out = (uint32_t*)c.as[6].p;
    uint8_t* d_oh = (uint8_t*)c.as[6].p + R_all * 4;
    if (N) {
        const int nb = (int)std::min<size_t>((N + 255) / 256, (size_t)c.sms * 8);
        bdg::centre_of_distinct32_kernel<<<nb, 256, 0, st>>>(d_ci, (const uint32_t*)c.dd[1].p /* first-seen -> node */,
                                                             (const uint32_t*)c.dd[6].p /* node -> barcode */, (uint32_t)N, d_cv, d_ch);
    }
    const int rb = (int)std::min<size_t>((R_all + 255) / 256, (size_t)c.sms * 8);
    bdg::assign_reads32_kernel<<<rb, 256, 0, st>>>(d_cv, d_ch, (const uint32_t*)c.dd[7].p, c.map_masked ? (const uint8_t*)c.as[1].p : nullptr,
                                                   c.map_masked ? (const uint32_t*)c.as[2].p : nullptr, (uint32_t)R_all, d_out, d_oh,
                                                   (unsigned long long*)c.as[7].p);
    g_launches += 2;
    CU_TRY(cudaGetLastError());
    unsigned long long cnt = 0;
    CU_TRY(cudaMemcpyAsync(centre_per_row, d_out, R_all * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(has_centre, d_oh, R_all, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&cnt, c.as[7].p, 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (n_assigned) *n_assigned = (size_t)cnt;
    return BDG_OK;
}

int bdg_assign_reads(unsigned long long token, const int32_t* centre_idx, size_t N, uint64_t* centre_per_row, size_t R_all, size_t* n_assigned)
{
    if (n_assigned) *n_assigned = 0;
    if (R_all == 0) return BDG_OK;
    if (!centre_per_row || (N && !centre_idx)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    if (token == 0 || token != c.map_token) return fail(BDG_ERR_ARG, "stale read-map token: a later dedup call has reused the workspaces");
    if (N != c.map_distinct || R_all != c.map_rows) return fail(BDG_ERR_ARG, "sizes do not match the dedup call the token came from");
    CU_TRY(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    // as[4] centre index per node (upload), as[5] centre value per distinct barcode in first-seen order, as[6] result per row, as[7] counter
    if (int e = ensure(c.as[4], std::max<size_t>(N, 1) * 4)) return e;
    if (int e = ensure(c.as[5], std::max<size_t>(N, 1) * 8)) return e;
    if (int e = ensure(c.as[6], R_all * 8)) return e;
    if (int e = ensure(c.as[7], 8)) return e;
    CU_TRY(cudaMemcpyAsync(c.as[4].p, centre_idx, N * 4, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(c.as[7].p, 0, 8, st));
    if (N) {
        const int nb = (int)std::min<size_t>((N + 255) / 256, (size_t)c.sms * 8);
        bdg::centre_of_distinct_kernel<<<nb, 256, 0, st>>>((const int32_t*)c.as[4].p, (const uint32_t*)c.dd[1].p /* first-seen -> node */,
                                                           (const uint32_t*)c.dd[6].p /* node -> barcode */, (uint32_t)N, (uint64_t*)c.as[5].p);
    }
    const int rb = (int)std::min<size_t>((R_all + 255) / 256, (size_t)c.sms * 8);
    bdg::assign_reads_kernel<<<rb, 256, 0, st>>>((const uint64_t*)c.as[5].p, (const uint32_t*)c.dd[7].p, c.map_masked ? (const uint8_t*)c.as[1].p : nullptr,
                                                 c.map_masked ? (const uint32_t*)c.as[2].p : nullptr, (uint32_t)R_all, (uint64_t*)c.as[6].p,
                                                 (unsigned long long*)c.as[7].p);
    g_launches += 2;
    CU_TRY(cudaGetLastError());
    unsigned long long cnt = 0;
    CU_TRY(cudaMemcpyAsync(centre_per_row, c.as[6].p, R_all * 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(&cnt, c.as[7].p, 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (n_assigned) *n_assigned = (size_t)cnt;
    return BDG_OK;
}

int bdg_pack16(const char* seqs, size_t R, uint32_t* out, uint8_t* valid)
{
    if (R == 0) return BDG_OK;
    if (!seqs || !out || !valid) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    CU_TRY(cudaSetDevice(c.dev));
    const size_t sizes[3] = {R * 16, R * 4, R};               // grow-only workspaces: letters, ranks, validity
    for (int k = 0; k < 3; k++)
        if (cudaError_t e = (cudaError_t)c.io[k].ensure(sizes[k]))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", sizes[k], cudaGetErrorString(e));
    CU_TRY(cudaMemcpyAsync(c.io[0].p, seqs, R * 16, cudaMemcpyHostToDevice, c.stream));
    if (int rc = bdg_dev_pack16((const char*)c.io[0].p, R, (uint32_t*)c.io[1].p, (uint8_t*)c.io[2].p, c.stream)) return rc;
    CU_TRY(cudaMemcpyAsync(out, c.io[1].p, R * 4, cudaMemcpyDeviceToHost, c.stream));
    CU_TRY(cudaMemcpyAsync(valid, c.io[2].p, R, cudaMemcpyDeviceToHost, c.stream));
    CU_TRY(cudaStreamSynchronize(c.stream));
    return BDG_OK;
}


// badger.py:82-88 for the array pipeline: the whitelist records packed (a-1), the invalid ones dropped, sorted and made
// distinct on the device (cub radix sort + unique); out_sorted has room for R entries, *n receives the number kept.
int bdg_pack16_sorted(const char* seqs, size_t R, uint32_t* out_sorted, size_t* n)
{
    if (!n) return fail(BDG_ERR_ARG, "NULL result pointer");
    *n = 0;
    if (R == 0) return BDG_OK;
    if (!seqs || !out_sorted) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (R > 0x7FFFFFFFull) return fail(BDG_ERR_ARG, "more than 2^31 records in one call");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    CU_TRY(cudaSetDevice(c.dev));
    cudaStream_t st = c.stream;
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    // io[0] letters, io[1] ranks, io[2] validity; ce[0] kept ranks, ce[1] sorted, ce[2] distinct, ce[7] count, ce[8] cub scratch
    if (int e = ensure(c.io[0], R * 16)) return e;
    if (int e = ensure(c.io[1], R * 4)) return e;
    if (int e = ensure(c.io[2], R)) return e;
    for (int k = 0; k < 3; k++) if (int e = ensure(c.ce[k], R * 4)) return e;
    if (int e = ensure(c.ce[7], 64)) return e;
    uint32_t* d_n = (uint32_t*)c.ce[7].p;
    CU_TRY(cudaMemcpyAsync(c.io[0].p, seqs, R * 16, cudaMemcpyHostToDevice, st));
    if (int rc = bdg_dev_pack16((const char*)c.io[0].p, R, (uint32_t*)c.io[1].p, (uint8_t*)c.io[2].p, st)) return rc;
    size_t t1 = 0, t2 = 0, t3 = 0;
    CU_TRY(cub::DeviceSelect::Flagged(nullptr, t1, (const uint32_t*)c.io[1].p, (const uint8_t*)c.io[2].p, (uint32_t*)c.ce[0].p, d_n, (int)R, st));
    CU_TRY(cub::DeviceRadixSort::SortKeys(nullptr, t2, (const uint32_t*)c.ce[0].p, (uint32_t*)c.ce[1].p, (int)R, 0, 32, st));
    CU_TRY(cub::DeviceSelect::Unique(nullptr, t3, (const uint32_t*)c.ce[1].p, (uint32_t*)c.ce[2].p, d_n, (int)R, st));
    if (int e = ensure(c.ce[8], std::max(t1, std::max(t2, t3)))) return e;
    CU_TRY(cub::DeviceSelect::Flagged(c.ce[8].p, t1, (const uint32_t*)c.io[1].p, (const uint8_t*)c.io[2].p, (uint32_t*)c.ce[0].p, d_n, (int)R, st));
    uint32_t kept = 0;
    CU_TRY(cudaMemcpyAsync(&kept, d_n, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (kept == 0) return BDG_OK;
    CU_TRY(cub::DeviceRadixSort::SortKeys(c.ce[8].p, t2, (const uint32_t*)c.ce[0].p, (uint32_t*)c.ce[1].p, (int)kept, 0, 32, st));
    CU_TRY(cub::DeviceSelect::Unique(c.ce[8].p, t3, (const uint32_t*)c.ce[1].p, (uint32_t*)c.ce[2].p, d_n, (int)kept, st));
    uint32_t distinct = 0;
    CU_TRY(cudaMemcpyAsync(&distinct, d_n, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    CU_TRY(cudaMemcpyAsync(out_sorted, c.ce[2].p, (size_t)distinct * 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    *n = distinct;
    return BDG_OK;
}

// One device's share of a host-buffer edge build: upload, launch, read the count back (re-run once with the exact
// size when the guess was too small).  Runs on its own host thread when several devices take part, because the
// sparse passes read a tile count back between their kernels.
// src_dev < 0: `sorted` is a host array; else it lives on device src_dev (peer copy over NVLink, or nothing to copy at all
// when that is this device's own workspace).
static int edges_on_device(DevCtx& c, const uint32_t* sorted, size_t N, int t, int part, int nparts, size_t* n_out, int src_dev = -1)
{
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    CU_TRY(cudaSetDevice(c.dev));
    if (int e = ensure(c.sorted, std::max<size_t>(N, 1) * 4)) return e;
    if (int e = ensure(c.count, 2 * sizeof(unsigned long long))) return e;      // [edge count | first unsorted index]
    if (src_dev < 0) CU_TRY(cudaMemcpyAsync(c.sorted.p, sorted, N * 4, cudaMemcpyHostToDevice, c.stream));
    else if (src_dev == c.dev) CU_TRY(cudaMemcpyAsync(c.sorted.p, sorted, N * 4, cudaMemcpyDeviceToDevice, c.stream));
    else CU_TRY(cudaMemcpyPeerAsync(c.sorted.p, c.dev, sorted, src_dev, N * 4, c.stream));
    // the input must be strictly increasing: checked on the device, read back together with the edge count
    unsigned long long* d_bad = (unsigned long long*)c.count.p + 1;
    unsigned long long first_bad = ~0ull;
    CU_TRY(cudaMemsetAsync(d_bad, 0xFF, 8, c.stream));
    if (N > 1) {
        bdg::sorted_check_kernel<true><<<(int)std::min<size_t>((N + 255) / 256, (size_t)c.sms * 8), 256, 0, c.stream>>>((const uint32_t*)c.sorted.p, (uint32_t)N, d_bad);
        g_launches++;
    }
    size_t cap = std::max(edge_cap_guess(N, t, nparts), c.ea.cap / 4);
    unsigned long long count = 0;
    bool done = false;
    for (int attempt = 0; attempt < 4 && !done; attempt++) {
        if (int e = ensure(c.ea, cap * 4)) return e;
        if (int e = ensure(c.eb, cap * 4)) return e;
        if (int e = ensure(c.ed, cap)) return e;
        if (int e = launch_edges((const uint32_t*)c.sorted.p, N, t, part, nparts, (uint32_t*)c.ea.p, (uint32_t*)c.eb.p, (uint8_t*)c.ed.p, cap,
                                 (unsigned long long*)c.count.p, c.stream, &c)) return e;
        unsigned long long both[2] = {0, 0};
        CU_TRY(cudaMemcpyAsync(both, c.count.p, sizeof(both), cudaMemcpyDeviceToHost, c.stream));
        CU_TRY(cudaStreamSynchronize(c.stream));
        count = both[0];
        first_bad = both[1];
        if (first_bad != ~0ull) return fail(BDG_ERR_ARG, "input not strictly increasing at index %llu", first_bad);
        if (count >> 63) {                     // rare: a sparse pass listed more tiles than the list holds; grow it and run again
            unsigned long long hdr[(PLAN_HDR / 8) * bdg::MAX_PASSES];
            CU_TRY(cudaMemcpy(hdr, c.plan.p, sizeof(hdr), cudaMemcpyDeviceToHost));
            unsigned long long need = 0;
            for (int p = 0; p < bdg::MAX_PASSES; p++) need = std::max(need, hdr[(PLAN_HDR / 8) * p + HDR_LIST / 8]);
            for (int p = 0; p < bdg::MAX_PASSES; p++)
                if (int e = ensure(c.tile_list[p], (size_t)need * sizeof(uint2))) return e;
        } else if (count > cap) {              // rare: the capacity guess was too small; the edge set is deterministic, so run again
            cap = (size_t)count;
        } else {
            done = true;
        }
    }
    if (!done) return fail(BDG_ERR_CUDA, "edge construction did not settle after growing its buffers (count %llu)", count);
    c.generation++;
    *n_out = (size_t)count;
    return BDG_OK;
}

static int edges_on_devices(const uint32_t* sorted, size_t N, int t, const std::vector<int>& ctx_idx,
                            const std::vector<int>& parts, int nparts, bdg_edges* res, int src_dev = -1)
{
    const double t0 = now_ms();
    const size_t G = ctx_idx.size();
    std::vector<int> rcs(G, BDG_OK);
    std::vector<std::string> errs(G);
    std::vector<size_t> counts(G, 0);
    std::vector<double> took(G, 0.0);
    auto work = [&](size_t g) {
        const double w0 = now_ms();
        rcs[g] = edges_on_device(g_ctx[ctx_idx[g]], sorted, N, t, parts[g], nparts, &counts[g], src_dev);
        took[g] = now_ms() - w0;
        if (rcs[g]) errs[g] = g_err;       // g_err is thread-local
    };
    if (G == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (size_t g = 0; g < G; g++) th.emplace_back(work, g);
        for (auto& x : th) x.join();
    }
    if (!g_ctx.empty()) cudaSetDevice(g_ctx[0].dev);
    for (size_t g = 0; g < G; g++)
        if (rcs[g]) return fail(rcs[g], "%s", errs[g].c_str());
    size_t total = 0;
    for (size_t g = 0; g < G; g++) {
        res->ctx.push_back(ctx_idx[g]);
        res->count.push_back(counts[g]);
        res->gen.push_back(g_ctx[ctx_idx[g]].generation);
        total += counts[g];
    }
    if (getenv("BDG_TRACE")) {
        fprintf(stderr, "[bdg] edges N=%zu t=%d devices=%zu: upload + kernels %.2f ms, edges %zu; per device (ms / edges):", N, t, G, now_ms() - t0, total);
        for (size_t g = 0; g < G; g++) fprintf(stderr, " %.1f/%zu", took[g], counts[g]);
        fprintf(stderr, "\n");
    }
    return BDG_OK;
}

int bdg_edges_build(const uint32_t* sorted_unique, size_t N, int t, bdg_edges** out)
{
    if (!out || (N && !sorted_unique)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    if (int rc = need_ctx()) return rc;                    // (the strictly-increasing check runs on the device)
    bdg_edges* res = new (std::nothrow) bdg_edges();
    if (!res) return fail(BDG_ERR_OOM, "host allocation failed");
    std::vector<int> idx, parts;
    const int G = (int)g_ctx.size();
    for (int g = 0; g < G; g++) { idx.push_back(g); parts.push_back(g); }
    int rc = BDG_OK;
    try { rc = edges_on_devices(sorted_unique, N, t, idx, parts, G, res); }
    catch (const std::bad_alloc&) { rc = fail(BDG_ERR_OOM, "host allocation failed"); }
    catch (const std::exception& e) { rc = fail(BDG_ERR_CUDA, "host thread failure: %s", e.what()); }
    if (rc) { delete res; return rc; }
    *out = res;
    return BDG_OK;
}

// The same over the ascending distinct barcodes a bdg_dedup_reads call left on the first device: nothing is uploaded, the other
// devices fetch the array with peer copies.
int bdg_edges_build_resident(unsigned long long token, int t, bdg_edges** out)
{
    if (!out) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    if (int rc = need_ctx()) return rc;
    DevCtx& c0 = g_ctx[0];
    if (token == 0 || token != c0.map_token) return fail(BDG_ERR_ARG, "stale read-map token: a later dedup call has reused the workspaces");
    bdg_edges* res = new (std::nothrow) bdg_edges();
    if (!res) return fail(BDG_ERR_OOM, "host allocation failed");
    std::vector<int> idx, parts;
    const int G = (int)g_ctx.size();
    for (int g = 0; g < G; g++) { idx.push_back(g); parts.push_back(g); }
    int rc = BDG_OK;
    try { rc = edges_on_devices((const uint32_t*)c0.dd[6].p, c0.map_distinct, t, idx, parts, G, res, c0.dev); }
    catch (const std::bad_alloc&) { rc = fail(BDG_ERR_OOM, "host allocation failed"); }
    catch (const std::exception& e) { rc = fail(BDG_ERR_CUDA, "host thread failure: %s", e.what()); }
    if (rc) { delete res; return rc; }
    *out = res;
    return BDG_OK;
}

// One part's edges straight into caller buffers of `cap` entries each (page-locked ones copy at PCIe speed): with the join
// form the finished edges of earlier seed conditions cross PCIe while later ones are still being joined.  *count receives the
// number of edges found; when it exceeds cap only the first cap were stored and the call has to be repeated with more room.
int bdg_edges_build_into(const uint32_t* sorted_unique, size_t N, int t, int part, int nparts, uint32_t* a, uint32_t* b, uint8_t* d, size_t cap,
                         size_t* count)
{
    if (!count || (N && !sorted_unique) || (cap && (!a || !b || !d))) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *count = 0;
    if (nparts < 1 || part < 0 || part >= nparts) return fail(BDG_ERR_ARG, "part %d of %d is not a valid part", part, nparts);
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    auto ensure = [&](Buf& bf, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)bf.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    CU_TRY(cudaSetDevice(c.dev));
    const size_t dcap = std::max<size_t>(cap, 1);
    if (int e = ensure(c.sorted, std::max<size_t>(N, 1) * 4)) return e;
    if (int e = ensure(c.count, 2 * sizeof(unsigned long long))) return e;
    if (int e = ensure(c.ea, dcap * 4)) return e;
    if (int e = ensure(c.eb, dcap * 4)) return e;
    if (int e = ensure(c.ed, dcap)) return e;
    c.generation++;                                         // older handles on this device lose their edges
    CU_TRY(cudaMemcpyAsync(c.sorted.p, sorted_unique, N * 4, cudaMemcpyHostToDevice, c.stream));
    unsigned long long* d_bad = (unsigned long long*)c.count.p + 1;
    CU_TRY(cudaMemsetAsync(d_bad, 0xFF, 8, c.stream));
    if (N > 1) {
        bdg::sorted_check_kernel<true><<<(int)std::min<size_t>((N + 255) / 256, (size_t)c.sms * 8), 256, 0, c.stream>>>((const uint32_t*)c.sorted.p, (uint32_t)N, d_bad);
        g_launches++;
    }
    const bool streaming = t > 0 && N >= 2 && edge_mode_for(t, N) == 2;
    JoinStream js{a, b, d, cap};
    for (int attempt = 0; attempt < 3; attempt++) {
        if (int e = launch_edges((const uint32_t*)c.sorted.p, N, t, part, nparts, (uint32_t*)c.ea.p, (uint32_t*)c.eb.p, (uint8_t*)c.ed.p, cap,
                                 (unsigned long long*)c.count.p, c.stream, &c, streaming ? &js : nullptr)) return e;
        unsigned long long both[2] = {0, 0};
        CU_TRY(cudaMemcpyAsync(both, c.count.p, sizeof(both), cudaMemcpyDeviceToHost, c.stream));
        CU_TRY(cudaStreamSynchronize(c.stream));
        if (both[1] != ~0ull) return fail(BDG_ERR_ARG, "input not strictly increasing at index %llu", both[1]);
        if (both[0] >> 63) {                                // a sparse pass listed more tiles than its list holds: grow it and run again
            unsigned long long hdr[(PLAN_HDR / 8) * bdg::MAX_PASSES];
            CU_TRY(cudaMemcpy(hdr, c.plan.p, sizeof(hdr), cudaMemcpyDeviceToHost));
            unsigned long long need = 0;
            for (int p = 0; p < bdg::MAX_PASSES; p++) need = std::max(need, hdr[(PLAN_HDR / 8) * p + HDR_LIST / 8]);
            for (int p = 0; p < bdg::MAX_PASSES; p++)
                if (int e = ensure(c.tile_list[p], (size_t)need * sizeof(uint2))) return e;
            continue;
        }
        *count = (size_t)both[0];
        if (!streaming) {
            const size_t k = std::min<size_t>(*count, cap);
            if (k) {
                CU_TRY(cudaMemcpyAsync(a, c.ea.p, k * 4, cudaMemcpyDeviceToHost, c.stream));
                CU_TRY(cudaMemcpyAsync(b, c.eb.p, k * 4, cudaMemcpyDeviceToHost, c.stream));
                CU_TRY(cudaMemcpyAsync(d, c.ed.p, k, cudaMemcpyDeviceToHost, c.stream));
                CU_TRY(cudaStreamSynchronize(c.stream));
            }
        }
        return BDG_OK;
    }
    return fail(BDG_ERR_CUDA, "edge construction did not settle after growing its tile lists");
}

int bdg_edges_build_part(const uint32_t* sorted_unique, size_t N, int t, int part, int nparts, bdg_edges** out)
{
    if (!out || (N && !sorted_unique)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    if (nparts < 1 || part < 0 || part >= nparts) return fail(BDG_ERR_ARG, "part %d of %d is not a valid part", part, nparts);
    if (int rc = need_ctx()) return rc;
    bdg_edges* res = new (std::nothrow) bdg_edges();
    if (!res) return fail(BDG_ERR_OOM, "host allocation failed");
    int rc = BDG_OK;
    try { rc = edges_on_devices(sorted_unique, N, t, {0}, {part}, nparts, res); }
    catch (const std::bad_alloc&) { rc = fail(BDG_ERR_OOM, "host allocation failed"); }
    if (rc) { delete res; return rc; }
    *out = res;
    return BDG_OK;
}

size_t bdg_edges_count(const bdg_edges* e)
{
    size_t n = 0;
    if (e) for (size_t c : e->count) n += c;
    return n;
}

int bdg_edges_copy(const bdg_edges* e, uint32_t* a, uint32_t* b, uint8_t* d)
{
    if (!e) return fail(BDG_ERR_ARG, "NULL edge handle");
    const size_t n = bdg_edges_count(e);
    if (n && (!a || !b || !d)) return fail(BDG_ERR_ARG, "NULL output pointer");
    const double tc0 = now_ms();
    size_t off = 0;
    for (size_t g = 0; g < e->ctx.size(); g++) {
        if (e->ctx[g] < 0 || (size_t)e->ctx[g] >= g_ctx.size() || g_ctx[e->ctx[g]].generation != e->gen[g])
            return fail(BDG_ERR_ARG, "stale edge handle: a later edge build on the same device has reused its buffers");
        DevCtx& c = g_ctx[e->ctx[g]];
        const size_t k = e->count[g];
        if (k) {
            CU_TRY(cudaSetDevice(c.dev));
            CU_TRY(cudaMemcpyAsync(a + off, c.ea.p, k * 4, cudaMemcpyDeviceToHost, c.stream));
            CU_TRY(cudaMemcpyAsync(b + off, c.eb.p, k * 4, cudaMemcpyDeviceToHost, c.stream));
            CU_TRY(cudaMemcpyAsync(d + off, c.ed.p, k, cudaMemcpyDeviceToHost, c.stream));
        }
        off += k;
    }
    for (size_t g = 0; g < e->ctx.size(); g++)
        if (e->count[g]) { CU_TRY(cudaSetDevice(g_ctx[e->ctx[g]].dev)); CU_TRY(cudaStreamSynchronize(g_ctx[e->ctx[g]].stream)); }
    if (!g_ctx.empty()) cudaSetDevice(g_ctx[0].dev);
    if (getenv("BDG_TRACE")) fprintf(stderr, "[bdg] edges_copy %zu edges %.3f ms\n", n, now_ms() - tc0);
    return BDG_OK;
}

void bdg_edges_free(bdg_edges* e) { delete e; }


// ---- f-3  cluster(): barcode_graph.py:279-301 -------------------------------------------------------------
// d_ea / d_eb hold barcode VALUES on entry and node indices on return (converted in place).
// centre_idx / level may be NULL: the result then only stays on the device (cl[1] / cl[2]) for bdg_assign_reads*.
// values -> node indices of the E edges (ea, eb) on device c (its own copy of the sorted array), written to (oa, ob); on c.stream
static int index_edges(DevCtx& c, size_t N, const uint32_t* d_ea, const uint32_t* d_eb, size_t E, uint32_t* d_oa, uint32_t* d_ob)
{
    if (cudaError_t e = (cudaError_t)c.top16.ensure(65537 * 4))
        return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation: %s", cudaGetErrorString(e));
    const int nb = (int)std::min<size_t>((N + 256) / 256, (size_t)c.sms * 8);
    bdg::top_start_kernel<<<nb, 256, 0, c.stream>>>((const uint32_t*)c.sorted.p, (uint32_t)N, (uint32_t*)c.top16.p);
    if (E) {
        const int eb = (int)std::min<size_t>((E + 255) / 256, (size_t)c.sms * 16);
        bdg::cluster_index_kernel<<<eb, 256, 0, c.stream>>>((const uint32_t*)c.sorted.p, (const uint32_t*)c.top16.p, d_ea, d_eb, E, d_oa, d_ob);
    }
    g_launches += 2;
    CU_TRY(cudaGetLastError());
    return BDG_OK;
}

static int cluster_on_device(DevCtx& c, const uint32_t* d_sorted, size_t N, uint32_t* d_ea, uint32_t* d_eb, size_t E,
                             const uint32_t* centres, size_t C, int rounds, int32_t* centre_idx, uint8_t* level, size_t* n_has_edge = nullptr,
                             bool indexed = false)
{
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    cudaStream_t st = c.stream;
    if (int e = ensure(c.cl[0], std::max<size_t>(C, 1) * 4)) return e;
    if (int e = ensure(c.cl[1], N * 4)) return e;
    if (int e = ensure(c.cl[2], N)) return e;
    if (int e = ensure(c.cl[3], N * 4)) return e;
    if (int e = ensure(c.cl[4], N * 4)) return e;
    if (int e = ensure(c.cl[5], 16)) return e;
    c.cl_token = 0;                                        // the resident result is about to be overwritten
    unsigned int* d_bad = (unsigned int*)c.cl[5].p;
    CU_TRY(cudaMemsetAsync(d_bad, 0, 4, c.stream));
    uint32_t* d_cen = (uint32_t*)c.cl[0].p;
    int32_t *d_ci = (int32_t*)c.cl[1].p, *d_min = (int32_t*)c.cl[3].p, *d_max = (int32_t*)c.cl[4].p;
    uint8_t* d_lv = (uint8_t*)c.cl[2].p;
    const int nb = (int)std::min<size_t>((N + 255) / 256, (size_t)c.sms * 8);
    const int eb = (int)std::min<size_t>((E + 255) / 256, (size_t)c.sms * 16);
    const bool trace = getenv("BDG_TRACE") != nullptr;
    double tr0 = 0, tr1 = 0, tr2 = 0;
    if (trace) { cudaStreamSynchronize(st); tr0 = now_ms(); }
    CU_TRY(cudaMemcpyAsync(d_cen, centres, C * 4, cudaMemcpyHostToDevice, st));
    bdg::cluster_init_kernel<<<nb, 256, 0, st>>>(d_ci, d_lv, d_min, d_max, (uint32_t)N);
    if (C) bdg::cluster_seed_kernel<<<(int)std::min<size_t>((C + 255) / 256, (size_t)c.sms * 8), 256, 0, st>>>(d_sorted, (uint32_t)N, d_cen, (uint32_t)C, d_ci, d_lv);
    g_launches += 2;
    if (E) {
        if (!indexed)                                       // (d_sorted is c.sorted on every path that reaches here)
            if (int e = index_edges(c, N, d_ea, d_eb, E, d_ea, d_eb)) return e;
        bdg::cluster_mark_kernel<<<eb, 256, 0, st>>>(d_ea, d_eb, E, d_lv, d_bad);
        g_launches++;
        if (trace) { cudaStreamSynchronize(st); tr1 = now_ms(); }
        for (int r = 1; r <= rounds; r++) {
            bdg::cluster_claim_kernel<<<eb, 256, 0, st>>>(d_ea, d_eb, E, r, d_ci, d_lv, d_min, d_max);
            bdg::cluster_resolve_kernel<<<nb, 256, 0, st>>>(d_ci, d_lv, d_min, d_max, (uint32_t)N, r);
            g_launches += 2;
        }
    }
    CU_TRY(cudaGetLastError());
    if (trace) { cudaStreamSynchronize(st); tr2 = now_ms(); }
    if (centre_idx) CU_TRY(cudaMemcpyAsync(centre_idx, d_ci, N * 4, cudaMemcpyDeviceToHost, st));
    if (level) CU_TRY(cudaMemcpyAsync(level, d_lv, N, cudaMemcpyDeviceToHost, st));
    unsigned long long has_edge = 0;
    if (n_has_edge) {
        unsigned long long* d_cnt = (unsigned long long*)((char*)c.cl[5].p + 8);
        CU_TRY(cudaMemsetAsync(d_cnt, 0, 8, st));
        bdg::count_has_edge_kernel<<<nb, 256, 0, st>>>(d_lv, (uint32_t)N, d_cnt);
        g_launches++;
        CU_TRY(cudaMemcpyAsync(&has_edge, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    }
    unsigned int bad = 0;
    CU_TRY(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (bad) return fail(BDG_ERR_ARG, "an edge end point is not in the barcode array");
    if (n_has_edge) *n_has_edge = (size_t)has_edge;
    if (trace)
        fprintf(stderr, "[bdg] cluster N=%zu E=%zu: (gather +) seed + index %.2f ms, %d rounds %.2f ms, read-back %.2f ms\n", N, E, tr1 - tr0, rounds,
                tr2 - tr1, now_ms() - tr2);
    return BDG_OK;
}

int bdg_cluster_levels(const uint32_t* sorted_unique, size_t N, const uint32_t* ea, const uint32_t* eb, size_t E, const uint32_t* centres,
                       size_t C, int rounds, int32_t* centre_idx, uint8_t* level)
{
    if (N == 0) return BDG_OK;
    if (!sorted_unique || !centre_idx || !level || (E && (!ea || !eb)) || (C && !centres)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (N > 0x7FFFFFFFull || rounds < 0 || rounds > 253) return fail(BDG_ERR_ARG, "N must be < 2^31 and 0 <= rounds <= 253");
    if (int rc = check_sorted(sorted_unique, N)) return rc;
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    CU_TRY(cudaSetDevice(c.dev));
    auto ensure = [&](Buf& b, size_t bytes) -> int {
        if (cudaError_t e = (cudaError_t)b.ensure(bytes))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e));
        return BDG_OK;
    };
    // the edge workspaces of an earlier build are reused for the index arrays: older handles become stale
    if (int e = ensure(c.sorted, N * 4)) return e;
    if (int e = ensure(c.ea, std::max<size_t>(E, 1) * 4)) return e;
    if (int e = ensure(c.eb, std::max<size_t>(E, 1) * 4)) return e;
    c.generation++;
    CU_TRY(cudaMemcpyAsync(c.sorted.p, sorted_unique, N * 4, cudaMemcpyHostToDevice, c.stream));
    CU_TRY(cudaMemcpyAsync(c.ea.p, ea, E * 4, cudaMemcpyHostToDevice, c.stream));
    CU_TRY(cudaMemcpyAsync(c.eb.p, eb, E * 4, cudaMemcpyHostToDevice, c.stream));
    return cluster_on_device(c, (const uint32_t*)c.sorted.p, N, (uint32_t*)c.ea.p, (uint32_t*)c.eb.p, E, centres, C, rounds, centre_idx, level);
}

static int cluster_from_edges(bdg_edges* e, size_t N, const uint32_t* centres, size_t C, int rounds, int32_t* centre_idx, uint8_t* level, size_t* n_has_edge);

int bdg_cluster_levels_from_edges(bdg_edges* e, size_t N, const uint32_t* centres, size_t C, int rounds, int32_t* centre_idx, uint8_t* level)
{
    if (N && (!centre_idx || !level)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    return cluster_from_edges(e, N, centres, C, rounds, centre_idx, level, nullptr);
}

// The same with the result left on the device for bdg_assign_reads32(token, NULL, ...): no per-node array crosses PCIe.
int bdg_cluster_resident(bdg_edges* e, size_t N, const uint32_t* centres, size_t C, int rounds, size_t* n_has_edge)
{
    if (n_has_edge) *n_has_edge = 0;
    if (int rc = cluster_from_edges(e, N, centres, C, rounds, nullptr, nullptr, n_has_edge)) return rc;
    if (e && !e->ctx.empty() && e->ctx[0] == 0 && N == g_ctx[0].map_distinct) g_ctx[0].cl_token = g_ctx[0].map_token;   // the result of THIS dedup's graph
    return BDG_OK;
}

static int cluster_from_edges(bdg_edges* e, size_t N, const uint32_t* centres, size_t C, int rounds, int32_t* centre_idx, uint8_t* level, size_t* n_has_edge)
{
    if (!e) return fail(BDG_ERR_ARG, "NULL edge handle");
    if (N == 0) return BDG_OK;
    if (C && !centres) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (e->ctx.empty()) return fail(BDG_ERR_ARG, "empty edge handle");
    if (N > 0x7FFFFFFFull || rounds < 0 || rounds > 253) return fail(BDG_ERR_ARG, "N must be < 2^31 and 0 <= rounds <= 253");
    for (size_t g = 0; g < e->ctx.size(); g++)
        if (e->ctx[g] < 0 || (size_t)e->ctx[g] >= g_ctx.size() || g_ctx[e->ctx[g]].generation != e->gen[g])
            return fail(BDG_ERR_ARG, "stale edge handle: a later edge build on the same device has reused its buffers");
    DevCtx& c = g_ctx[e->ctx[0]];
    if (c.sorted.cap < N * 4) return fail(BDG_ERR_ARG, "N does not match the array the edges were built from");
    CU_TRY(cudaSetDevice(c.dev));
    if (e->ctx.size() > 1) {
        // Multi-device handle: the parts' edge lists are disjoint, so the clustering rounds need their plain
        // concatenation.  It is gathered on the first device with peer copies (NVLink; every part's stream was synchronised
        // by the build), into buffers of its own - the handle's edges stay intact on their devices.
        size_t total = 0;
        for (size_t k : e->count) total += k;
        if (cudaError_t err = (cudaError_t)c.gather_a.ensure(std::max<size_t>(total, 1) * 4))
            return fail(err == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "gather buffer of %zu edges: %s", total, cudaGetErrorString(err));
        if (cudaError_t err = (cudaError_t)c.gather_b.ensure(std::max<size_t>(total, 1) * 4))
            return fail(err == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "gather buffer of %zu edges: %s", total, cudaGetErrorString(err));
        // every device turns its own edges into node indices (its copy of the array, its SMs), then the indices travel
        size_t off = 0;
        for (size_t g = 0; g < e->ctx.size(); g++) {
            DevCtx& src = g_ctx[e->ctx[g]];
            const size_t k = e->count[g];
            if (k == 0) continue;
            if (&src == &c) {
                CU_TRY(cudaSetDevice(c.dev));
                if (int rc = index_edges(c, N, (const uint32_t*)c.ea.p, (const uint32_t*)c.eb.p, k, (uint32_t*)c.gather_a.p + off, (uint32_t*)c.gather_b.p + off)) return rc;
            } else {
                CU_TRY(cudaSetDevice(src.dev));
                if (cudaError_t err = (cudaError_t)src.gather_a.ensure(k * 4)) return fail(BDG_ERR_OOM, "index buffer: %s", cudaGetErrorString(err));
                if (cudaError_t err = (cudaError_t)src.gather_b.ensure(k * 4)) return fail(BDG_ERR_OOM, "index buffer: %s", cudaGetErrorString(err));
                if (int rc = index_edges(src, N, (const uint32_t*)src.ea.p, (const uint32_t*)src.eb.p, k, (uint32_t*)src.gather_a.p, (uint32_t*)src.gather_b.p)) return rc;
                CU_TRY(cudaMemcpyPeerAsync((uint32_t*)c.gather_a.p + off, c.dev, src.gather_a.p, src.dev, k * 4, src.stream));
                CU_TRY(cudaMemcpyPeerAsync((uint32_t*)c.gather_b.p + off, c.dev, src.gather_b.p, src.dev, k * 4, src.stream));
            }
            off += k;
        }
        for (size_t g = 0; g < e->ctx.size(); g++) {             // the rounds start when every part has arrived
            DevCtx& src = g_ctx[e->ctx[g]];
            if (&src == &c || e->count[g] == 0) continue;
            CU_TRY(cudaSetDevice(src.dev));
            CU_TRY(cudaStreamSynchronize(src.stream));
        }
        CU_TRY(cudaSetDevice(c.dev));
        return cluster_on_device(c, (const uint32_t*)c.sorted.p, N, (uint32_t*)c.gather_a.p, (uint32_t*)c.gather_b.p, total, centres, C, rounds, centre_idx, level, n_has_edge, true);
    }
    // the handle's edge VALUES are turned into node indices in place: the handle is consumed (stale afterwards)
    c.generation++;
    return cluster_on_device(c, (const uint32_t*)c.sorted.p, N, (uint32_t*)c.ea.p, (uint32_t*)c.eb.p, e->count[0], centres, C, rounds, centre_idx, level, n_has_edge);
}

int bdg_member_sorted(const uint32_t* sorted_wl, size_t W, const uint32_t* q, size_t Q, uint8_t* hit)
{
    if (Q == 0) return BDG_OK;
    if (!q || !hit || (W && !sorted_wl)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (W > 0xFFFFFFFFull || Q > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "size exceeds 2^32");
    if (int rc = need_ctx()) return rc;
    DevCtx& c = g_ctx[0];
    CU_TRY(cudaSetDevice(c.dev));
    const size_t sizes[4] = {std::max<size_t>(W, 1) * 4, Q * 4, Q, 8};       // grow-only workspaces: whitelist, queries, hits, check flag
    for (int k = 0; k < 4; k++)
        if (cudaError_t e = (cudaError_t)c.io[k].ensure(sizes[k]))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", sizes[k], cudaGetErrorString(e));
    uint32_t *d_wl = (uint32_t*)c.io[0].p, *d_q = (uint32_t*)c.io[1].p;
    uint8_t* d_hit = (uint8_t*)c.io[2].p;
    unsigned long long* d_bad = (unsigned long long*)c.io[3].p;
    CU_TRY(cudaMemcpyAsync(d_wl, sorted_wl, W * 4, cudaMemcpyHostToDevice, c.stream));
    CU_TRY(cudaMemcpyAsync(d_q, q, Q * 4, cudaMemcpyHostToDevice, c.stream));
    CU_TRY(cudaMemsetAsync(d_bad, 0xFF, 8, c.stream));
    if (W > 1) {                                            // the whitelist must be sorted: checked on the device, read back with the hits
        bdg::sorted_check_kernel<false><<<(int)std::min<size_t>((W + 255) / 256, (size_t)c.sms * 8), 256, 0, c.stream>>>(d_wl, (uint32_t)W, d_bad);
        g_launches++;
    }
    if (int rc = bdg_dev_member_sorted(d_wl, W, d_q, Q, d_hit, c.stream)) return rc;
    unsigned long long first_bad = ~0ull;
    CU_TRY(cudaMemcpyAsync(hit, d_hit, Q, cudaMemcpyDeviceToHost, c.stream));
    CU_TRY(cudaMemcpyAsync(&first_bad, d_bad, 8, cudaMemcpyDeviceToHost, c.stream));
    CU_TRY(cudaStreamSynchronize(c.stream));
    if (first_bad != ~0ull) return fail(BDG_ERR_ARG, "whitelist not sorted at index %llu", first_bad);
    return BDG_OK;
}

int bdg_nearest_bounded(const uint32_t* q, size_t Q, const uint32_t* targets, size_t W, int max_d, int32_t* argmin, uint8_t* dist)
{
    if (Q == 0) return BDG_OK;
    if (!q || !argmin || !dist || (W && !targets)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (int rc = need_ctx()) return rc;
    // queries are independent: deal contiguous slices to the devices, targets replicated
    const int G = (int)g_ctx.size();
    struct Run { uint32_t *d_q = nullptr, *d_t = nullptr, *d_keys = nullptr; int32_t* d_arg = nullptr; uint8_t* d_dist = nullptr; size_t lo = 0, n = 0; };
    std::vector<Run> runs(G);
    int rc = BDG_OK;
    const size_t per = (Q + G - 1) / G;
    for (int g = 0; g < G && rc == BDG_OK; g++) {
        Run& r = runs[g];
        r.lo = std::min(Q, per * g);
        r.n = std::min(Q, per * (g + 1)) - r.lo;
        if (r.n == 0) continue;
        DevCtx& c = g_ctx[g];
        rc = [&]() -> int {
            CU_TRY(cudaSetDevice(c.dev));
            CU_TRY(cudaMallocAsync((void**)&r.d_q, r.n * 4, c.stream));
            CU_TRY(cudaMallocAsync((void**)&r.d_t, std::max<size_t>(W, 1) * 4, c.stream));
            CU_TRY(cudaMallocAsync((void**)&r.d_keys, r.n * 4, c.stream));
            CU_TRY(cudaMallocAsync((void**)&r.d_arg, r.n * 4, c.stream));
            CU_TRY(cudaMallocAsync((void**)&r.d_dist, r.n, c.stream));
            CU_TRY(cudaMemcpyAsync(r.d_q, q + r.lo, r.n * 4, cudaMemcpyHostToDevice, c.stream));
            CU_TRY(cudaMemcpyAsync(r.d_t, targets, W * 4, cudaMemcpyHostToDevice, c.stream));
            if (int e = bdg_dev_nearest_bounded(r.d_q, r.n, r.d_t, W, max_d, r.d_keys, r.d_arg, r.d_dist, c.stream)) return e;
            CU_TRY(cudaMemcpyAsync(argmin + r.lo, r.d_arg, r.n * 4, cudaMemcpyDeviceToHost, c.stream));
            CU_TRY(cudaMemcpyAsync(dist + r.lo, r.d_dist, r.n, cudaMemcpyDeviceToHost, c.stream));
            return BDG_OK;
        }();
    }
    for (int g = 0; g < G; g++) {
        Run& r = runs[g];
        if (r.n == 0) continue;
        DevCtx& c = g_ctx[g];
        cudaSetDevice(c.dev);
        cudaFreeAsync(r.d_q, c.stream); cudaFreeAsync(r.d_t, c.stream); cudaFreeAsync(r.d_keys, c.stream);
        cudaFreeAsync(r.d_arg, c.stream); cudaFreeAsync(r.d_dist, c.stream);
        cudaError_t e = cudaStreamSynchronize(c.stream);
        if (e != cudaSuccess && rc == BDG_OK) rc = fail(BDG_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
    }
    cudaSetDevice(g_ctx[0].dev);
    return rc;
}

// ---- a-5  KmerIndexer / QGramIndex: the known strings stay on the device between queries ------------------
static int kmer_index_create(const uint32_t* wl, size_t W, bool postings, bdg_kmer_index** out)
{
    if (!out || (W && !wl)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    if (W > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "size exceeds 2^32");
    if (int rc = need_ctx()) return rc;
    bdg_kmer_index* ix = new (std::nothrow) bdg_kmer_index();
    if (!ix) return fail(BDG_ERR_OOM, "host allocation failed");
    ix->dev = g_ctx[0].dev;
    ix->W = W;
    CU_TRY(cudaSetDevice(ix->dev));
    if (cudaError_t e = (cudaError_t)ix->wl.ensure(std::max<size_t>(W, 1) * 4)) {
        delete ix;
        return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", W * 4, cudaGetErrorString(e));
    }
    cudaError_t e = cudaMemcpyAsync(ix->wl.p, wl, W * 4, cudaMemcpyHostToDevice, g_ctx[0].stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx[0].stream);
    if (e != cudaSuccess) { bdg_kmer_index_free(ix); return fail(BDG_ERR_CUDA, "upload of the index failed: %s", cudaGetErrorString(e)); }
    // posting lists of the strings' 6-mers (kmer_indexer.py:29-32): worth it once a bucket is a small share of the strings
    size_t min_w = 4096;
    if (const char* v = getenv("BDG_KMER_POST_MIN_W")) min_w = (size_t)std::max(0ll, atoll(v));
    if (postings && W >= min_w && W > 0 && W <= 0xFFFFFFFFull / 11) {          // (the postings are counted in 32 bits)
        cudaStream_t st = g_ctx[0].stream;
        const int sms = g_ctx[0].sms;
        size_t tmp = 0;
        bool ok = ix->kstart.ensure(4097 * 4) == BDG_OK && ix->scratch.ensure(2 * 4097 * 4) == BDG_OK;
        ok = ok && cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, 4097, st) == cudaSuccess;
        Buf cubtmp;
        ok = ok && cubtmp.ensure(tmp) == BDG_OK;
        if (ok) {
            uint32_t* hist = (uint32_t*)ix->scratch.p;
            uint32_t* fill = hist + 4097;
            const int nb = (int)std::min<size_t>((W + 255) / 256, (size_t)sms * 8);
            ok = cudaMemsetAsync(hist, 0, 2 * 4097 * 4, st) == cudaSuccess;
            bdg::kidx_hist_kernel<<<nb, 256, 0, st>>>((const uint32_t*)ix->wl.p, (uint32_t)W, hist);
            ok = ok && cub::DeviceScan::ExclusiveSum(cubtmp.p, tmp, (const uint32_t*)hist, (uint32_t*)ix->kstart.p, 4097, st) == cudaSuccess;
            uint32_t n_post = 0;
            ok = ok && cudaMemcpyAsync(&n_post, (uint32_t*)ix->kstart.p + 4096, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
            ok = ok && ix->post.ensure(std::max<size_t>(n_post, 1) * 4) == BDG_OK;
            if (ok) {
                bdg::kidx_scatter_kernel<<<nb, 256, 0, st>>>((const uint32_t*)ix->wl.p, (uint32_t)W, (const uint32_t*)ix->kstart.p, fill, (uint32_t*)ix->post.p);
                g_launches += 2;
                ok = cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess;
            }
        }
        cubtmp.release();
        if (!ok) { bdg_kmer_index_free(ix); return fail(BDG_ERR_CUDA, "building the 6-mer posting lists failed: %s", cudaGetErrorString(cudaGetLastError())); }
        ix->posted = true;
    }
    if (cudaEventCreate(&ix->ev[0]) != cudaSuccess || cudaEventCreate(&ix->ev[1]) != cudaSuccess) {
        bdg_kmer_index_free(ix);
        return fail(BDG_ERR_CUDA, "event creation failed");
    }
    *out = ix;
    return BDG_OK;
}

int bdg_kmer_index_create(const uint32_t* wl, size_t W, bdg_kmer_index** out) { return kmer_index_create(wl, W, true, out); }

int bdg_kmer_index_info(const bdg_kmer_index* ix, int* posted, double* kernel_ms)
{
    if (!ix) return fail(BDG_ERR_ARG, "NULL index handle");
    if (posted) *posted = ix->posted ? 1 : 0;
    if (kernel_ms) {
        float ms = 0.f;
        if (ix->timed) CU_TRY(cudaEventElapsedTime(&ms, ix->ev[0], ix->ev[1]));
        *kernel_ms = ms;
    }
    return BDG_OK;
}

void bdg_kmer_index_free(bdg_kmer_index* ix)
{
    if (!ix) return;
    cudaSetDevice(ix->dev);
    ix->wl.release(); ix->kstart.release(); ix->post.release(); ix->scratch.release();
    for (auto& b : ix->outb) b.release();
    for (auto& e : ix->ev) if (e) cudaEventDestroy(e);
    delete ix;
}

int bdg_kmer_index_query(bdg_kmer_index* ix, const uint32_t* q, size_t Q, int min_kmers, size_t cap, uint32_t* hit_q, uint32_t* hit_w,
                         uint8_t* cnt, uint64_t* mult, size_t* total)
{
    if (!total) return fail(BDG_ERR_ARG, "NULL total pointer");
    *total = 0;
    if (!ix) return fail(BDG_ERR_ARG, "NULL index handle");
    const size_t W = ix->W;
    if (Q == 0 || W == 0) return BDG_OK;
    if (!q || (cap && (!hit_q || !hit_w || !cnt || !mult))) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (Q > 0xFFFFFFFFull) return fail(BDG_ERR_ARG, "size exceeds 2^32");
    DevCtx* c = nullptr;
    for (auto& x : g_ctx) if (x.dev == ix->dev) c = &x;
    if (!c) return fail(BDG_ERR_NODEVICE, "the device the index lives on is no longer claimed (bdg_shutdown?)");
    CU_TRY(cudaSetDevice(c->dev));
    const size_t capa = std::max<size_t>(cap, 1);
    const size_t sizes[6] = {Q * 4, capa * 4, capa * 4, capa, capa * 8, 8};     // queries, hit_q, hit_w, cnt, mult, total
    for (int k = 0; k < 6; k++)
        if (cudaError_t e = (cudaError_t)ix->outb[k].ensure(sizes[k]))
            return fail(e == cudaErrorMemoryAllocation ? BDG_ERR_OOM : BDG_ERR_CUDA, "device allocation of %zu bytes: %s", sizes[k], cudaGetErrorString(e));
    uint32_t *d_q = (uint32_t*)ix->outb[0].p, *d_hq = (uint32_t*)ix->outb[1].p, *d_hw = (uint32_t*)ix->outb[2].p;
    uint8_t* d_cnt = (uint8_t*)ix->outb[3].p;
    unsigned long long *d_mult = (unsigned long long*)ix->outb[4].p, *d_total = (unsigned long long*)ix->outb[5].p;
    CU_TRY(cudaMemsetAsync(d_total, 0, 8, c->stream));
    CU_TRY(cudaMemcpyAsync(d_q, q, Q * 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaEventRecord(ix->ev[0], c->stream));
    if (ix->posted) {                                       // walk the buckets of the queries' own 6-mers
        const unsigned blocks = (unsigned)std::min<unsigned long long>((unsigned long long)Q * 11, (unsigned long long)c->sms * 64);
        bdg::kmer_post_kernel<<<blocks, bdg::NT, 0, c->stream>>>(d_q, (uint32_t)Q, (const uint32_t*)ix->wl.p, (const uint32_t*)ix->kstart.p,
                                                               (const uint32_t*)ix->post.p, std::max(min_kmers, 1), (unsigned long long)cap, d_hq, d_hw,
                                                               d_cnt, d_mult, d_total);
    } else {
        const uint32_t gy_total = (uint32_t)((Q + bdg::KS_QB - 1) / bdg::KS_QB);
        const uint32_t gx = (uint32_t)((W + bdg::NT - 1) / bdg::NT);
        if (gy_total > 65535u) return fail(BDG_ERR_ARG, "Q too large for one call (> 65535*256 queries)");
        bdg::kmer_score_kernel<<<dim3(gx, gy_total), bdg::NT, 0, c->stream>>>(d_q, (uint32_t)Q, (const uint32_t*)ix->wl.p, (uint32_t)W, min_kmers,
                                                                           (unsigned long long)cap, d_hq, d_hw, d_cnt, d_mult, d_total);
    }
    g_launches++;
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(ix->ev[1], c->stream));
    ix->timed = true;
    unsigned long long tot = 0;
    CU_TRY(cudaMemcpyAsync(&tot, d_total, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)std::min<unsigned long long>(tot, cap);
    if (n) {
        CU_TRY(cudaMemcpyAsync(hit_q, d_hq, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaMemcpyAsync(hit_w, d_hw, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaMemcpyAsync(cnt, d_cnt, n, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaMemcpyAsync(mult, d_mult, n * 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(cudaStreamSynchronize(c->stream));
    }
    *total = (size_t)tot;
    if (tot > cap) return fail(BDG_ERR_CAPACITY, "%llu hits but room for %zu", tot, cap);
    return BDG_OK;
}

int bdg_kmer_score(const uint32_t* q, size_t Q, const uint32_t* wl, size_t W, int min_kmers, size_t cap, uint32_t* hit_q,
                   uint32_t* hit_w, uint8_t* cnt, uint64_t* mult, size_t* total)
{
    if (!total) return fail(BDG_ERR_ARG, "NULL total pointer");
    *total = 0;
    if (Q == 0 || W == 0) return BDG_OK;
    if (!q || !wl || (cap && (!hit_q || !hit_w || !cnt || !mult))) return fail(BDG_ERR_ARG, "NULL pointer argument");
    bdg_kmer_index* ix = nullptr;
    if (int rc = kmer_index_create(wl, W, Q >= 64, &ix)) return rc;       // a one-off call with few queries: the scan costs less than the lists
    const int rc = bdg_kmer_index_query(ix, q, Q, min_kmers, cap, hit_q, hit_w, cnt, mult, total);
    bdg_kmer_index_free(ix);
    return rc;
}

// ---- f-2 / f-4  extraction TSV in, assignment TSV out, whitelist file (host only; bdg_tsv.hpp) -----------------
struct bdg_tsv { tsvio::Tsv t; };
struct bdg_lines16 { tsvio::Lines16 l; };

int bdg_tsv_open(const char* path, int bc_len, int threads, bdg_tsv** out)
{
    if (!path || !out) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    bdg_tsv* h = new (std::nothrow) bdg_tsv();
    if (!h) return fail(BDG_ERR_OOM, "out of host memory");
    bool io = false;
    std::string why;
    try {
        why = tsvio::tsv_parse(h->t, path, bc_len, threads, &io);
    } catch (const std::bad_alloc&) {
        delete h;
        return fail(BDG_ERR_OOM, "out of host memory while reading %s", path);
    }
    if (!why.empty()) {
        delete h;
        return fail(io ? BDG_ERR_IO : BDG_ERR_UNSUPPORTED, "%s%s", io ? "" : "native TSV reader refuses the file: ", why.c_str());
    }
    *out = h;
    return BDG_OK;
}

size_t bdg_tsv_rows(const bdg_tsv* t) { return t ? t->t.rows : 0; }

int bdg_tsv_barcodes(const bdg_tsv* t, char* seqs16, uint8_t* kind)
{
    if (!t) return fail(BDG_ERR_ARG, "NULL TSV handle");
    if (t->t.rows && (!seqs16 || !kind)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (t->t.rows) {
        memcpy(seqs16, t->t.seqs.data(), t->t.rows * 16);
        memcpy(kind, t->t.kind.data(), t->t.rows);
    }
    return BDG_OK;
}

int bdg_tsv_write_assignments(const bdg_tsv* t, const char* out_path, const uint64_t* centre_per_row, int threads)
{
    if (!t || !out_path) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (t->t.rows && !centre_per_row) return fail(BDG_ERR_ARG, "NULL pointer argument");
    std::string why;
    try {
        why = tsvio::tsv_write(t->t, out_path, centre_per_row, threads);
    } catch (const std::bad_alloc&) {
        return fail(BDG_ERR_OOM, "out of host memory while writing %s", out_path);
    }
    if (!why.empty()) return fail(BDG_ERR_IO, "%s", why.c_str());
    return BDG_OK;
}

int bdg_tsv_write_assignments32(const bdg_tsv* t, const char* out_path, const uint32_t* centre_per_row, const uint8_t* has_centre, int threads)
{
    if (!t || !out_path) return fail(BDG_ERR_ARG, "NULL pointer argument");
    if (t->t.rows && (!centre_per_row || !has_centre)) return fail(BDG_ERR_ARG, "NULL pointer argument");
    std::string why;
    try {
        why = tsvio::tsv_write(t->t, out_path, nullptr, threads, centre_per_row, has_centre);
    } catch (const std::bad_alloc&) {
        return fail(BDG_ERR_OOM, "out of host memory while writing %s", out_path);
    }
    if (!why.empty()) return fail(BDG_ERR_IO, "%s", why.c_str());
    return BDG_OK;
}

void bdg_tsv_close(bdg_tsv* t) { delete t; }

int bdg_lines16_open(const char* path, bdg_lines16** out)
{
    if (!path || !out) return fail(BDG_ERR_ARG, "NULL pointer argument");
    *out = nullptr;
    bdg_lines16* h = new (std::nothrow) bdg_lines16();
    if (!h) return fail(BDG_ERR_OOM, "out of host memory");
    bool io = false;
    std::string why;
    try {
        why = tsvio::lines16_read(h->l, path, &io);
    } catch (const std::bad_alloc&) {
        delete h;
        return fail(BDG_ERR_OOM, "out of host memory while reading %s", path);
    }
    if (!why.empty()) {
        delete h;
        return fail(BDG_ERR_IO, "%s", why.c_str());
    }
    *out = h;
    return BDG_OK;
}

size_t bdg_lines16_count(const bdg_lines16* l) { return l ? l->l.count : 0; }
const char* bdg_lines16_data(const bdg_lines16* l) { return l ? l->l.seqs.data() : nullptr; }
void bdg_lines16_close(bdg_lines16* l) { delete l; }

}  // extern "C"
