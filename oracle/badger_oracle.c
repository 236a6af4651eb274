/*
 * badger_oracle.c -- CPU restatement of algbio/Badger's barcode edit-distance hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (badger_b200/, libbadger_b200.so) never does and has no CPU fallback.
 *
 * Parity is PINNED: every function below is checked in tests/test_oracle_golden.py against
 * fixtures produced by running the unmodified reference in the authoring container
 * (oracle/make_golden.py -> tests/golden/), and, when /root/reference is present, against
 * the live reference (tests/test_oracle_vs_reference.py).
 *
 * Each function cites the reference file:line it restates.  The reference is Python; the
 * algorithms are restated here in plain C over the reference's own 2-bit packing.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BC 16
#define Q 6
#define NK (BC - Q + 1) /* 11 six-mers per barcode */

/* ---- common.py:11-25  rank(): r = sum code(s[i]) * 4^i, A0 C1 G2 T3; other letters -> KeyError ---- */
int orc_rank16(const char *s, uint32_t *out)
{
    uint32_t r = 0;
    for (int i = 0; i < BC; i++) {
        uint32_t c;
        switch (s[i]) {
        case 'A': c = 0; break;
        case 'C': c = 1; break;
        case 'G': c = 2; break;
        case 'T': c = 3; break;
        default: return -1; /* KeyError in the reference (common.py:24) */
        }
        r += c << (2 * i);
    }
    *out = r;
    return 0;
}

/* ---- common.py:27-38  unrank() ---- */
void orc_unrank16(uint32_t r, char *out)
{
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int i = 0; i < BC; i++) {
        out[i] = L[r % 4];
        r /= 4;
    }
}

/* ---- editdistance.eval (third-party, unpinned; unit-cost Levenshtein by definition) ----
 * Plain two-row dynamic programme over the first la / lb bases of the packed words. */
int orc_ed(uint32_t a, int la, uint32_t b, int lb)
{
    int prev[BC + 1], cur[BC + 1];
    for (int j = 0; j <= lb; j++) prev[j] = j;
    for (int i = 1; i <= la; i++) {
        uint32_t ca = (a >> (2 * (i - 1))) & 3u;
        cur[0] = i;
        for (int j = 1; j <= lb; j++) {
            uint32_t cb = (b >> (2 * (j - 1))) & 3u;
            int best = prev[j - 1] + (ca != cb);
            if (prev[j] + 1 < best) best = prev[j] + 1;
            if (cur[j - 1] + 1 < best) best = cur[j - 1] + 1;
            cur[j] = best;
        }
        memcpy(prev, cur, sizeof(int) * (lb + 1));
    }
    return prev[lb];
}

/* ---- barcode_graph.py:96 / :243  dist = min(ed(a,b), ed(a[:-1],b), ed(a,b[:-1])) ---- */
int orc_D(uint32_t a, uint32_t b)
{
    int d = orc_ed(a, BC, b, BC);
    int d1 = orc_ed(a, BC - 1, b, BC);
    int d2 = orc_ed(a, BC, b, BC - 1);
    if (d1 < d) d = d1;
    if (d2 < d) d = d2;
    return d;
}

/* ---- index.py:19-24  threshold = bc_len - q + 1 - q*t, replaced by 4 when <= 0 ---- */
int orc_T(int t)
{
    int T = BC - Q + 1 - Q * t;
    return T <= 0 ? 4 : T;
}

/* 6-mer at position p of a packed barcode == QGramIndex.rank/update_rank (index.py:68-75) */
static inline uint32_t kmer_at(uint32_t r, int p) { return (r >> (2 * p)) & 0xFFFu; }

/* ---- index.py:29-35 + :77-93  S(a,b) = sum over the query's 11 six-mers (repeats included) of the
 *      multiplicity of that six-mer in b  ==  #{(p,q): kmer_a[p] == kmer_b[q]} ---- */
int orc_S(uint32_t a, uint32_t b)
{
    int s = 0;
    for (int p = 0; p < NK; p++)
        for (int q = 0; q < NK; q++)
            s += kmer_at(a, p) == kmer_at(b, q);
    return s;
}

/* Edge predicate of barcode_graph.py:224-249: candidate (S >= T, index.py:91-92) AND D <= t
 * (barcode_graph.py:245).  Returns D if (a,b) is an edge, else -1.  Requires a != b. */
int orc_edge(uint32_t a, uint32_t b, int t)
{
    if (a == b) return -1; /* barcode_graph.py:239 */
    if (orc_S(a, b) < orc_T(t)) return -1;
    int d = orc_D(a, b);
    return d <= t ? d : -1;
}

/* Brute force over all unordered pairs: the predicate applied without any index.
 * out_* may be NULL (count only).  Returns the number of edges (may exceed cap). */
int64_t orc_edges_brute(const uint32_t *r, size_t n, int t, uint32_t *oa, uint32_t *ob, uint8_t *od, size_t cap)
{
    int64_t cnt = 0;
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {
            if (r[i] >= r[j]) continue; /* index.py:87 "j > number" */
            int d = orc_edge(r[i], r[j], t);
            if (d < 0) continue;
            if (oa && (size_t)cnt < cap) { oa[cnt] = r[i]; ob[cnt] = r[j]; od[cnt] = (uint8_t)d; }
            cnt++;
        }
    return cnt;
}

/* ---------------------------------------------------------------------------------------------
 * The reference's own algorithm: 4096 buckets {rank -> multiplicity} (index.py:29-41), get_close
 * walking the query's 11 buckets and summing multiplicities of larger ranks (index.py:77-93), then the
 * 3-way edit-distance verify (barcode_graph.py:233-249).  This is what bench.py times as the CPU
 * baseline ("port"), threaded over query rows the way barcode_graph.py:164-189 spreads chunks over
 * processes.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    size_t n;
    const uint32_t *ranks;
    uint64_t *start; /* 4097 bucket offsets */
    uint32_t *ids;   /* entry -> index into ranks */
    uint8_t *mult;   /* entry -> multiplicity of the six-mer inside that barcode */
} orc_index;

orc_index *orc_index_build(const uint32_t *ranks, size_t n)
{
    orc_index *ix = (orc_index *)calloc(1, sizeof(*ix));
    ix->n = n;
    ix->ranks = ranks;
    ix->start = (uint64_t *)calloc(4097, sizeof(uint64_t));
    /* pass 1: count DISTINCT six-mers per barcode per bucket (add_to_index increments the same dict slot) */
    for (size_t i = 0; i < n; i++) {
        for (int p = 0; p < NK; p++) {
            uint32_t k = kmer_at(ranks[i], p);
            int seen = 0;
            for (int q = 0; q < p; q++) seen |= kmer_at(ranks[i], q) == k;
            if (!seen) ix->start[k + 1]++;
        }
    }
    for (int k = 0; k < 4096; k++) ix->start[k + 1] += ix->start[k];
    uint64_t total = ix->start[4096];
    ix->ids = (uint32_t *)malloc(sizeof(uint32_t) * (total ? total : 1));
    ix->mult = (uint8_t *)malloc(total ? total : 1);
    uint64_t *fill = (uint64_t *)malloc(sizeof(uint64_t) * 4096);
    memcpy(fill, ix->start, sizeof(uint64_t) * 4096);
    for (size_t i = 0; i < n; i++) {
        for (int p = 0; p < NK; p++) {
            uint32_t k = kmer_at(ranks[i], p);
            int seen = 0, m = 0;
            for (int q = 0; q < p; q++) seen |= kmer_at(ranks[i], q) == k;
            if (seen) continue;
            for (int q = p; q < NK; q++) m += kmer_at(ranks[i], q) == k;
            ix->ids[fill[k]] = (uint32_t)i;
            ix->mult[fill[k]] = (uint8_t)m;
            fill[k]++;
        }
    }
    free(fill);
    return ix;
}

void orc_index_free(orc_index *ix)
{
    if (!ix) return;
    free(ix->start); free(ix->ids); free(ix->mult); free(ix);
}

/* get_close for one query row (index.py:77-93).  acc is a zeroed n-byte scratch, touched an n-entry
 * scratch; returns the number of candidates written to cand (indices into ranks). */
static size_t get_close_row_dir(const orc_index *ix, size_t row, int T, uint8_t *acc, uint32_t *touched, uint32_t *cand, int smaller);
static size_t get_close_row(const orc_index *ix, size_t row, int T, uint8_t *acc, uint32_t *touched, uint32_t *cand)
{
    return get_close_row_dir(ix, row, T, acc, touched, cand, 0);
}
/* smaller = 0: index.py:77-93 as written (candidates with rank > number).  smaller = 1: the candidates with rank < number,
 * i.e. the queries whose own get_close returns this row - S is symmetric, so walking the row's buckets finds them. */
static size_t get_close_row_dir(const orc_index *ix, size_t row, int T, uint8_t *acc, uint32_t *touched, uint32_t *cand, int smaller)
{
    uint32_t number = ix->ranks[row];
    size_t nt = 0, nc = 0;
    for (int p = 0; p < NK; p++) { /* the query's six-mers, repeats walked again (index.py:85) */
        uint32_t k = kmer_at(number, p);
        for (uint64_t e = ix->start[k]; e < ix->start[k + 1]; e++) {
            uint32_t j = ix->ids[e];
            if (smaller ? ix->ranks[j] < number : ix->ranks[j] > number) { /* index.py:87 */
                if (!acc[j]) touched[nt++] = j;
                acc[j] += ix->mult[e];
            }
        }
    }
    for (size_t x = 0; x < nt; x++) {
        uint32_t j = touched[x];
        if (acc[j] >= T) cand[nc++] = j; /* index.py:91-92 */
        acc[j] = 0;
    }
    return nc;
}

/* Candidate ranks of one query (QGramIndex.get_close); out must hold n entries. */
size_t orc_get_close(const orc_index *ix, size_t row, int t, uint32_t *out_ranks)
{
    uint8_t *acc = (uint8_t *)calloc(ix->n ? ix->n : 1, 1);
    uint32_t *touched = (uint32_t *)malloc(sizeof(uint32_t) * (ix->n ? ix->n : 1));
    uint32_t *cand = (uint32_t *)malloc(sizeof(uint32_t) * (ix->n ? ix->n : 1));
    size_t nc = get_close_row(ix, row, orc_T(t), acc, touched, cand);
    for (size_t i = 0; i < nc; i++) out_ranks[i] = ix->ranks[cand[i]];
    free(acc); free(touched); free(cand);
    return nc;
}

/* graph_construction over `rows` (NULL = every row).  Edges appended (a<b) up to cap; returns the edge
 * count; *verified receives the number of candidates that went through the 3-way verify. */
int64_t orc_edges_index(const orc_index *ix, int t, const uint32_t *rows, size_t nrows, int threads,
                        uint32_t *oa, uint32_t *ob, uint8_t *od, size_t cap, uint64_t *verified)
{
    size_t n = ix->n;
    if (!rows) nrows = n;
    int T = orc_T(t);
    int64_t cnt = 0;
    uint64_t ver = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        uint8_t *acc = (uint8_t *)calloc(n ? n : 1, 1);
        uint32_t *touched = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
        uint32_t *cand = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
#pragma omp for schedule(dynamic, 64) reduction(+ : ver)
        for (size_t x = 0; x < nrows; x++) {
            size_t row = rows ? rows[x] : x;
            uint32_t a = ix->ranks[row];
            size_t nc = get_close_row(ix, row, T, acc, touched, cand);
            ver += nc;
            for (size_t c = 0; c < nc; c++) {
                uint32_t b = ix->ranks[cand[c]];
                if (b == a) continue; /* barcode_graph.py:239 */
                int d = orc_D(a, b);  /* barcode_graph.py:243 */
                if (d <= t) {          /* barcode_graph.py:245 */
                    int64_t slot;
#pragma omp atomic capture
                    slot = cnt++;
                    if (oa && (size_t)slot < cap) { oa[slot] = a; ob[slot] = b; od[slot] = (uint8_t)d; }
                }
            }
        }
        free(acc); free(touched); free(cand);
    }
    if (verified) *verified = ver;
    return cnt;
}

/* Every edge that has one of `rows` as an END POINT (smaller or larger barcode): what graph_construction leaves in
 * edges[rows[x]] (barcode_graph.py:245-249 appends both directions).  Used for sampled-row parity at sizes where the whole
 * index walk is out of reach.  An edge between two sampled rows is reported from both of them. */
int64_t orc_edges_rows_both(const orc_index *ix, int t, const uint32_t *rows, size_t nrows, int threads,
                            uint32_t *oa, uint32_t *ob, uint8_t *od, size_t cap)
{
    size_t n = ix->n;
    int T = orc_T(t);
    int64_t cnt = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        uint8_t *acc = (uint8_t *)calloc(n ? n : 1, 1);
        uint32_t *touched = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
        uint32_t *cand = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
#pragma omp for schedule(dynamic, 16)
        for (size_t x = 0; x < nrows; x++) {
            uint32_t a = ix->ranks[rows[x]];
            for (int dir = 0; dir < 2; dir++) {
                size_t nc = get_close_row_dir(ix, rows[x], T, acc, touched, cand, dir);
                for (size_t c = 0; c < nc; c++) {
                    uint32_t b = ix->ranks[cand[c]];
                    if (b == a) continue;
                    int d = orc_D(a, b);
                    if (d <= t) {
                        int64_t slot;
#pragma omp atomic capture
                        slot = cnt++;
                        if (oa && (size_t)slot < cap) { oa[slot] = a < b ? a : b; ob[slot] = a < b ? b : a; od[slot] = (uint8_t)d; }
                    }
                }
            }
        }
        free(acc); free(touched); free(cand);
    }
    return cnt;
}

/* ---- barcode_graph.py:192-204  dedup + count in first-seen order over reads that survived the
 *      length rules (the caller has already applied :195-197; valid[i]==0 rows are skipped).
 *      Open-addressing hash; out_rank/out_count sized for n.  Returns the number of distinct. ---- */
size_t orc_dedup_count(const uint32_t *reads, const uint8_t *valid, size_t n, uint32_t *out_rank, uint32_t *out_count)
{
    size_t cap = 16;
    while (cap < 2 * n + 16) cap <<= 1;
    int64_t *slot = (int64_t *)malloc(sizeof(int64_t) * cap);
    for (size_t i = 0; i < cap; i++) slot[i] = -1;
    size_t nd = 0;
    for (size_t i = 0; i < n; i++) {
        if (valid && !valid[i]) continue;
        uint32_t r = reads[i];
        size_t h = ((uint64_t)r * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
        for (;;) {
            if (slot[h] < 0) { slot[h] = (int64_t)nd; out_rank[nd] = r; out_count[nd] = 1; nd++; break; }
            if (out_rank[slot[h]] == r) { out_count[slot[h]]++; break; }
            h = (h + 1) & (cap - 1);
        }
    }
    free(slot);
    return nd;
}

/* ---- barcode_graph.py:262-267  `unrank(r) in barcode_list`: exact set membership ---- */
void orc_member(const uint32_t *sorted_wl, size_t W, const uint32_t *q, size_t nq, uint8_t *hit)
{
    for (size_t i = 0; i < nq; i++) {
        size_t lo = 0, hi = W;
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (sorted_wl[mid] < q[i]) lo = mid + 1; else hi = mid;
        }
        hit[i] = (lo < W && sorted_wl[lo] == q[i]);
    }
}

/* ---- barcode_graph.py:370-385  postprocessing: first strict minimum of the PLAIN edit distance over
 *      the centres in iteration order, starting from min_dist = 16; accepted when min_dist < 3.
 *      argmin = -1 / dist = 255 when rejected.  max_d = 2 reproduces `< 3`. ---- */
void orc_nearest(const uint32_t *q, size_t nq, const uint32_t *targets, size_t W, int max_d, int32_t *argmin, uint8_t *dist)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < nq; i++) {
        int best = 16, bi = -1;
        for (size_t j = 0; j < W; j++) {
            int d = orc_ed(q[i], BC, targets[j], BC);
            if (d < best) { best = d; bi = (int)j; }
        }
        if (best <= max_d && bi >= 0) { argmin[i] = bi; dist[i] = (uint8_t)best; }
        else { argmin[i] = -1; dist[i] = 255; }
    }
}

/* ---- kmer_indexer.py:49-61 for packed 16-mers, k = 6: cnt = #{(p,q'): kmer_q[p]==kmer_B[q']} and
 *      mult[p] = number of target positions matching query position p (so that the reference's
 *      `positions` list is p repeated mult[p] times, ascending).  Dense Q x W output. ---- */
void orc_kmer_score(const uint32_t *q, size_t nq, const uint32_t *wl, size_t W, uint8_t *cnt /*nq*W*/, uint8_t *mult /*nq*W*11 or NULL*/)
{
    for (size_t i = 0; i < nq; i++)
        for (size_t j = 0; j < W; j++) {
            int s = 0;
            for (int p = 0; p < NK; p++) {
                int m = 0;
                for (int x = 0; x < NK; x++) m += kmer_at(q[i], p) == kmer_at(wl[j], x);
                if (mult) mult[(i * W + j) * NK + p] = (uint8_t)m;
                s += m;
            }
            cnt[i * W + j] = (uint8_t)s;
        }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
