// Host build of badger_b200/csrc/bdg_core.cuh for the CPU-side unit tests (tests/test_core_host.py).
// Test infrastructure: lets the per-pair device arithmetic be checked against the oracle without a GPU.
#include "../badger_b200/csrc/bdg_core.cuh"
#include <stddef.h>

extern "C" {
void shim_pairs(const uint32_t* a, const uint32_t* b, size_t n, uint8_t* pre1, uint8_t* pre2, uint8_t* dsmall,
                uint8_t* dplain, uint8_t* dfull, uint8_t* da15, uint8_t* db15, uint8_t* S, uint64_t* mult)
{
    for (size_t i = 0; i < n; i++) {
        pre1[i] = bdg::prefilter_t1(a[i], b[i]);
        pre2[i] = bdg::prefilter_t2(a[i], b[i]);
        dsmall[i] = (uint8_t)bdg::dist_small(a[i], b[i], false);
        dplain[i] = (uint8_t)bdg::dist_small(a[i], b[i], true);
        bdg::Dist3 r = bdg::myers3(a[i], b[i]);
        dfull[i] = (uint8_t)r.full; da15[i] = (uint8_t)r.a15; db15[i] = (uint8_t)r.b15;
        S[i] = (uint8_t)bdg::qgram_score(a[i], b[i], &mult[i]);
    }
}
int shim_edge(uint32_t a, uint32_t b, int t) { return bdg::edge_dist(a, b, t); }
// split prefilters of the tiled edge kernel: light (per pair) and top (excluded per tile)
void shim_split(const uint32_t* a, const uint32_t* b, size_t n, uint8_t* l1, uint8_t* t1, uint8_t* l2, uint8_t* t2)
{
    for (size_t i = 0; i < n; i++) {
        l1[i] = bdg::t1_light(a[i], b[i]); t1[i] = bdg::t1_top(a[i], b[i]);
        l2[i] = bdg::t2_light(a[i], b[i]); t2[i] = bdg::t2_top(a[i], b[i]);
    }
}
void shim_pass_pred(int t, const uint32_t* a, const uint32_t* b, size_t n, uint8_t* mask /* bit p = pass p fires */)
{
    for (size_t i = 0; i < n; i++) {
        uint8_t m = 0;
        for (int p = 0; p < bdg::n_passes(t); p++) {
            const bool f = bdg::pass_pred(t, p, a[i], b[i]);
            const int r = bdg::pass_rot(t, p);
            if (f != bdg::pass_pred_rot(t, p, bdg::rotl32(a[i], r), bdg::rotl32(b[i], r))) m |= 0x80;   // must agree
            if (f) m |= (uint8_t)(1u << p);
        }
        mask[i] = m;
    }
}
int shim_pass_possible(int t, int p, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi)
{
    return bdg::pass_possible(t, p, alo, ahi, blo, bhi);
}
void shim_quick(int t, const uint32_t* a, const uint32_t* b, size_t n, uint8_t* out)
{
    for (size_t i = 0; i < n; i++) out[i] = bdg::quick_pass(a[i], b[i], t);
}
void shim_quick_any(int t, const uint32_t* a, const uint32_t* b, size_t n, uint8_t* out)
{
    for (size_t i = 0; i < n; i++) out[i] = bdg::quick_pass_any(a[i], b[i], t);
}
// two-block seeds (t = 2): first condition met by (a, b), -1 if none; and the two join keys of condition c
void shim_seed2_first(const uint32_t* a, const uint32_t* b, size_t n, int8_t* out)
{
    for (size_t i = 0; i < n; i++) out[i] = (int8_t)bdg::seed2_first(a[i], b[i]);
}
void shim_seed2_keys(int c, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* ka, uint32_t* kb)
{
    for (size_t i = 0; i < n; i++) { ka[i] = bdg::seed2_key_a(c, a[i]); kb[i] = bdg::seed2_key_b(c, b[i]); }
}
int shim_seed2_count(void) { return bdg::SEED2_N; }
void shim_seed2_permute(int c, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* pa, uint32_t* pb, uint32_t* ua, uint32_t* ub)
{
    const bdg::SeedPerm A = bdg::seed2_perm_a(c), B = bdg::seed2_perm_b(c);
    for (size_t i = 0; i < n; i++) {
        pa[i] = bdg::seed_permute(a[i], A); pb[i] = bdg::seed_permute(b[i], B);
        ua[i] = bdg::seed_unpermute(pa[i], A); ub[i] = bdg::seed_unpermute(pb[i], B);
    }
}
int shim_seed2_key_bits(int c) { return bdg::seed_key_bits(bdg::seed2_perm_a(c)); }
int shim_seed_tiles_meet(uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, int kb) { return bdg::seed_tiles_meet(alo, ahi, blo, bhi, kb); }
int shim_top_possible(int t, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi)
{
    return t == 1 ? bdg::t1_top_possible(alo, ahi, blo, bhi) : bdg::t2_top_possible(alo, ahi, blo, bhi);
}
}
