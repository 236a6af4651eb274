"""Per-pair device arithmetic (badger_b200/csrc/bdg_core.cuh), compiled for the host, against the oracle.

The same header is what the CUDA kernels include; this test is the CPU-side proof that the prefilters
are sound (never reject a pair with D <= t) and that the exact stage equals the reference's distances.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from badger_b200 import synth
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "core_host_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "core_host_shim.cpp"), "-o", so])
    L = C.CDLL(so)
    L.shim_pairs.argtypes = [u32p, u32p, C.c_size_t] + [u8p] * 8 + [u64p]
    L.shim_edge.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    L.shim_edge.restype = C.c_int
    L.shim_split.argtypes = [u32p, u32p, C.c_size_t] + [u8p] * 4
    L.shim_pass_pred.argtypes = [C.c_int, u32p, u32p, C.c_size_t, u8p]
    L.shim_pass_possible.argtypes = [C.c_int, C.c_int] + [C.c_uint32] * 4
    L.shim_pass_possible.restype = C.c_int
    L.shim_quick.argtypes = [C.c_int, u32p, u32p, C.c_size_t, u8p]
    L.shim_quick_any.argtypes = [C.c_int, u32p, u32p, C.c_size_t, u8p]
    L.shim_top_possible.argtypes = [C.c_int] + [C.c_uint32] * 4
    L.shim_top_possible.restype = C.c_int
    L.shim_seed2_first.argtypes = [u32p, u32p, C.c_size_t, np.ctypeslib.ndpointer(np.int8, flags="C")]
    L.shim_seed2_keys.argtypes = [C.c_int, u32p, u32p, C.c_size_t, u32p, u32p]
    L.shim_seed2_count.restype = C.c_int
    L.shim_seed2_permute.argtypes = [C.c_int, u32p, u32p, C.c_size_t, u32p, u32p, u32p, u32p]
    L.shim_seed_tiles_meet.argtypes = [C.c_uint32] * 4 + [C.c_int]
    return L


def run(shim, a, b):
    n = a.size
    outs = [np.zeros(n, np.uint8) for _ in range(8)]
    mult = np.zeros(n, np.uint64)
    shim.shim_pairs(a, b, n, *outs, mult)
    return dict(zip(["pre1", "pre2", "dsmall", "dplain", "dfull", "da15", "db15", "S"], outs)), mult


def edit_ops(rng, a, k):
    """Apply k random edit operations to the 16-mer a, keep 16 bases (pad randomly)."""
    s = [(a >> (2 * i)) & 3 for i in range(16)]
    for _ in range(k):
        op = rng.integers(0, 3); pos = int(rng.integers(0, len(s)))
        if op == 0:
            s[pos] = (s[pos] + int(rng.integers(1, 4))) & 3
        elif op == 1:
            s.insert(pos, int(rng.integers(0, 4)))
        else:
            del s[pos]
    while len(s) < 16:
        s.append(int(rng.integers(0, 4)))
    return sum(c << (2 * i) for i, c in enumerate(s[:16]))


def make_pairs(seed, n_near=60000, n_rand=20000, n_low=20000):
    rng = np.random.default_rng(seed)
    a, b = [], []
    for _ in range(n_near):
        x = int(rng.integers(0, 1 << 32))
        if rng.random() < 0.3:   # low-complexity seeds exercise the multiplicity side of S
            unit = int(rng.integers(0, 1 << (2 * int(rng.integers(1, 5)))))
            ul = max(1, unit.bit_length() + 1) // 2 or 1
            x = 0
            for i in range(16):
                x |= ((unit >> (2 * (i % ul))) & 3) << (2 * i)
        y = edit_ops(rng, x, int(rng.integers(1, 5)))
        if x != y:
            a.append(x); b.append(y)
    ra = rng.integers(0, 1 << 32, n_rand, dtype=np.uint64); rb = rng.integers(0, 1 << 32, n_rand, dtype=np.uint64)
    a += ra.tolist(); b += rb.tolist()
    # pairs sharing long stretches on shifted diagonals
    for _ in range(n_low):
        x = int(rng.integers(0, 1 << 32)); sh = int(rng.integers(1, 4)) * 2
        y = ((x << sh) | int(rng.integers(0, 1 << sh))) & 0xFFFFFFFF if rng.random() < 0.5 else (x >> sh) | (int(rng.integers(0, 1 << sh)) << (32 - sh))
        if rng.random() < 0.5:
            y = edit_ops(rng, y, 1)
        if x != y:
            a.append(x); b.append(y)
    a = np.asarray(a, dtype=np.uint32); b = np.asarray(b, dtype=np.uint32)
    keep = a != b
    return a[keep], b[keep]


def test_core_vs_oracle(shim):
    a, b = make_pairs(5)
    res, mult = run(shim, a, b)
    L = orc.lib()
    n = a.size
    ed = np.fromiter((L.orc_ed(int(x), 16, int(y), 16) for x, y in zip(a, b)), np.int32, n)
    ea = np.fromiter((L.orc_ed(int(x), 15, int(y), 16) for x, y in zip(a, b)), np.int32, n)
    eb = np.fromiter((L.orc_ed(int(x), 16, int(y), 15) for x, y in zip(a, b)), np.int32, n)
    D = np.minimum(ed, np.minimum(ea, eb))
    S = np.fromiter((L.orc_S(int(x), int(y)) for x, y in zip(a, b)), np.int32, n)
    assert np.array_equal(res["dfull"], ed)
    assert np.array_equal(res["da15"], ea)
    assert np.array_equal(res["db15"], eb)
    assert np.array_equal(res["S"], S)
    assert np.array_equal(res["dsmall"], np.minimum(D, 3))
    assert np.array_equal(res["dplain"], np.minimum(ed, 3))
    # soundness of the prefilters: D<=t  =>  prefilter passes
    assert res["pre1"][D <= 1].all()
    assert res["pre2"][D <= 2].all()
    assert (D <= 1).sum() > 5000 and ((D == 2).sum() > 5000)
    # the prefilters do reject random pairs
    rng = np.random.default_rng(1)
    ra = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    rb = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    rr, _ = run(shim, ra, rb)
    assert rr["pre1"].mean() < 4e-4      # expected 2*2^-16 + 2*2^-14 = 1.5e-4
    assert rr["pre2"].mean() < 0.012     # expected 9 * 2^-10   = 0.88 %
    # symmetry
    res2, _ = run(shim, b, a)
    for k in ("pre1", "dsmall", "dplain", "S"):   # pre2 is frame-dependent (sound either way)
        assert np.array_equal(res[k], res2[k]), k
    # multiplicities: sum over positions == S, and equal to the oracle's kmer_score multiplicities
    nib = np.stack([(mult >> np.uint64(4 * p)) & np.uint64(15) for p in range(11)], 1).astype(np.int32)
    assert np.array_equal(nib.sum(1), S)
    idx = np.arange(0, n, 97)
    for i in idx:
        _, m = orc.kmer_score(a[i:i + 1], b[i:i + 1])
        assert m[0, 0].tolist() == nib[i].tolist()


def test_edge_predicate_on_golden(shim, gold_pairs):
    for p in gold_pairs["pairs"]:
        for t in range(0, 5):
            want = p["D"] if (p["S"] >= orc.T(t) and p["D"] <= t) else 0
            assert shim.shim_edge(p["ra"], p["rb"], t) == want
            assert shim.shim_edge(p["rb"], p["ra"], t) == want


def test_exhaustive_neighbourhood(shim):
    """Every string within two edit operations of a few seeds (incl. truncation effects)."""
    rng = np.random.default_rng(9)
    seeds = [int(rng.integers(0, 1 << 32)) for _ in range(3)] + [0, 0x11111111 * 0 + 0x44444444, int(synth.rank_many(["ACACACACACACACAC"])[0])]
    for x in seeds:
        s = [(x >> (2 * i)) & 3 for i in range(16)]
        neigh = set()

        def one(seq):
            out = []
            for pos in range(len(seq)):
                for c in range(4):
                    if c != seq[pos]:
                        out.append(seq[:pos] + [c] + seq[pos + 1:])
                out.append(seq[:pos] + seq[pos + 1:])
            for pos in range(len(seq) + 1):
                for c in range(4):
                    out.append(seq[:pos] + [c] + seq[pos:])
            return out

        lvl1 = one(s)
        sample = [lvl1[i] for i in rng.choice(len(lvl1), 40, replace=False)]
        lvl2 = [y for z in sample for y in one(z)]
        for seq in lvl1 + lvl2:
            for pad in range(4):
                q = (seq + [pad, pad])[:16] if len(seq) < 16 else seq[:16]
                neigh.add(sum(c << (2 * i) for i, c in enumerate(q)))
        neigh.discard(x)
        b = np.fromiter(neigh, dtype=np.uint32)
        a = np.full(b.size, x, dtype=np.uint32)
        res, _ = run(shim, a, b)
        L = orc.lib()
        D = np.fromiter((L.orc_D(int(x), int(y)) for y in b), np.int32, b.size)
        assert np.array_equal(res["dsmall"], np.minimum(D, 3))
        assert res["pre1"][D <= 1].all() and res["pre2"][D <= 2].all()


def split(shim, a, b):
    outs = [np.zeros(a.size, np.uint8) for _ in range(4)]
    shim.shim_split(a, b, a.size, *outs)
    return dict(zip(["l1", "t1", "l2", "t2"], outs))


def test_split_filters_sound(shim):
    """light || top is a necessary condition for D <= t in both frames (bdg_core.cuh, tiled edge kernel)."""
    L = orc.lib()
    tot1 = tot2 = 0
    for seed in (5, 6, 7):
        a, b = make_pairs(seed)
        D = np.fromiter((L.orc_D(int(x), int(y)) for x, y in zip(a, b)), np.int32, a.size)
        for x, y in ((a, b), (b, a)):
            r = split(shim, x, y)
            assert (r["l1"] | r["t1"])[D <= 1].all()
            assert (r["l2"] | r["t2"])[D <= 2].all()
        tot1 += int((D <= 1).sum()); tot2 += int((D == 2).sum())
    assert tot1 > 15000 and tot2 > 15000
    rng = np.random.default_rng(2)
    ra = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    rb = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    r = split(shim, ra, rb)
    assert r["l1"].mean() < 6e-5 and r["t1"].mean() < 3e-4         # 2^-15 ; 2^-16 + 2*2^-14
    assert r["l2"].mean() < 6e-3 and r["t2"].mean() < 4.5e-3       # 4*2^-10 ; 3*2^-10


def test_split_filters_exhaustive_neighbourhood(shim):
    """All strings within two operations of a few seeds, all paddings: no edge candidate is lost."""
    rng = np.random.default_rng(11)
    L = orc.lib()
    seeds = [int(rng.integers(0, 1 << 32)) for _ in range(4)] + [0, 0xFFFFFFFF, 0x44444444, int(synth.rank_many(["ACACACACACACACAC"])[0])]
    for x in seeds:
        s = [(x >> (2 * i)) & 3 for i in range(16)]

        def one(seq):
            out = []
            for pos in range(len(seq)):
                for c in range(4):
                    if c != seq[pos]:
                        out.append(seq[:pos] + [c] + seq[pos + 1:])
                out.append(seq[:pos] + seq[pos + 1:])
            for pos in range(len(seq) + 1):
                for c in range(4):
                    out.append(seq[:pos] + [c] + seq[pos:])
            return out

        lvl1 = one(s)
        sample = [lvl1[i] for i in rng.choice(len(lvl1), 60, replace=False)]
        neigh = set()
        for seq in lvl1 + [y for z in sample for y in one(z)]:
            for pad in range(16):
                q = (seq + [pad & 3, pad >> 2])[:16]
                neigh.add(sum(c << (2 * i) for i, c in enumerate(q)))
        neigh.discard(x)
        b = np.fromiter(neigh, dtype=np.uint32)
        a = np.full(b.size, x, dtype=np.uint32)
        D = np.fromiter((L.orc_D(int(x), int(y)) for y in b), np.int32, b.size)
        for p, q in ((a, b), (b, a)):
            r = split(shim, p, q)
            assert (r["l1"] | r["t1"])[D <= 1].all()
            assert (r["l2"] | r["t2"])[D <= 2].all()


@pytest.mark.parametrize("t", [1, 2])
def test_top_possible_is_conservative(shim, t):
    """If some pair of a (row run) x (column run) tile of a sorted array meets a top condition, the interval test
    on the runs' end points must say so."""
    rng = np.random.default_rng(20 + t)
    checked = hits = skipped = 0
    for n, rows, cols in ((3000, 64, 32), (20000, 256, 128), (200000, 256, 256)):
        base = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        # plant near-duplicates so that top conditions do occur
        extra = (base[: n // 4] ^ rng.integers(0, 1 << 12, n // 4, dtype=np.uint64).astype(np.uint32))
        s = np.unique(np.concatenate([base, extra]))
        for _ in range(300):
            r0 = int(rng.integers(0, s.size - rows)); c0 = int(rng.integers(0, s.size - cols))
            if rng.random() < 0.3:
                c0 = min(s.size - cols, r0 + int(rng.integers(0, rows)))
            A = s[r0:r0 + rows]; B = s[c0:c0 + cols]
            aa = np.repeat(A, B.size); bb = np.tile(B, A.size)
            r = split(shim, aa, bb)
            any_top = bool(r["t%d" % t].any())
            poss = bool(shim.shim_top_possible(t, int(A[0]), int(A[-1]), int(B[0]), int(B[-1])))
            assert poss or not any_top
            checked += 1; hits += any_top; skipped += (not poss)
    assert hits > 50 and skipped > 200


def pass_mask(shim, t, a, b):
    m = np.zeros(a.size, np.uint8)
    shim.shim_pass_pred(t, a, b, a.size, m)
    assert not (m & 0x80).any(), "pass_pred and pass_pred_rot disagree"
    return m


def rotl(v, r):
    v = v.astype(np.uint32)
    return v if r == 0 else ((v << np.uint32(r)) | (v >> np.uint32(32 - r))).astype(np.uint32)


PASS_ROT = {1: [0, 16], 2: [0, 10, 20]}


@pytest.mark.parametrize("t", [1, 2])
def test_pass_predicates_cover_and_are_symmetric(shim, t):
    """Sparse edge passes: every pair with D <= t fires in at least one pass; predicates are orientation-free."""
    L = orc.lib()
    tot = 0
    for seed in (5, 8):
        a, b = make_pairs(seed)
        D = np.fromiter((L.orc_D(int(x), int(y)) for x, y in zip(a, b)), np.int32, a.size)
        m = pass_mask(shim, t, a, b)
        assert np.array_equal(m, pass_mask(shim, t, b, a))
        assert (m[D <= t] != 0).all()
        tot += int((D <= t).sum())
    assert tot > 10000
    rng = np.random.default_rng(3)
    ra = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    rb = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    assert (pass_mask(shim, t, ra, rb) != 0).mean() < (4e-4 if t == 1 else 0.014)


@pytest.mark.parametrize("t", [1, 2])
def test_pass_possible_is_conservative(shim, t):
    """Interval test of every pass on arrays sorted by that pass's rotated key."""
    rng = np.random.default_rng(40 + t)
    for p, rot in enumerate(PASS_ROT[t]):
        hits = skipped = 0
        for n, rows, cols in ((3000, 64, 32), (20000, 256, 128), (200000, 256, 128)):
            base = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
            noise = rng.integers(0, 1 << 12, n // 4, dtype=np.uint64).astype(np.uint32)
            extra = base[: n // 4] ^ rotl(noise, int(rng.integers(0, 32)))
            s = np.unique(rotl(np.concatenate([base, extra]), rot))       # sorted by the rotated key
            for _ in range(250):
                r0 = int(rng.integers(0, s.size - rows)); c0 = int(rng.integers(0, s.size - cols))
                if rng.random() < 0.3:
                    c0 = min(s.size - cols, r0 + int(rng.integers(0, rows)))
                A = s[r0:r0 + rows]; B = s[c0:c0 + cols]
                aa = np.repeat(A, B.size); bb = np.tile(B, A.size)
                fires = ((pass_mask(shim, t, rotl(aa, (32 - rot) % 32), rotl(bb, (32 - rot) % 32)) >> p) & 1).any()
                poss = bool(shim.shim_pass_possible(t, p, int(A[0]), int(A[-1]), int(B[0]), int(B[-1])))
                assert poss or not fires, (t, p)
                hits += bool(fires); skipped += (not poss)
        assert hits > 30 and skipped > 150, (t, p, hits, skipped)


@pytest.mark.parametrize("t", [1, 2])
def test_quick_pass_sound(shim, t):
    """The per-pair test of the sparse passes never rejects a pair with D <= t (both frames), and rejects most others."""
    L = orc.lib()
    for seed in (5, 12):
        a, b = make_pairs(seed)
        D = np.fromiter((L.orc_D(int(x), int(y)) for x, y in zip(a, b)), np.int32, a.size)
        for x, y in ((a, b), (b, a)):
            out = np.zeros(x.size, np.uint8)
            shim.shim_quick(t, x, y, x.size, out)
            assert out[D <= t].all()
    rng = np.random.default_rng(4)
    ra = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    rb = rng.integers(0, 1 << 32, 1 << 20, dtype=np.uint64).astype(np.uint32)
    out = np.zeros(ra.size, np.uint8)
    shim.shim_quick(t, ra, rb, ra.size, out)
    assert out.mean() < (0.004 if t == 1 else 0.03)


@pytest.mark.parametrize("t", [2, 3, 5])
def test_quick_pass_any_sound(shim, t):
    """Generic quick test of the dense kernel (t >= 3): never rejects a pair with D <= t, in both frames."""
    L = orc.lib()
    a, b = make_pairs(30 + t, n_near=25000, n_rand=3000, n_low=8000)
    rng = np.random.default_rng(t)
    extra_a, extra_b = [], []
    for _ in range(12000):                      # more operations than make_pairs uses, so that D = 3..5 is well covered
        x = int(rng.integers(0, 1 << 32)); y = edit_ops(rng, x, int(rng.integers(3, 7)))
        if x != y:
            extra_a.append(x); extra_b.append(y)
    a = np.concatenate([a, np.asarray(extra_a, np.uint32)]); b = np.concatenate([b, np.asarray(extra_b, np.uint32)])
    D = np.fromiter((L.orc_D(int(x), int(y)) for x, y in zip(a, b)), np.int32, a.size)
    for x, y in ((a, b), (b, a)):
        out = np.zeros(x.size, np.uint8)
        shim.shim_quick_any(t, x, y, x.size, out)
        assert out[D <= t].all()
    assert (D == t).sum() > 1000
    ra = rng.integers(0, 1 << 32, 1 << 18, dtype=np.uint64).astype(np.uint32)
    rb = rng.integers(0, 1 << 32, 1 << 18, dtype=np.uint64).astype(np.uint32)
    out = np.zeros(ra.size, np.uint8)
    shim.shim_quick_any(t, ra, rb, ra.size, out)
    print("t", t, "random pass rate", out.mean())


# ------------------------------------------------------------------------------------------ two-block seeds (t = 2)
def seed2_first(shim, a, b):
    out = np.zeros(a.size, np.int8)
    shim.shim_seed2_first(np.ascontiguousarray(a), np.ascontiguousarray(b), a.size, out)
    return out


def two_op_neighbourhood(x):
    """Every 16-mer reachable from x by at most two edit operations (all paddings of the shortened ones)."""
    def one(seq):
        out = []
        for pos in range(len(seq)):
            for c in range(4):
                if c != seq[pos]:
                    out.append(seq[:pos] + [c] + seq[pos + 1:])
            out.append(seq[:pos] + seq[pos + 1:])
        for pos in range(len(seq) + 1):
            for c in range(4):
                out.append(seq[:pos] + [c] + seq[pos:])
        return out

    s = [(x >> (2 * i)) & 3 for i in range(16)]
    lvl1 = one(s)
    neigh = set()
    for seq in lvl1 + [y for z in lvl1 for y in one(z)]:
        if len(seq) >= 16:
            neigh.add(sum(c << (2 * i) for i, c in enumerate(seq[:16])))
        else:
            for p1 in range(4):
                for p2 in range(4):
                    neigh.add(sum(c << (2 * i) for i, c in enumerate((seq + [p1, p2])[:16])))
    neigh.discard(x)
    return np.fromiter(neigh, dtype=np.uint32)


def test_seed2_conditions_are_necessary_for_d2(shim):
    """bdg_core.cuh two-block seeds: D(a,b) <= 2 implies that one of the 20 key equalities holds, under either labelling
    of the pair (so a join may fix a = min); checked on random near pairs, low-complexity and shifted pairs, and on the
    complete two-operation neighbourhoods of random and repetitive seeds."""
    L = orc.lib()
    assert shim.shim_seed2_count() == 20
    tot = 0
    for seed in (5, 8):
        a, b = make_pairs(seed)
        D = np.fromiter((L.orc_D(int(x), int(y)) for x, y in zip(a, b)), np.int32, a.size)
        near = D <= 2
        assert (seed2_first(shim, a, b)[near] >= 0).all()
        assert (seed2_first(shim, b, a)[near] >= 0).all()
        assert (seed2_first(shim, np.minimum(a, b), np.maximum(a, b))[near] >= 0).all()
        tot += int(near.sum())
    assert tot > 50000
    rng = np.random.default_rng(9)
    seeds = [int(rng.integers(0, 1 << 32)) for _ in range(3)] + [0, 0x44444444] + \
        [int(v) for v in synth.rank_many(["ACACACACACACACAC", "AAAAAAAACCCCCCCC", "ACGACGACGACGACGA", "TTTTGTTTTGTTTTGT"])]
    for x in seeds:
        b = two_op_neighbourhood(x)
        a = np.full(b.size, x, dtype=np.uint32)
        D = np.fromiter((L.orc_D(x, int(y)) for y in b), np.int32, b.size)
        near = D <= 2
        assert near.sum() > 1000
        assert (seed2_first(shim, a, b)[near] >= 0).all() and (seed2_first(shim, b, a)[near] >= 0).all()


def test_seed2_keys_and_selectivity(shim):
    """The join keys are what the predicate compares (14 or 16 bits), every condition fires on its own construction, and a
    random pair meets some condition with probability ~ 69 / 65536 (an order of magnitude below the single-block passes)."""
    rng = np.random.default_rng(4)
    n = 1 << 21
    ra = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    rb = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    first = seed2_first(shim, ra, rb)
    rate = (first >= 0).mean()
    assert 0.5 * 69 / 65536 < rate < 1.2 * 69 / 65536
    ka = np.zeros(n, np.uint32); kb = np.zeros(n, np.uint32)
    fired = np.zeros(n, bool)
    for c in range(20):
        shim.shim_seed2_keys(c, ra, rb, n, ka, kb)
        bits = 14 if (4 <= c <= 6 or c >= 10) else 16
        assert int(ka.max()) < (1 << bits) and int(kb.max()) < (1 << bits)
        eq = ka == kb
        assert np.array_equal(first == c, eq & ~fired)        # `first` is the lowest condition whose keys agree
        fired |= eq
        # construct a partner that meets condition c and nothing else is required: copy a's two blocks into b at the shifted place
        blocks = {0: [(0, 0), (1, 0)]}
        for k in range(1, 4): blocks[k] = [(1, k - 2), (2, k - 2)]
        for k in range(4, 7): blocks[k] = [(2, k - 5), (3, k - 5)]
        for k in range(7, 10): blocks[k] = [(0, 0), (2, k - 8)]
        for k in range(10, 13): blocks[k] = [(0, 0), (3, k - 11)]
        for k, (d1, d3) in enumerate([(-1, -1), (-1, 0), (0, -1), (0, 0), (0, 1), (1, 0), (1, 1)]):
            blocks[13 + k] = [(1, d1), (3, d3)]
        y = rb[:4096].astype(np.uint64)
        x = ra[:4096].astype(np.uint64)
        for blk, d in blocks[c]:
            lo, nb = 8 * blk, (6 if blk == 3 else 8)
            mask = np.uint64(((1 << nb) - 1) << (lo + 2 * d))
            y = (y & ~mask) | (((x >> np.uint64(lo)) & np.uint64((1 << nb) - 1)) << np.uint64(lo + 2 * d))
        shim.shim_seed2_keys(c, ra[:4096].copy(), y.astype(np.uint32), 4096, ka[:4096], kb[:4096])
        assert np.array_equal(ka[:4096], kb[:4096]), c


def test_seed2_sort_form(shim):
    """The permuted words of a condition: a bijection (unpermute gives the barcode back), the join key on top (equal top
    bits <=> the condition holds), order of the words = order of (key, remaining bits), and the tile test is conservative."""
    rng = np.random.default_rng(6)
    n = 1 << 18
    a = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    b = a.copy()
    flip = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) & rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) \
        & rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    b ^= flip                                                   # ~1/8 of the bits differ: many key hits, many misses
    pa, pb, ua, ub, ka, kb_ = (np.zeros(n, np.uint32) for _ in range(6))
    for c in range(20):
        shim.shim_seed2_permute(c, a, b, n, pa, pb, ua, ub)
        assert np.array_equal(ua, a) and np.array_equal(ub, b), c
        bits = shim.shim_seed2_key_bits(c)
        shim.shim_seed2_keys(c, a, b, n, ka, kb_)
        assert np.array_equal(pa >> np.uint32(32 - bits), ka) and np.array_equal(pb >> np.uint32(32 - bits), kb_), c
        assert np.unique(pa).size == np.unique(a).size               # injective
        # tiles of the two sorted sides: whenever a tile pair holds an equal key, the tile test says so
        sa, sb = np.sort(pa), np.sort(pb)
        for _ in range(300):
            i = int(rng.integers(0, n - 64)); j = int(np.searchsorted(sb >> np.uint32(32 - bits), sa[i] >> np.uint32(32 - bits)))
            j = min(max(j + int(rng.integers(-200, 200)), 0), n - 64)
            A, B = sa[i:i + 64], sb[j:j + 64]
            hit = np.intersect1d(A >> np.uint32(32 - bits), B >> np.uint32(32 - bits)).size > 0
            poss = bool(shim.shim_seed_tiles_meet(int(A[0]), int(A[-1]), int(B[0]), int(B[-1]), bits))
            assert poss or not hit


def test_seed2_join_passes_reproduce_the_oracle_edge_set(shim):
    """The planned join passes, emulated in numpy: for every condition both sides are permuted and sorted, equal-key buckets
    are paired, a pair is kept under the labelling a = min by the FIRST condition it meets, then the exact D and S decide.
    The union over the 20 passes must be the oracle's edge set at t = 2, every edge exactly once."""
    L = orc.lib()
    rng = synth.rng_for(91)
    cells = rng.integers(0, 1 << 32, 120, dtype=np.uint64).astype(np.uint32)
    obs, _ = synth.simulate_reads(cells, 9000, 0.06, rng)
    s = np.unique(obs)
    n = s.size
    wa, wb, wd, _ = orc.Index(s).edges(2)
    want = sorted(zip(wa.tolist(), wb.tolist(), wd.tolist()))
    got = []
    pa, pb, ua, ub = (np.zeros(n, np.uint32) for _ in range(4))
    n_cand = 0
    for c in range(20):
        shim.shim_seed2_permute(c, s, s, n, pa, pb, ua, ub)
        bits = shim.shim_seed2_key_bits(c)
        oa, ob = np.argsort(pa, kind="stable"), np.argsort(pb, kind="stable")          # sorted sides; the barcode rides along
        ka, kb = pa[oa] >> np.uint32(32 - bits), pb[ob] >> np.uint32(32 - bits)
        lo, hi = np.searchsorted(kb, ka, "left"), np.searchsorted(kb, ka, "right")
        rows = np.repeat(np.arange(n), hi - lo)
        cols = np.concatenate([np.arange(l, h) for l, h in zip(lo.tolist(), hi.tolist())]) if rows.size else np.empty(0, np.int64)
        x, y = s[oa[rows]], s[ob[cols]]
        keep = x < y
        x, y = np.ascontiguousarray(x[keep]), np.ascontiguousarray(y[keep])
        n_cand += x.size
        first = seed2_first(shim, x, y)
        for xv, yv in zip(x[first == c].tolist(), y[first == c].tolist()):
            d = L.orc_D(xv, yv)
            if d <= 2 and L.orc_S(xv, yv) >= 4:
                got.append((xv, yv, d))
    assert len(got) == len(set(got))                        # no pair twice
    assert sorted(got) == want and len(want) > 5000
    assert n_cand < 0.02 * n * (n - 1) / 2                  # the seeds leave a small share of the pairs (clustered data)
