"""ctypes front-end of the CPU oracle (oracle/badger_oracle.c) plus small pure-Python
restatements of the reference's host-side steps.  TEST INFRASTRUCTURE ONLY -- see the header
of badger_oracle.c.  Parity is pinned by tests/test_oracle_golden.py (fixtures generated from
the unmodified reference by oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections import defaultdict
from statistics import mean

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbadger_oracle.so")
_lib = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "badger_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_rank16.argtypes = [C.c_char_p, C.POINTER(C.c_uint32)]
        L.orc_rank16.restype = C.c_int
        L.orc_unrank16.argtypes = [C.c_uint32, C.c_char_p]
        L.orc_ed.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_int]
        L.orc_ed.restype = C.c_int
        for f in (L.orc_D, L.orc_S):
            f.argtypes = [C.c_uint32, C.c_uint32]
            f.restype = C.c_int
        L.orc_T.argtypes = [C.c_int]
        L.orc_T.restype = C.c_int
        L.orc_edge.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
        L.orc_edge.restype = C.c_int
        L.orc_edges_brute.argtypes = [u32p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_edges_brute.restype = C.c_int64
        L.orc_index_build.argtypes = [u32p, C.c_size_t]
        L.orc_index_build.restype = C.c_void_p
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_get_close.argtypes = [C.c_void_p, C.c_size_t, C.c_int, u32p]
        L.orc_get_close.restype = C.c_size_t
        L.orc_edges_index.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]
        L.orc_edges_index.restype = C.c_int64
        L.orc_edges_rows_both.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_edges_rows_both.restype = C.c_int64
        L.orc_dedup_count.argtypes = [u32p, C.c_void_p, C.c_size_t, u32p, u32p]
        L.orc_dedup_count.restype = C.c_size_t
        L.orc_member.argtypes = [u32p, C.c_size_t, u32p, C.c_size_t, u8p]
        L.orc_nearest.argtypes = [u32p, C.c_size_t, u32p, C.c_size_t, C.c_int, i32p, u8p]
        L.orc_kmer_score.argtypes = [u32p, C.c_size_t, u32p, C.c_size_t, u8p, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


# ------------------------------------------------------------------ scalar helpers
def rank(seq: str) -> int:
    out = C.c_uint32()
    if len(seq) < 16 or lib().orc_rank16(seq.encode(), C.byref(out)) != 0:
        raise KeyError(seq)
    return out.value


def unrank(r: int) -> str:
    buf = C.create_string_buffer(16)
    lib().orc_unrank16(r, buf)
    return buf.raw.decode()


def ed(a: int, b: int, la: int = 16, lb: int = 16) -> int:
    return lib().orc_ed(a, la, b, lb)


def D(a: int, b: int) -> int:
    return lib().orc_D(a, b)


def S(a: int, b: int) -> int:
    return lib().orc_S(a, b)


def T(t: int) -> int:
    return lib().orc_T(t)


def edge(a: int, b: int, t: int) -> int:
    return lib().orc_edge(a, b, t)


# ------------------------------------------------------------------ edge sets
def _canon(a, b, d):
    order = np.lexsort((b, a))
    return a[order], b[order], d[order]


def edges_brute(ranks: np.ndarray, t: int):
    """All-pairs application of the predicate; returns canonical (a,b,d) arrays sorted by (a,b)."""
    ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
    n = lib().orc_edges_brute(ranks, ranks.size, t, None, None, None, 0)
    a = np.empty(n, np.uint32); b = np.empty(n, np.uint32); d = np.empty(n, np.uint8)
    if n:
        lib().orc_edges_brute(ranks, ranks.size, t, a.ctypes.data, b.ctypes.data, d.ctypes.data, n)
    return _canon(a, b, d)


class Index:
    """The reference's 6-mer index (index.py:12-41) over a set of distinct ranks."""

    def __init__(self, ranks: np.ndarray):
        self.ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
        self._h = lib().orc_index_build(self.ranks, self.ranks.size)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_index_free(self._h)
            self._h = None

    def get_close(self, row: int, t: int) -> np.ndarray:
        out = np.empty(max(self.ranks.size, 1), np.uint32)
        n = lib().orc_get_close(self._h, row, t, out)
        return np.sort(out[:n])

    def edges(self, t: int, rows=None, threads: int = 0, cap: int | None = None):
        """graph_construction (barcode_graph.py:224-249) over all rows or a sample of rows.
        Returns (a, b, d, candidates_verified)."""
        rows_p, nrows = None, 0
        if rows is not None:
            rows = np.ascontiguousarray(rows, dtype=np.uint32)
            rows_p, nrows = rows.ctypes.data, rows.size
        ver = C.c_uint64()
        if cap is None:
            cap = lib().orc_edges_index(self._h, t, rows_p, nrows, threads, None, None, None, 0, C.byref(ver))
        a = np.empty(cap, np.uint32); b = np.empty(cap, np.uint32); d = np.empty(cap, np.uint8)
        n = lib().orc_edges_index(self._h, t, rows_p, nrows, threads,
                                  a.ctypes.data, b.ctypes.data, d.ctypes.data, cap, C.byref(ver))
        n = min(n, cap)
        a, b, d = _canon(a[:n], b[:n], d[:n])
        return a, b, d, ver.value


def edges_touching(ix: "Index", t: int, rows, threads: int = 0):
    """Every edge with one of `rows` (indices into ix.ranks) as its smaller OR larger end point, each once, canonical order:
    the adjacency lists graph_construction leaves for those rows (barcode_graph.py:245-249)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    cap = lib().orc_edges_rows_both(ix._h, t, rows.ctypes.data, rows.size, threads, None, None, None, 0)
    a = np.empty(cap, np.uint32); b = np.empty(cap, np.uint32); d = np.empty(cap, np.uint8)
    n = lib().orc_edges_rows_both(ix._h, t, rows.ctypes.data, rows.size, threads, a.ctypes.data, b.ctypes.data, d.ctypes.data, cap)
    assert n == cap
    key = (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)
    _, first = np.unique(key, return_index=True)            # an edge between two sampled rows was found from both
    return _canon(a[first], b[first], d[first])


def dedup_count(reads: np.ndarray, valid: np.ndarray | None = None):
    """barcode_graph.py:192-204 -> (ranks, counts) in first-seen order."""
    reads = np.ascontiguousarray(reads, dtype=np.uint32)
    vp = None
    if valid is not None:
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
        vp = valid.ctypes.data
    r = np.empty(max(reads.size, 1), np.uint32); c = np.empty(max(reads.size, 1), np.uint32)
    n = lib().orc_dedup_count(reads, vp, reads.size, r, c)
    return r[:n].copy(), c[:n].copy()


def member(sorted_wl: np.ndarray, q: np.ndarray) -> np.ndarray:
    sorted_wl = np.ascontiguousarray(sorted_wl, dtype=np.uint32)
    q = np.ascontiguousarray(q, dtype=np.uint32)
    hit = np.zeros(max(q.size, 1), np.uint8)
    lib().orc_member(sorted_wl if sorted_wl.size else np.zeros(1, np.uint32), sorted_wl.size, q if q.size else np.zeros(1, np.uint32), q.size, hit)
    return hit[:q.size]


def nearest(q: np.ndarray, targets: np.ndarray, max_d: int = 2):
    q = np.ascontiguousarray(q, dtype=np.uint32)
    targets = np.ascontiguousarray(targets, dtype=np.uint32)
    am = np.full(max(q.size, 1), -1, np.int32); dist = np.full(max(q.size, 1), 255, np.uint8)
    if q.size:
        lib().orc_nearest(q, q.size, targets if targets.size else np.zeros(1, np.uint32), targets.size, max_d, am, dist)
    return am[:q.size], dist[:q.size]


def kmer_score(q: np.ndarray, wl: np.ndarray, want_mult: bool = True):
    q = np.ascontiguousarray(q, dtype=np.uint32)
    wl = np.ascontiguousarray(wl, dtype=np.uint32)
    cnt = np.zeros((q.size, wl.size), np.uint8)
    mult = np.zeros((q.size, wl.size, 11), np.uint8) if want_mult else None
    if q.size and wl.size:
        lib().orc_kmer_score(q, q.size, wl, wl.size, cnt.reshape(-1), mult.ctypes.data if want_mult else None)
    return cnt, mult


# ------------------------------------------------------------------ host-side steps (pure Python, small cases)
def get_occurrences(known, sequence, k=6, max_hits=0, min_kmers=1, hits_delta=1, ignore_equal=False):
    """kmer_indexer.py:49-75 restated; returns [(string, count, positions)] in the reference's dict order."""
    index = defaultdict(list)
    for i, s in enumerate(known):                      # kmer_indexer.py:29-32
        for p in range(0, len(s) - k + 1):
            index[s[p:p + k]].append(i)
    counts, positions = {}, {}
    for pos in range(0, len(sequence) - k + 1):        # kmer_indexer.py:52-55
        for i in index.get(sequence[pos:pos + k], ()):
            counts[i] = counts.get(i, 0) + 1
            positions.setdefault(i, []).append(pos)
    result = []
    for i, c in counts.items():                         # first-touch order (dict insertion)
        if c < min_kmers:
            continue
        if ignore_equal and known[i] == sequence:
            continue
        result.append((known[i], c, positions[i]))
    if not result:
        return []
    top = max(r[1] for r in result)
    result = [r for r in result if r[1] >= top - hits_delta]
    result.sort(key=lambda r: r[1], reverse=True)       # stable
    if max_hits:
        result = result[:max_hits]
    out = {}
    for r in result:                                    # dict keyed by string: later duplicates overwrite value, keep position
        out[r[0]] = r
    return list(out.values())


def cluster_centers(counts: dict, n_cells: int, interval: int, true_ranks=None, whitelist: set | None = None):
    """barcode_graph.py:252-277.  counts: insertion-ordered {rank: count}; whitelist: set of ranks or None."""
    by_counts = [k for k, _ in sorted(counts.items(), key=lambda kv: kv[1], reverse=True)]
    cutoff = max(mean(list(counts.values())[:n_cells]) / 5.0, 5)
    tbcs, n, i = [], 0, 0
    hi = n_cells + n_cells * interval * 0.01
    if true_ranks:
        tbcs = list(true_ranks)
    elif whitelist:
        while i < len(by_counts) and counts[by_counts[i]] > cutoff and n <= hi:
            if by_counts[i] in whitelist:
                tbcs.append(by_counts[i]); n += 1
            i += 1
    else:
        while counts[by_counts[i]] > cutoff and n <= hi:
            tbcs.append(by_counts[i]); i += 1; n += 1
    while n < n_cells - n_cells * interval * 0.01:
        tbcs.append(by_counts[i]); i += 1; n += 1         # IndexError when N is small, as in the reference
    return tbcs


def cluster(adj: dict, centres):
    """barcode_graph.py:283-301 restated literally.  adj: {rank: iterable of neighbour ranks}."""
    clusters, clustering, clustered = {}, {}, defaultdict(bool)
    for t in centres:
        clusters[t] = [t]; clustering[t] = (t, 0); clustered[t] = True
    for i in (1, 2):
        for center in clusters.keys():
            for n in range(len(clusters[center])):
                node = clusters[center][n]
                for nb in adj.get(node, ()):
                    if not clustered[nb]:
                        clusters[center].append(nb); clustering[nb] = (center, i); clustered[nb] = True
                    elif clustering[nb][0] != center and clustering[nb][0] != -1:
                        if clustering[nb][1] == i:
                            clusters[clustering[nb][0]].remove(nb)
                            clustering[nb] = (-1, -1)
    return clustering


def num_threads() -> int:
    return lib().orc_num_threads()
