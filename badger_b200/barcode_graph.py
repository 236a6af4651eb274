"""BarcodeGraph with the reference's interface (algbio/Badger barcode_graph.py:38-410), backed by the
B200 kernels.  Same method names, arguments, attributes and results; the O(N^2) candidate search and the
edit-distance verification (barcode_graph.py:224-249, index.py:77-93) run in ``libbadger_b200.so``.

Dedup / count in first-seen order, whitelist membership, the two rounds of ``cluster`` and ``--high_sens`` scoring run on
the device as well (``ops.dedup_first_seen``, ``ops.member_sorted``, ``ops.cluster_levels``, ``ops.nearest_bounded``); the host
keeps the centre-selection walk over the count-sorted barcodes and the dict-shaped views downstream readers expect.
"""
from __future__ import annotations

import logging
import time
from collections import defaultdict
from contextlib import contextmanager
from statistics import StatisticsError

import numpy as np

from . import ops
from .common import rank, unrank
from .index import QGramIndex

logger = logging.getLogger("BarcodeGraph")

READ_CHUNK_SIZE = 100000   # kept for interface parity (barcode_graph.py:25-26); unused by the GPU path
BC_CHUNK_SIZE = 10000


def _unrank_many(ranks: np.ndarray, bc_len: int = 16) -> list:
    ranks = np.asarray(ranks, dtype=np.uint32)
    codes = np.empty((ranks.size, bc_len), dtype=np.uint8)
    for i in range(bc_len):
        codes[:, i] = (ranks >> np.uint32(2 * i)) & np.uint32(3)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    return [s.decode() for s in letters.view("S%d" % bc_len).reshape(-1).tolist()]


class _EdgeView:
    """Read-only stand-in for ``defaultdict(list)`` rank -> neighbour ranks (barcode_graph.py:44) over CSR
    arrays.  A missing key yields an empty list and, like the defaultdict it replaces, is remembered as a key
    (badger.py:131 prints ``len(counts) - len(edges.keys())``)."""

    def __init__(self, nodes: np.ndarray, indptr: np.ndarray, nbrs: np.ndarray):
        self._nodes, self._indptr, self._nbrs = nodes, indptr, nbrs
        self._touched = set()

    def _slot(self, key):
        if not 0 <= int(key) <= 0xFFFFFFFF:
            return -1
        i = int(np.searchsorted(self._nodes, np.uint32(key)))     # typed key: a Python int would upcast the whole array per call
        return i if i < self._nodes.size and int(self._nodes[i]) == int(key) else -1

    def __getitem__(self, key):
        i = self._slot(key)
        if i < 0:
            self._touched.add(int(key))
            return []
        return self._nbrs[self._indptr[i]:self._indptr[i + 1]].tolist()

    def get(self, key, default=None):
        i = self._slot(key)
        return default if i < 0 else self._nbrs[self._indptr[i]:self._indptr[i + 1]].tolist()

    def __contains__(self, key):
        return self._slot(key) >= 0 or int(key) in self._touched

    def keys(self):
        extra = [k for k in self._touched if self._slot(k) < 0]
        return self._nodes.tolist() + extra

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def items(self):
        return ((k, self[k]) for k in self.keys())


class _DistView:
    """Stand-in for ``defaultdict(int)`` (a,b) -> D holding both orientations (barcode_graph.py:45,248-249)."""

    def __init__(self, a: np.ndarray, b: np.ndarray, d: np.ndarray):
        self._key = (a.astype(np.uint64) << np.uint64(32)) | b.astype(np.uint64)   # sorted by (a,b), a<b
        self._d = d

    def __getitem__(self, pair):
        x, y = int(pair[0]), int(pair[1])
        if x > y:
            x, y = y, x
        k = np.uint64((x << 32) | y)
        i = int(np.searchsorted(self._key, k))
        return int(self._d[i]) if i < self._key.size and self._key[i] == k else 0

    def __len__(self):
        return 2 * int(self._key.size)

    def keys(self):
        a = (self._key >> np.uint64(32)).tolist(); b = (self._key & np.uint64(0xFFFFFFFF)).tolist()
        return [(x, y) for x, y in zip(a, b)] + [(y, x) for x, y in zip(a, b)]

    def items(self):
        return ((k, self[k]) for k in self.keys())


def rest_by_counts(ranks, counts, cutoff, need):
    """bc_by_counts beyond the barcodes above the cutoff: the next `need` + 1 entries with count <= cutoff in count-descending
    order, ties by first sighting (the short stretch the top-up loop of barcode_graph.py:273-276 may reach)."""
    N = ranks.size
    rest = np.nonzero(counts <= cutoff)[0]
    m = min(rest.size, max(int(need) + 1, 1))
    key = -counts[rest] * np.int64(N) + rest                                # unique: count-descending, then first-seen
    pick = np.argpartition(key, m - 1)[:m] if m < rest.size else np.arange(rest.size)
    return ranks[rest[pick[np.argsort(key[pick], kind="stable")]]]


def walk_centres(N, n_cells, interval, top, hits, true_barcodes, have_list, rest_fn, bc_len=16):
    """The walk of barcode_graph.py:259-277 over the head of `bc_by_counts`: top = the barcodes with count > cutoff in that
    order, hits = their whitelist membership (None without a whitelist), rest_fn(need) = the entries that follow them, asked
    for only when the top-up loop runs past the cutoff.  Same list, same order, same IndexError when N is too small."""
    hi = n_cells + n_cells * interval * 0.01
    lo = n_cells - n_cells * interval * 0.01
    n_above = int(top.size)
    rest_sorted = []

    def by_counts_at(i, need):
        if i < n_above:
            return int(top[i])
        if not rest_sorted:
            rest_sorted.append(rest_fn(need))
        j = i - n_above
        if j >= rest_sorted[0].size:
            raise IndexError("list index out of range")                     # :274 in the reference
        return int(rest_sorted[0][j])

    tbcs, n, i = [], 0, 0
    if true_barcodes:
        tbcs = [bc if isinstance(bc, (int, np.integer)) else rank(bc, bc_len) for bc in true_barcodes]   # packed callers pass ranks
    elif have_list:
        csum = np.cumsum(hits)
        want = int(np.floor(hi)) + 1                                         # loop runs while n <= hi
        if csum.size and csum[-1] >= want:
            i = int(np.searchsorted(csum, want, side="left")) + 1
            n = want
        else:
            i = n_above
            n = int(csum[-1]) if csum.size else 0
        tbcs = top[:i][hits[:i]].tolist()
    else:
        if n_above >= N and N <= int(np.floor(hi)) + 1:
            raise IndexError("list index out of range")                     # :269 runs off the list
        n = i = min(n_above, int(np.floor(hi)) + 1)
        tbcs = top[:i].tolist()
    missing = int(np.ceil(lo - n)) if n < lo else 0
    while n < lo:
        if i >= N:
            raise IndexError("list index out of range")                     # :274 in the reference
        tbcs.append(by_counts_at(i, missing))
        i += 1
        n += 1
    return tbcs


class BarcodeGraph:

    def __init__(self, threshold):
        self.threshold = threshold
        self.counts = defaultdict(int)      # rank -> count, first-seen order (barcode_graph.py:43)
        self.edges = defaultdict(list)
        self.dists = defaultdict(int)
        self.clusters = defaultdict(list)
        self.clustering = dict()
        self.clustered = defaultdict(bool)
        self.index = None
        self.timings = defaultdict(float)          # seconds per stage of this graph's life (tools/pipeline_bench.py)
        self._ranks = np.empty(0, np.uint32)      # distinct ranks, first-seen order
        self._cnt = np.empty(0, np.int64)
        self._edge_arrays = (np.empty(0, np.uint32), np.empty(0, np.uint32), np.empty(0, np.uint8))

    @contextmanager
    def _timed(self, stage):
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.timings[stage] += time.perf_counter() - t0

    @classmethod
    def from_arrays(cls, threshold, ranks_first_seen, counts, edges=None, with_dict=True):
        """Array entry point: distinct ranks in first-seen order with their counts, optionally an edge list
        (a, b, d).  Used by bench.py / tests and by callers that already hold packed barcodes (with_dict=False skips
        the `counts` dict, which only the dict-shaped readers of the reference need)."""
        g = cls(threshold)
        g._ranks = np.ascontiguousarray(ranks_first_seen, dtype=np.uint32)
        g._cnt = np.ascontiguousarray(counts, dtype=np.int64)
        if with_dict:
            g.counts.update(zip(g._ranks.tolist(), g._cnt.tolist()))
        if edges is not None:
            g._set_edges(*edges)
        return g

    # ------------------------------------------------------------------ a-1/a-2: pack, dedup, count
    def index_bc_single_thread(self, barcodes, bc_len):
        """barcode_graph.py:192-204: 17-mers lose their last base, other lengths are skipped, the rest is
        ranked (GPU, ops.pack16) and counted in first-seen order."""
        if bc_len != 16:
            raise NotImplementedError("the B200 path packs 16-bp barcodes into uint32; bc_len=%r is not supported" % bc_len)
        with self._timed("index: length filter (python loop)"):
            keep = []
            for s in barcodes:
                n = len(s)
                if n == bc_len + 1:
                    keep.append(s[:-1])
                elif n == bc_len:
                    keep.append(s)
        if not keep:
            return
        with self._timed("index: pack16 (join + GPU)"):
            ranks, valid = ops.pack16(keep)
        if not valid.all():
            bad = keep[int(np.argmin(valid))]
            raise KeyError(next(c for c in bad if c not in "ACGT"))      # common.py:24 raises KeyError(letter)
        with self._timed("index: dedup/count first-seen (GPU)"):
            new_r, new_c = ops.dedup_first_seen(ranks)
        if self._ranks.size:                                               # merge with an earlier call
            for r, c in zip(new_r.tolist(), new_c.tolist()):
                self.counts[r] += c
            self._ranks = np.fromiter(self.counts.keys(), dtype=np.uint32, count=len(self.counts))
            self._cnt = np.fromiter(self.counts.values(), dtype=np.int64, count=len(self.counts))
        else:
            self._ranks, self._cnt = new_r, new_c
            with self._timed("index: counts dict"):
                self.counts.update(zip(new_r.tolist(), new_c.tolist()))

    index_bc_in_parallel = lambda self, barcodes, bc_len, threads: self.index_bc_single_thread(barcodes, bc_len)  # noqa: E731

    # ------------------------------------------------------------------ a-3/a-4: edges
    def graph_construction(self, barcodes, bc_len, threads):
        """barcode_graph.py:207-249.  ``threads`` is accepted for compatibility; the GPUs claimed by
        ``badger_b200.init`` replace the process pool."""
        self.index = QGramIndex(self.threshold, bc_len, 6)
        self.index_bc_single_thread(barcodes, bc_len)
        self.index._adopt(self._ranks)
        with self._timed("edges: sort + GPU edge construction"):
            a, b, d = ops.edges_build(np.sort(self._ranks), self.threshold)
        with self._timed("edges: canonical order + CSR views"):
            self._set_edges(a, b, d)

    def compare_in_parallel(self, bc_len, threads):
        a, b, d = ops.edges_build(np.sort(self._ranks), self.threshold)
        self._set_edges(a, b, d)

    def _set_edges(self, a, b, d):
        a, b, d = ops.canonical(np.asarray(a, np.uint32), np.asarray(b, np.uint32), np.asarray(d, np.uint8))
        self._edge_arrays = (a, b, d)
        src = np.concatenate([a, b]); dst = np.concatenate([b, a])
        order = np.lexsort((dst, src))
        src, dst = src[order], dst[order]
        nodes, start = np.unique(src, return_index=True)
        indptr = np.append(start, src.size).astype(np.int64)
        self._csr = (nodes, indptr, dst)
        self.edges = _EdgeView(nodes, indptr, dst)
        self.dists = _DistView(a, b, d)

    def edge_arrays(self):
        """(a, b, d) with a < b, sorted by (a, b): the array form of ``edges``/``dists``."""
        return self._edge_arrays

    # ------------------------------------------------------------------ a-6: centres
    def _whitelist_hits(self, barcode_list, bc_len, ranks=None):
        """`unrank(r, bc_len) in barcode_list` (barcode_graph.py:264) for every distinct barcode: the set is
        packed once, sorted, and probed on the GPU (ops.member_sorted)."""
        cache = getattr(self, "_wl_cache", None)
        if cache is None or cache[0] is not barcode_list:
            with self._timed("centres: pack + sort whitelist"):
                good = [s for s in barcode_list if len(s) == bc_len]       # other lengths can never equal an unranked barcode
                if good:
                    r, ok = ops.pack16(good)                                # entries with letters outside ACGT are flagged invalid
                    wl = np.sort(r[ok])
                else:
                    wl = np.empty(0, np.uint32)
            self._wl_cache = cache = (barcode_list, wl)
        return ops.member_sorted(cache[1], self._ranks if ranks is None else ranks)

    def get_cluster_centers(self, true_barcodes, bc_len, barcode_list, n_cells, interval):
        """barcode_graph.py:252-277; same list, same order, same IndexError when N is too small.  The reference sorts
        all barcodes by count (stable, so ties keep first-seen order) and walks the list; only the entries above the
        cutoff - a few thousand - and, rarely, a short stretch below it are ever looked at, so only those are sorted."""
        N = self._ranks.size
        if N == 0:
            raise StatisticsError("mean requires at least one data point")      # statistics.mean([]) at :255
        first = self._cnt[:n_cells]
        if first.size == 0:
            raise StatisticsError("mean requires at least one data point")
        cutoff = max((int(first.sum()) / first.size) / 5.0, 5)
        above = np.nonzero(self._cnt > cutoff)[0]                               # first-seen order
        above = above[np.argsort(-self._cnt[above], kind="stable")]             # count-descending, ties by first sighting
        top = self._ranks[above]                                                # bc_by_counts[:n_above]
        hits = None
        if not true_barcodes and barcode_list:
            hits = self._whitelist_hits(barcode_list, bc_len, top) if top.size else np.zeros(0, bool)
        return walk_centres(N, n_cells, interval, top, hits, true_barcodes, bool(barcode_list),
                            lambda need: rest_by_counts(self._ranks, self._cnt, cutoff, need), bc_len)

    # ------------------------------------------------------------------ clustering (GPU rounds, dict views on the host)
    def cluster(self, true_barcodes, barcode_list, n_cells, bc_len, interval):
        """barcode_graph.py:279-301.  The reference's two rounds are level-synchronous and independent of the
        adjacency order: a node joins centre c in round i iff every claim it receives in that round
        comes from c; two different centres in the same round evict it.  The rounds run edge-parallel on the GPU
        (ops.cluster_levels); this method fills the reference's dict attributes from the resulting arrays."""
        tbcs = self.get_cluster_centers(true_barcodes, bc_len, barcode_list, n_cells, interval)
        centres = list(dict.fromkeys(int(t) for t in tbcs))
        for t in centres:
            self.clusters[t] = [t]
            self.clustering[t] = (t, 0)
            self.clustered[t] = True
            _ = self.edges[t]                       # the reference touches edges[centre] (:293)
        print(1)                                     # barcode_graph.py:289 prints the round number
        print(2)
        a, b, _d = self._edge_arrays
        with self._timed("cluster: GPU rounds"):
            s = np.sort(self._ranks)
            ci, lv = ops.cluster_levels(s, a, b, np.asarray(centres, dtype=np.uint32), 2)
        self._cluster_arrays = (s, ci, lv)
        with self._timed("cluster: dict views"):
            got = np.nonzero((ci != -2) & (lv != 0))[0]
            nodes = s[got].tolist()
            cen = np.where(ci[got] >= 0, s[np.maximum(ci[got], 0)], 0).tolist()
            ok = (ci[got] >= 0).tolist()
            lvl = lv[got].tolist()
            for node, c, good, l in zip(nodes, cen, ok, lvl):
                if good:
                    self.clusters[c].append(node)
                    self.clustering[node] = (c, l)
                else:
                    self.clustering[node] = (-1, -1)
                self.clustered[node] = True

    # ------------------------------------------------------------------ assignment / post-processing / output
    def assign_by_cluster(self, bc_len):
        """barcode_graph.py:322-329 (dict built in `counts` order, which fixes the iteration order of
        ``set(assignments.values())`` used by postprocessing)."""
        observed_assignments = defaultdict(str)
        nodes = [n for n in self.counts.keys() if n in self.clustering and self.clustering[n][0] != -1]
        if nodes:
            bcs = _unrank_many(np.asarray(nodes, dtype=np.uint32), bc_len)
            cens = _unrank_many(np.asarray([self.clustering[n][0] for n in nodes], dtype=np.uint32), bc_len)
            for bc, tbc in zip(bcs, cens):
                observed_assignments[bc] = tbc
        return observed_assignments

    def postprocessing(self, assignments, bc_len, _centre_order=None):
        """barcode_graph.py:370-385 (--high_sens): every still-unassigned distinct barcode goes to the first
        centre, in the iteration order of ``set(assignments.values())``, at minimum plain edit distance when
        that distance is < 3.  The Q x W scoring runs on the GPU (ops.nearest_bounded)."""
        cluster_centers = list(set(assignments.values())) if _centre_order is None else list(_centre_order)
        all_bc = _unrank_many(self._ranks, bc_len)
        todo_idx = [i for i, bc in enumerate(all_bc) if assignments[bc] == "" or assignments[bc] == "*"]
        good = [c for c in cluster_centers if len(c) == bc_len and not (set(c) - set("ACGT"))]
        if not todo_idx or not good:
            return assignments
        pos = [j for j, c in enumerate(cluster_centers) if len(c) == bc_len and not (set(c) - set("ACGT"))]
        targets = ops.pack16(good)[0]
        am, _ = ops.nearest_bounded(self._ranks[todo_idx], targets, 2)
        for i, j in zip(todo_idx, am.tolist()):
            if j >= 0:
                assignments[all_bc[i]] = cluster_centers[pos[j]]
        return assignments

    def output_file(self, read_assignment, out, true_barcodes, bc_len, post):
        """barcode_graph.py:388-410: `<out>_output_file.tsv`, header readID/barcode, '*' when unassigned."""
        import pandas as pd
        assignments = self.assign_by_cluster(bc_len)
        if post:
            assignments = self.postprocessing(assignments, bc_len)
        read_ids, results = [], []
        for read in read_assignment:
            observed_bc = read[1]
            assigned_bc = "*"
            if observed_bc != "*":
                assigned_bc = assignments[observed_bc]
                if assigned_bc == "":
                    assigned_bc = "*"
            read_ids.append(read[0])
            results.append(assigned_bc)
        out_file = out + "_output_file.tsv"
        res = pd.DataFrame({"readID": read_ids, "barcode": results})
        res.to_csv(out_file, sep='\t', index=False)
