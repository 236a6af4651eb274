"""Array form of the correction step (reference badger.py:120-129 with barcode_graph.py:192-410 behind it): packed
barcodes per read in, cluster-centre barcode per read out, every stage on the GPU operators, no per-read Python.

    read barcodes (uint32 per row + valid mask)
      -> ops.dedup_reads                 barcode_graph.py:192-204   valid rows compacted on the device; distinct barcodes in
                                                                    first-seen order, counts; the read map STAYS on the device
      -> ops.edges_handle                index.py:77-93 + barcode_graph.py:224-249   every claimed GPU builds its part; edges stay there
      -> BarcodeGraph.get_cluster_centers   barcode_graph.py:252-277   (count-sorted scan, whitelist membership on the GPU)
      -> EdgeHandle.cluster_levels       barcode_graph.py:279-301   two rounds, same-round conflicts evict (parts gathered over NVLink)
      -> ops.nearest_bounded             barcode_graph.py:370-385   only with high_sens; patches the node -> centre array
      -> ops.assign_reads                barcode_graph.py:322-329,395-404   centre of every row: gather kernel over the resident read map

The string-based mirror (``BarcodeGraph`` + ``badger.py``) gives the same assignments; this module exists for callers
that already hold packed barcodes and for the reads/s figure of ``bench.py``.
"""
from __future__ import annotations

import time

import numpy as np

from . import ops
from .barcode_graph import BarcodeGraph, _unrank_many
from .common import rank

NONE = np.uint64(1) << np.uint64(32)      # "no centre" marker in the uint64 result (every uint32 is a valid barcode)


class _Token:
    """Stands in for the whitelist set the string API passes around (truthy, compared by identity)."""

    def __bool__(self):
        return True


def assign_packed(ranks, valid=None, *, threshold, n_cells, interval=25, whitelist_sorted=None, true_barcodes=None,
                  high_sens=False, centre_order=None, timings=None):
    """Returns (centre uint64[R], info).  centre[i] == pipeline.NONE where the reference would write ``*``.

    ranks / valid: packed barcode and validity per read (valid=None: all valid); whitelist_sorted: ascending uint32 array
    or None; true_barcodes: iterable of packed centres or None (badger.py --true_barcodes); centre_order: order in which
    --high_sens tries the centres (the reference iterates a Python set, barcode_graph.py:372; default: ascending; "set":
    the iteration order of that very set of strings in this process, which is what badger_b200.BarcodeGraph meets too)."""
    T = timings if timings is not None else {}

    def tick(name, t0):
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0

    ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
    R = ranks.size
    if valid is not None:
        valid = np.ascontiguousarray(valid, dtype=bool)
    info = {"reads": int(R)}
    if R == 0:
        info["valid_reads"] = 0
        return np.full(0, NONE, dtype=np.uint64), info

    t0 = time.perf_counter()
    rm = ops.dedup_reads(ranks, valid)                    # valid rows compacted on the device, read map stays there
    distinct, counts, spos = rm.distinct, rm.counts, rm.sorted_pos
    tick("dedup_first_seen", t0)
    info["valid_reads"] = int(rm.n_valid)
    if rm.n_valid == 0:
        return np.full(R, NONE, dtype=np.uint64), info
    t0 = time.perf_counter()
    s = rm.sorted_distinct                                # ascending order straight from the dedup's sort
    handle = ops.edges_handle_resident(rm, threshold)     # the array is already on the device: no upload, peer copies to the other GPUs
    tick("edges", t0)
    info.update(distinct=int(distinct.size), edges=int(handle.count))

    t0 = time.perf_counter()
    g = BarcodeGraph.from_arrays(threshold, distinct, counts, with_dict=False)
    token = None
    if whitelist_sorted is not None:
        token = _Token()
        g._wl_cache = (token, np.ascontiguousarray(whitelist_sorted, dtype=np.uint32))
    tb = None if true_barcodes is None else [int(x) for x in true_barcodes]
    centres = np.asarray(list(dict.fromkeys(g.get_cluster_centers(tb, 16, token, n_cells, interval))), dtype=np.uint32)
    tick("centres", t0)
    info["centres"] = int(centres.size)

    t0 = time.perf_counter()
    ci, lv, has_edge = handle.cluster_levels(centres, 2, want_has_edge=True)   # consumes the handle's edges (no copy to the host)
    handle.free()
    # badger.py:131 `len(counts) - len(edges.keys())`: the keys are the nodes with an edge plus every centre cluster() touched
    info["disconnected"] = int(distinct.size) - (int(has_edge.sum()) + int(centres.size))
    tick("cluster", t0)

    if high_sens:
        # barcode_graph.py:370-385 on node positions: every node without a centre looks for the nearest used centre
        t0 = time.perf_counter()
        todo = np.nonzero(ci < 0)[0]
        used_nodes = ops.sorted_unique(ci[ci >= 0])                              # set(assignments.values()), as nodes (ascending = by value)
        if centre_order is None:
            targets = s[used_nodes]
        elif isinstance(centre_order, str) and centre_order == "set":
            # the order the string route meets in this process: `set(assignments.values())` with the dict filled in
            # first-seen order of the distinct barcodes (barcode_graph.py:322-329,372).  Re-adding a member leaves a set's
            # table untouched, so the set of the first occurrences, added in that order, iterates identically.
            cfs = ci[spos]                                                       # centre node per distinct barcode, first-seen order
            vals = cfs[cfs >= 0]
            order = np.argsort(vals, kind="stable")                              # first occurrence of every value, in order of appearance
            sv = vals[order]
            head = np.ones(sv.size, bool)
            head[1:] = sv[1:] != sv[:-1]
            strs = _unrank_many(s[vals[np.sort(order[head])]])
            targets = np.asarray([rank(c, 16) for c in set(strs)], np.uint32)
        else:
            used = set(s[used_nodes].tolist())
            targets = np.asarray([c for c in centre_order if c in used], np.uint32)
        if todo.size and targets.size:
            am, _ = ops.nearest_bounded(s[todo], targets, 2)
            hit = am >= 0
            ci[todo[hit]] = np.searchsorted(s, targets[am[hit]]).astype(np.int32)
        tick("high_sens", t0)

    t0 = time.perf_counter()
    out, n_assigned = ops.assign_reads(rm, ci)            # per-read gather on the device
    tick("gather", t0)
    info["assigned_reads"] = int(n_assigned)
    return out, info
