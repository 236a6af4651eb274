"""Array form of the correction step (reference badger.py:120-129 with barcode_graph.py:192-410 behind it): packed
barcodes per read in, cluster-centre barcode per read out, every stage on the GPU operators, no per-read Python.

    read barcodes (uint32 per row + valid mask)
      -> ops.dedup_reads                 barcode_graph.py:192-204   valid rows compacted on the device; distinct barcodes in
                                                                    first-seen order, counts; the read map STAYS on the device
      -> ops.edges_handle                index.py:77-93 + barcode_graph.py:224-249   every claimed GPU builds its part; edges stay there
      -> ops.centres_above + walk_centres   barcode_graph.py:252-277   cutoff, count-ordered head of the list and its whitelist hits on the GPU
      -> EdgeHandle.cluster_resident     barcode_graph.py:279-301   two rounds, same-round conflicts evict (parts gathered over NVLink)
      -> ops.nearest_bounded             barcode_graph.py:370-385   only with high_sens; patches the node -> centre array
      -> ops.assign_reads32              barcode_graph.py:322-329,395-404   centre of every row: gather kernel over the resident read map

The string-based mirror (``BarcodeGraph`` + ``badger.py``) gives the same assignments; this module exists for callers
that already hold packed barcodes and for the reads/s figure of ``bench.py``.
"""
from __future__ import annotations

import time

import numpy as np

from . import ops
from .barcode_graph import _unrank_many
from .common import rank

NONE = np.uint64(1) << np.uint64(32)      # "no centre" marker in the uint64 result (every uint32 is a valid barcode)


class _Token:
    """Stands in for the whitelist set the string API passes around (truthy, compared by identity)."""

    def __bool__(self):
        return True


def select_centres(rm, n_cells, interval, whitelist_sorted, true_barcodes):
    """barcode_graph.py:252-277 with the counting, the count-ordered head of the barcode list and its whitelist membership on
    the device (ops.centres_above, and ops.centres_rest if the top-up loop has to go below the cutoff): no per-barcode array is
    downloaded."""
    from statistics import StatisticsError
    from .barcode_graph import walk_centres
    if rm.n_distinct == 0 or n_cells <= 0:
        raise StatisticsError("mean requires at least one data point")
    tb = None if true_barcodes is None else [int(x) for x in true_barcodes]
    cutoff, top, _, hits = ops.centres_above(rm, n_cells, None if tb else whitelist_sorted)
    have_list = whitelist_sorted is not None
    if hits is None and have_list and not tb:
        hits = np.zeros(0, bool)
    return walk_centres(rm.n_distinct, n_cells, interval, top, hits, tb, have_list,
                        lambda need: ops.centres_rest(rm, cutoff, int(need) + 1))


def assign_packed(ranks, valid=None, *, threshold, n_cells, interval=25, whitelist_sorted=None, true_barcodes=None,
                  high_sens=False, centre_order=None, timings=None, form="u64"):
    """Returns (centre uint64[R], info) - centre[i] == pipeline.NONE where the reference would write ``*`` - or, with
    form="u32", ((centre uint32[R], has_centre uint8[R]), info): 5 instead of 8 bytes per row cross PCIe.

    ranks / valid: packed barcode and validity per read (valid=None: all valid); whitelist_sorted: ascending uint32 array
    or None; true_barcodes: iterable of packed centres or None (badger.py --true_barcodes); centre_order: order in which
    --high_sens tries the centres (the reference iterates a Python set, barcode_graph.py:372; default: ascending; "set":
    the iteration order of that very set of strings in this process, which is what badger_b200.BarcodeGraph meets too).

    Without high_sens nothing per-barcode ever reaches the host: the distinct barcodes, their counts, the edges, the
    clustering result and the read map stay on the device between the stages."""
    T = timings if timings is not None else {}

    def tick(name, t0):
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0

    def result(c32, has):
        if form == "u32":
            return c32, has
        out = c32.astype(np.uint64)
        out[has == 0] = NONE
        return out

    ranks = np.ascontiguousarray(ranks, dtype=np.uint32)
    R = ranks.size
    if valid is not None:
        valid = np.ascontiguousarray(valid, dtype=bool)
    info = {"reads": int(R)}
    if R == 0:
        info["valid_reads"] = 0
        return result(np.zeros(0, np.uint32), np.zeros(0, np.uint8)), info

    t0 = time.perf_counter()
    rm = ops.dedup_reads(ranks, valid)                    # valid rows compacted on the device; everything stays there
    tick("dedup_first_seen", t0)
    info["valid_reads"] = int(rm.n_valid)
    if rm.n_valid == 0:
        return result(np.zeros(R, np.uint32), np.zeros(R, np.uint8)), info
    t0 = time.perf_counter()
    handle = ops.edges_handle_resident(rm, threshold)     # the ascending array is already on the device: no upload, peer copies to the other GPUs
    tick("edges", t0)
    info.update(distinct=int(rm.n_distinct), edges=int(handle.count))

    t0 = time.perf_counter()
    centres = np.asarray(list(dict.fromkeys(select_centres(rm, n_cells, interval, whitelist_sorted, true_barcodes))), dtype=np.uint32)
    tick("centres", t0)
    info["centres"] = int(centres.size)

    if not high_sens:
        t0 = time.perf_counter()
        n_has_edge = handle.cluster_resident(centres, 2)  # consumes the handle's edges; the node -> centre array stays on the device
        handle.free()
        # badger.py:131 `len(counts) - len(edges.keys())`: the keys are the nodes with an edge plus every centre cluster() touched
        info["disconnected"] = int(rm.n_distinct) - (int(n_has_edge) + int(centres.size))
        tick("cluster", t0)
        t0 = time.perf_counter()
        c32, has, n_assigned = ops.assign_reads32(rm, None)
        tick("gather", t0)
        info["assigned_reads"] = int(n_assigned)
        return result(c32, has), info

    t0 = time.perf_counter()
    s, spos = rm.sorted_distinct, rm.sorted_pos
    ci, lv, has_edge = handle.cluster_levels(centres, 2, want_has_edge=True)   # consumes the handle's edges (no copy to the host)
    handle.free()
    info["disconnected"] = int(rm.n_distinct) - (int(has_edge.sum()) + int(centres.size))
    tick("cluster", t0)
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
    c32, has, n_assigned = ops.assign_reads32(rm, ci)     # per-read gather on the device
    tick("gather", t0)
    info["assigned_reads"] = int(n_assigned)
    return result(c32, has), info
