"""Run the UNMODIFIED reference (/root/reference) in this container.  TEST INFRASTRUCTURE ONLY.

The reference is pure Python but leans on pip packages that are not installed here
(editdistance, ssw, pysam, Bio, matplotlib, igraph, edlib, Levenshtein).  ``oracle/stubs``
provides empty stand-ins plus a real Levenshtein for ``editdistance.eval`` (mathematically
pinned); with those on ``sys.path`` the reference imports and runs as is.

Nothing in the product (``badger_b200/``), ``bench.py`` or the ``-m gpu`` tests imports this
module: /root/reference does not exist on the GPU box.  It is used by
``oracle/make_golden.py`` (to write ``tests/golden/``) and by the ``not gpu`` tests that
re-check the C restatement against the live reference when the reference is present.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

REFERENCE_DIR = os.environ.get("BADGER_REFERENCE_DIR", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "barcode_graph.py"))


def load():
    """Import the reference's modules (once) and return them in a dict."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_DIR)
    # The reference's process-pool variant pickles BarcodeGraph by module name
    # (barcode_graph.py:142-143,177-178), so the modules stay importable under their own names
    # and the two directories stay on sys.path for the life of this (test-only) process.
    for k in ("common", "index", "barcode_graph", "stats", "badger", "extract_raw_barcodes"):
        sys.modules.pop(k, None)
    for k in [k for k in sys.modules if k == "barcode_extraction" or k.startswith("barcode_extraction.")]:
        sys.modules.pop(k, None)
    sys.path[:0] = [_STUBS, REFERENCE_DIR]
    with contextlib.redirect_stdout(io.StringIO()):     # index.py:21 prints "k: 6" at import
        mods = {name: importlib.import_module(name) for name in
                ("common", "index", "barcode_graph", "badger")}
        mods["kmer_indexer"] = importlib.import_module("barcode_extraction.kmer_indexer")
    _mods = mods
    return mods


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ref_graph(barcodes, threshold: int, bc_len: int = 16, threads: int = 1):
    """barcode_graph.py:207-249 -> (graph object, counts [(rank,count)...] in first-seen order,
    sorted edge list [(a,b,d)] with a<b)."""
    m = load()
    with _quiet():
        g = m["barcode_graph"].BarcodeGraph(threshold)
        g.graph_construction(list(barcodes), bc_len, threads)
    counts = [(int(k), int(v)) for k, v in g.counts.items()]
    edges = sorted((int(a), int(b), int(d)) for (a, b), d in g.dists.items() if a < b)
    # adjacency and dists must describe the same undirected edge set
    adj = set()
    for a, nb in g.edges.items():
        for b in nb:
            adj.add((min(a, b), max(a, b)))
    assert adj == {(a, b) for a, b, _ in edges}
    return g, counts, edges


def ref_pair(a: str, b: str):
    """ed, D (barcode_graph.py:243) and S (index.py:77-93 accumulated multiplicities) for one pair."""
    m = load()
    import editdistance  # the stub, already on sys.modules via load()
    ed = editdistance.eval(a, b)
    D = min(ed, editdistance.eval(a[:-1], b), editdistance.eval(a, b[:-1]))
    with _quiet():
        ix = m["index"].QGramIndex(1, 16, 6)
    ix.add_to_index(b, 1)
    kmer = ix.rank(a[:6])
    S = ix.index[kmer].get(1, 0)
    for i in range(6, len(a)):
        kmer = ix.update_rank(kmer, a[i])
        S += ix.index[kmer].get(1, 0)
    return ed, D, S


def ref_threshold(t: int) -> int:
    m = load()
    with _quiet():
        return m["index"].QGramIndex(t, 16, 6).threshold


def ref_get_close(ranks, query_rank: int, threshold: int):
    """QGramIndex.get_close over an index holding ``ranks`` (index.py:29-35,77-93)."""
    m = load()
    unrank = m["common"].unrank
    with _quiet():
        ix = m["index"].QGramIndex(threshold, 16, 6)
    for r in ranks:
        ix.add_to_index(unrank(int(r), 16), int(r))
    return sorted(int(x) for x in ix.get_close(unrank(int(query_rank), 16), int(query_rank)))


def ref_cluster(g, true_barcodes, barcode_list, n_cells: int, interval: int, bc_len: int = 16):
    """barcode_graph.py:252-301 on a constructed graph -> (centres, clustering dict)."""
    m = load()
    with _quiet():
        g2_centres = g.get_cluster_centers(true_barcodes, bc_len, barcode_list, n_cells, interval)
        g.cluster(true_barcodes, barcode_list, n_cells, bc_len, interval)
    return [int(x) for x in g2_centres], {int(k): (int(v[0]), int(v[1])) for k, v in g.clustering.items()}


def ref_main(argv):
    """Run badger.main(argv) (badger.py:62-175); returns captured stdout."""
    m = load()
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            m["badger"].main(list(argv))
    finally:
        import logging
        lg = logging.getLogger("BarcodeGraph")
        for h in list(lg.handlers):
            lg.removeHandler(h)
    return buf.getvalue()


def ref_get_occurrences(known, query, kmer_size=6, array=False, **kw):
    m = load()
    cls = m["kmer_indexer"].ArrayKmerIndexer if array else m["kmer_indexer"].KmerIndexer
    ix = cls(list(known), kmer_size)
    res = ix.get_occurrences(query, **kw)
    return [(k, int(v[1]), [int(p) for p in v[2]]) for k, v in res.items()]
