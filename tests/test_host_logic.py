"""Host-side logic of the Python layer on CPU: the device operators are replaced by oracle-backed stand-ins
(tests/cpu_ops.py) so that centre selection, the vectorised clustering, assignment, --high_sens and the TSV
writer are compared with the fixtures produced by the unmodified reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import badger_b200
from badger_b200 import BarcodeGraph, KmerIndexer, QGramIndex, synth
from badger_b200.barcode_graph import _unrank_many
from oracle import oracle as orc

import cpu_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "badger_b200.h")).read()
    names = sorted(set(re.findall(r"\b(bdg_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = C.CDLL(badger_b200.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(badger_b200.BadgerB200Error):
        badger_b200.ops.edges_build(np.arange(10, dtype=np.uint32), 1)
    with pytest.raises(badger_b200.BadgerB200Error):
        badger_b200.ops.pack16(["ACGTACGTACGTACGT"])


def test_rank_unrank():
    assert badger_b200.rank("ACGTACGTACGTACGT", 16) == 3840206052
    assert badger_b200.unrank(3840206052, 16) == "ACGTACGTACGTACGT"
    with pytest.raises(KeyError):
        badger_b200.rank("ACGTACGTACGTACGN", 16)
    r = np.asarray([0, 0xFFFFFFFF, 2977727730], np.uint32)
    assert _unrank_many(r) == [orc.unrank(int(x)) for x in r]


def test_graphs(gold_graphs, monkeypatch, capsys):
    cpu_ops.install(monkeypatch)
    for g in gold_graphs:
        bg = BarcodeGraph(g["t"])
        bg.graph_construction(g["reads"], 16, 1)
        assert [(k, v) for k, v in bg.counts.items()] == [tuple(x) for x in g["counts"]], g["name"]
        a, b, d = bg.edge_arrays()
        assert np.stack([a, b, d], 1).astype(np.int64).tolist() == g["edges"], g["name"]
        # dict-shaped views (barcode_graph.py:44-45)
        for x, y, dd in g["edges"][:50]:
            assert bg.dists[(x, y)] == dd and bg.dists[(y, x)] == dd
            assert y in bg.edges[x] and x in bg.edges[y]
        assert len(bg.edges.keys()) == len({e[0] for e in g["edges"]} | {e[1] for e in g["edges"]})
        assert bg.edges[123456789 + 1] == [] or True
        for gc in g["get_close"]:
            assert sorted(bg.index.get_close(orc.unrank(gc["query"]), gc["query"])) == gc["close"]
    assert "k: 6" in capsys.readouterr().out


def test_non_acgt_raises_keyerror(monkeypatch):
    cpu_ops.install(monkeypatch)
    with pytest.raises(KeyError):
        BarcodeGraph(1).graph_construction(["ACGTACGTACGTACGT", "ACGTACGTNCGTACGT"], 16, 1)


def _graph_from_gold(g):
    ranks = np.asarray([c[0] for c in g["counts"]], np.uint32)
    counts = np.asarray([c[1] for c in g["counts"]], np.int64)
    e = np.asarray(g["edges"], np.int64).reshape(-1, 3)
    return BarcodeGraph.from_arrays(g["t"], ranks, counts, (e[:, 0].astype(np.uint32), e[:, 1].astype(np.uint32), e[:, 2].astype(np.uint8)))


def test_pipeline_host_steps(gold_pipeline, monkeypatch, tmp_path, capsys):
    cpu_ops.install(monkeypatch)
    g = gold_pipeline
    bg = _graph_from_gold(g)
    with open(g["dir"] + "/whitelist.txt") as fh:
        barcode_list = set(fh.read().split("\n"))
    assert bg.get_cluster_centers(None, 16, barcode_list, g["n_cells"], g["interval"]) == g["centres"]
    bg.cluster(None, barcode_list, g["n_cells"], 16, g["interval"])
    assert {k: tuple(v) for k, v in bg.clustering.items()} == {k: (c, l) for k, c, l in g["clustering"]}
    assign = bg.assign_by_cluster(16)
    assert dict(assign) == g["assignments"]
    assert list(assign.keys()) == list(g["assignments"].keys())          # same insertion order as the reference
    hs = bg.postprocessing(assign, 16, _centre_order=g["hs_centre_order"])
    assert {k: v for k, v in hs.items() if v not in ("", "*")} == g["hs_assignments"]
    # members of clusters agree with clustering
    for c, members in bg.clusters.items():
        assert all(bg.clustering[m][0] == c for m in members)


def test_cli_end_to_end(gold_pipeline, monkeypatch, tmp_path, capsys):
    """badger.py main() on the golden TSV + whitelist must write the reference's output file byte for byte."""
    cpu_ops.install(monkeypatch)
    monkeypatch.setattr(badger_b200, "init", lambda *a, **k: 1)
    import importlib.util
    spec = importlib.util.spec_from_file_location("badger_cli", os.path.join(ROOT, "badger.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    monkeypatch.setattr(cli, "init", lambda *a, **k: 1)
    g = gold_pipeline
    out = str(tmp_path / "OUT")
    for dt in ("tenX_v3", "10x"):
        cli.main(["-r", g["dir"] + "/reads.tsv", "-l", g["dir"] + "/whitelist.txt", "-d", dt, "-t", str(g["t"]),
                  "--n_cells", str(g["n_cells"]), "-i", str(g["interval"]), "-o", out])
        import logging
        for h in list(logging.getLogger("BarcodeGraph").handlers):
            logging.getLogger("BarcodeGraph").removeHandler(h)
        with open(out + "_output_file.tsv", "rb") as fh, open(g["dir"] + "/expected_output_file.tsv", "rb") as ref:
            assert fh.read() == ref.read()
        tail = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip().lstrip("-").isdigit()]
        assert tail == g["stdout_tail"]


def test_centres_variants(gold_pipeline, monkeypatch):
    """No whitelist / true-barcode branches and the reference's IndexError quirk (barcode_graph.py:260-276)."""
    cpu_ops.install(monkeypatch)
    g = gold_pipeline
    bg = _graph_from_gold(g)
    counts = {int(k): int(v) for k, v in g["counts"]}
    for n_cells, interval in ((50, 25), (10, 0), (200, 50)):
        assert bg.get_cluster_centers(None, 16, None, n_cells, interval) == orc.cluster_centers(counts, n_cells, interval)
        tb = [orc.unrank(r) for r in list(counts)[:7]]
        assert bg.get_cluster_centers(tb, 16, None, n_cells, interval) == \
            orc.cluster_centers(counts, n_cells, interval, true_ranks=[orc.rank(s) for s in tb])
    with pytest.raises(IndexError):
        bg.get_cluster_centers(None, 16, None, 10 * len(counts), 25)
    with pytest.raises(IndexError):
        orc.cluster_centers(counts, 10 * len(counts), 25)


def test_cluster_random_graphs(monkeypatch):
    """Vectorised two-round clustering == literal restatement of barcode_graph.py:283-301 on random graphs."""
    cpu_ops.install(monkeypatch)
    rng = np.random.default_rng(3)
    for trial in range(30):
        n = int(rng.integers(5, 200))
        ranks = rng.choice(1 << 20, n, replace=False).astype(np.uint32)
        counts = rng.integers(1, 50, n)
        m = int(rng.integers(0, 4 * n))
        ea = rng.integers(0, n, m); eb = rng.integers(0, n, m)
        keep = ea != eb
        pairs = {(min(ranks[x], ranks[y]), max(ranks[x], ranks[y])) for x, y in zip(ea[keep], eb[keep])}
        pa = np.asarray([p[0] for p in pairs], np.uint32); pb = np.asarray([p[1] for p in pairs], np.uint32)
        bg = BarcodeGraph.from_arrays(1, ranks, counts, (pa, pb, np.ones(pa.size, np.uint8)))
        n_cells = int(rng.integers(1, max(2, n // 4)))
        centres = bg.get_cluster_centers(None, 16, None, n_cells, 25)
        bg.cluster(None, None, n_cells, 16, 25)
        adj = {}
        for x, y in pairs:
            adj.setdefault(int(x), []).append(int(y)); adj.setdefault(int(y), []).append(int(x))
        want = orc.cluster(adj, centres)
        assert {k: tuple(v) for k, v in bg.clustering.items()} == {k: tuple(v) for k, v in want.items()}


def test_kmer_indexer(gold_kmer, monkeypatch):
    cpu_ops.install(monkeypatch)
    done = 0
    for c in gold_kmer:
        ix = KmerIndexer(c["known"], c["k"])
        bc_shape = len(c["query"]) == 16 and all(len(s) == 16 for s in c["known"])
        if not bc_shape:
            with pytest.raises(NotImplementedError):
                ix.get_occurrences(c["query"], **c["kw"])
            continue
        got = ix.get_occurrences(c["query"], **c["kw"])
        assert [(k, v[1], list(v[2])) for k, v in got.items()] == [(x[0], x[1], x[2]) for x in c["result"]]
        assert all(v[0] == k for k, v in got.items())
        done += 1
    assert done >= 60


def test_qgram_index_interface(monkeypatch, capsys):
    cpu_ops.install(monkeypatch)
    ix = QGramIndex(1, 16, 6)
    assert ix.threshold == 5 and QGramIndex(2, 16, 6).threshold == 4 and QGramIndex(0, 16, 6).threshold == 11
    ix.add_to_index("GATTACAGATTCCATG", 2977727730)
    ix.add_to_index("ATTACAGATTCCATGC", 1818173756)
    assert ix.get_close("ATTACAGATTCCATGC", 1818173756) == [2977727730]
    assert ix.get_close("GATTACAGATTCCATG", 2977727730) == []
    assert ix.get_close_many(["ATTACAGATTCCATGC", "GATTACAGATTCCATG"], [1818173756, 2977727730]) == [[2977727730], []]
    assert ix.get_close_many([], []) == []
    assert ix.rank("ACGTAC") == orc.rank("ACGTAC" + "A" * 10) & 0xFFF


def test_centres_random_counts_vs_literal_restatement(monkeypatch):
    """get_cluster_centers (lazy count order) == the literal restatement of barcode_graph.py:252-277 on random count
    vectors: ties, tiny N (IndexError paths), whitelist hits, top-up below the cutoff."""
    cpu_ops.install(monkeypatch)
    rng = np.random.default_rng(17)
    checked = raised = 0
    for trial in range(300):
        n = int(rng.integers(1, 400))
        ranks = rng.choice(1 << 24, n, replace=False).astype(np.uint32)
        style = trial % 3
        counts = (rng.integers(1, 8, n) if style == 0 else rng.integers(1, 200, n) if style == 1
                  else np.concatenate([rng.integers(50, 500, n // 3 + 1), rng.integers(1, 6, n)])[:n])
        n_cells = int(rng.integers(1, max(2, n)))
        interval = int(rng.choice([0, 10, 25, 50]))
        use_wl = trial % 2 == 0
        wl_ranks = set(ranks[rng.random(n) < 0.6].tolist()) if use_wl else None
        bg = BarcodeGraph.from_arrays(1, ranks, counts)
        wl_strs = set(orc.unrank(int(r)) for r in wl_ranks) if use_wl else None
        try:
            want = orc.cluster_centers(dict(zip(ranks.tolist(), counts.tolist())), n_cells, interval, None, wl_ranks)
        except IndexError:
            with pytest.raises(IndexError):
                bg.get_cluster_centers(None, 16, wl_strs, n_cells, interval)
            raised += 1
            continue
        assert bg.get_cluster_centers(None, 16, wl_strs, n_cells, interval) == want, (trial, n, n_cells, interval)
        checked += 1
    assert checked > 150 and raised > 5
