"""The CPU oracle (oracle/) against the fixtures produced by the unmodified reference
(oracle/make_golden.py).  This is what "pins" the oracle (the reference has no
tests or golden vectors of its own)."""
import hashlib

import numpy as np

from badger_b200 import synth
from oracle import oracle as orc


def test_rank_unrank_kat():
    # common.py:21-38; KAT from SURVEY.md §8(a-1)
    assert orc.rank("ACGTACGTACGTACGT") == 3840206052
    assert orc.unrank(3840206052) == "ACGTACGTACGTACGT"
    assert orc.rank("GATTACAGATTCCATG") == 2977727730
    assert orc.rank("ATTACAGATTCCATGC") == 1818173756
    try:
        orc.rank("ACGTACGTACGTACGN")
        assert False
    except KeyError:
        pass


def test_thresholds(gold_pairs):
    for t, T in gold_pairs["thresholds"].items():
        assert orc.T(int(t)) == T


def test_pairs(gold_pairs):
    for p in gold_pairs["pairs"]:
        ra, rb = orc.rank(p["a"]), orc.rank(p["b"])
        assert (ra, rb) == (p["ra"], p["rb"])
        assert orc.unrank(ra) == p["a"]
        assert orc.ed(ra, rb) == p["ed"]
        assert orc.D(ra, rb) == p["D"] == orc.D(rb, ra)
        assert orc.S(ra, rb) == p["S"] == orc.S(rb, ra)
        for t in range(4):
            want = p["D"] if (p["S"] >= orc.T(t) and p["D"] <= t) else -1
            assert orc.edge(ra, rb, t) == want


def _prep(reads):
    """barcode_graph.py:195-197 length rules, then pack."""
    keep = []
    for s in reads:
        if len(s) == 17:
            s = s[:-1]
        if len(s) == 16:
            keep.append(s)
    return synth.rank_many(keep) if keep else np.empty(0, np.uint32)


def test_graphs(gold_graphs):
    for g in gold_graphs:
        packed = _prep(g["reads"])
        ranks, counts = orc.dedup_count(packed)
        assert [(int(r), int(c)) for r, c in zip(ranks, counts)] == [tuple(x) for x in g["counts"]], g["name"]
        want = np.asarray(g["edges"], dtype=np.int64).reshape(-1, 3)
        ix = orc.Index(ranks)
        a, b, d, _ = ix.edges(g["t"])
        assert np.array_equal(np.stack([a, b, d], 1).astype(np.int64), want), g["name"]
        a2, b2, d2 = orc.edges_brute(ranks, g["t"])
        assert np.array_equal(np.stack([a2, b2, d2], 1).astype(np.int64), want), g["name"]
        rl = ranks.tolist()
        for gc in g["get_close"]:
            got = ix.get_close(rl.index(gc["query"]), g["t"]).tolist()
            assert got == gc["close"], g["name"]


def test_get_occurrences(gold_kmer):
    for c in gold_kmer:
        got = orc.get_occurrences(c["known"], c["query"], c["k"], **c["kw"])
        assert [(s, n, list(p)) for s, n, p in got] == [tuple(x) if not isinstance(x, list) else (x[0], x[1], x[2]) for x in c["result"]]


def test_kmer_score_matches_get_occurrences(gold_kmer):
    for c in gold_kmer:
        if any(len(s) != 16 for s in c["known"]) or len(c["query"]) != 16:
            continue
        q = synth.rank_many([c["query"]]); wl = synth.rank_many(c["known"])
        cnt, mult = orc.kmer_score(q, wl)
        ref = orc.get_occurrences(c["known"], c["query"], 6, max_hits=0, min_kmers=1, hits_delta=1000)
        by_str = {s: (n, p) for s, n, p in ref}
        for j, s in enumerate(c["known"]):
            n = int(cnt[0, j])
            if n == 0:
                assert s not in by_str
            else:
                pos = [p for p in range(11) for _ in range(int(mult[0, j, p]))]
                assert by_str[s] == (n, pos)


def test_pipeline(gold_pipeline):
    g = gold_pipeline
    counts = {int(k): int(v) for k, v in g["counts"]}
    edges = np.asarray(g["edges"], dtype=np.int64).reshape(-1, 3)
    ranks = np.fromiter(counts.keys(), dtype=np.uint32)
    a, b, d, _ = orc.Index(ranks).edges(g["t"])
    assert np.array_equal(np.stack([a, b, d], 1).astype(np.int64), edges)
    with open(g["dir"] + "/whitelist.txt") as fh:
        wl = set(fh.read().split("\n"))
    wl_ranks = np.sort(synth.rank_many([s for s in wl if s]))
    hit = orc.member(wl_ranks, ranks)
    wl_set = {int(r) for r, h in zip(ranks, hit) if h}
    centres = orc.cluster_centers(counts, g["n_cells"], g["interval"], None, wl_set)
    assert centres == g["centres"]
    adj = {}
    for x, y, _ in edges.tolist():
        adj.setdefault(x, []).append(y); adj.setdefault(y, []).append(x)
    clustering = orc.cluster(adj, centres)
    assert {k: tuple(v) for k, v in clustering.items()} == {k: (c, l) for k, c, l in g["clustering"]}
    assign = {orc.unrank(k): orc.unrank(v[0]) for k, v in clustering.items() if v[0] != -1 and k in counts}
    assert assign == g["assignments"]
    # --high_sens (barcode_graph.py:370-385) with the centre order the reference iterated
    order = synth.rank_many(g["hs_centre_order"])
    un = np.asarray([r for r in counts if orc.unrank(r) not in assign], dtype=np.uint32)
    am, dist = orc.nearest(un, order, 2)
    hs = dict(assign)
    for r, j in zip(un.tolist(), am.tolist()):
        if j >= 0:
            hs[orc.unrank(r)] = g["hs_centre_order"][j]
    assert hs == g["hs_assignments"]


def test_c1(gold_c1):
    wl, cells, obs, valid, cfg = synth.make_dataset("C1")
    ranks, counts = orc.dedup_count(obs, valid)
    assert ranks.size == gold_c1["n_distinct"]
    pairs = np.stack([ranks, counts], 1).astype(np.uint64)
    assert hashlib.sha256(pairs.tobytes()).hexdigest() == gold_c1["counts_sha256"]
    a, b, d, _ = orc.Index(ranks).edges(cfg["threshold"])
    assert np.array_equal(np.stack([a, b, d], 1).astype(np.int64), np.asarray(gold_c1["edges"], dtype=np.int64))
