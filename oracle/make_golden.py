#!/usr/bin/env python3
"""Generate tests/golden/*.json(+tsv) by running the UNMODIFIED reference (/root/reference).

Run in the authoring container only:  python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so these fixtures --
outputs of the reference itself on seeded inputs -- are what pins the oracle and the CUDA path.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from badger_b200 import synth                 # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ALPH = "ACGT"


def mutate(s: str, rng, n_edits: int) -> str:
    s = list(s)
    for _ in range(n_edits):
        kind = rng.integers(0, 3)
        pos = int(rng.integers(0, len(s)))
        if kind == 0:
            s[pos] = ALPH[(ALPH.index(s[pos]) + int(rng.integers(1, 4))) % 4]
        elif kind == 1:
            s.insert(pos, ALPH[int(rng.integers(0, 4))])
        else:
            del s[pos]
    while len(s) < 16:
        s.append(ALPH[int(rng.integers(0, 4))])
    return "".join(s[:16])


def rand_bc(rng) -> str:
    return "".join(ALPH[int(x)] for x in rng.integers(0, 4, 16))


def low_complexity(rng) -> str:
    kind = int(rng.integers(0, 4))
    if kind == 0:
        return ALPH[int(rng.integers(0, 4))] * 16
    if kind == 1:
        u = "".join(ALPH[int(x)] for x in rng.integers(0, 4, 2))
        return (u * 8)[:16]
    if kind == 2:
        u = "".join(ALPH[int(x)] for x in rng.integers(0, 4, 3))
        return (u * 6)[:16]
    u = "".join(ALPH[int(x)] for x in rng.integers(0, 4, int(rng.integers(4, 8))))
    return (u * 4)[:16]


def gold_pairs():
    rng = synth.rng_for(7)
    base = "GATTACAGATTCCATG"
    pairs = [(base, x) for x in ("ATTACAGATTCCATGC", "GATTAGAGATTGCATG", "GATTACAGTTTCCATG", "GATTACGATTCCATGA",
                                 "GATTACCAGATTCCAT", "GTTTACAGATTCCTTG", "GATTACAGCCTCCATG")]
    pairs += [("A" * 16, "A" * 15 + "C"), ("AC" * 8, "CA" * 8), ("ACGT" * 4, "ACGT" * 3 + "ACGA")]
    for _ in range(400):
        a = rand_bc(rng) if rng.random() < 0.7 else low_complexity(rng)
        b = mutate(a, rng, int(rng.integers(1, 5)))
        if a != b:
            pairs.append((a, b))
    for _ in range(60):
        a, b = low_complexity(rng), low_complexity(rng)
        if a != b:
            pairs.append((a, b))
    for _ in range(60):
        pairs.append((rand_bc(rng), rand_bc(rng)))
    m = rh.load()
    out = []
    for a, b in pairs:
        ed, D, S = rh.ref_pair(a, b)
        out.append(dict(a=a, b=b, ra=m["common"].rank(a, 16), rb=m["common"].rank(b, 16), ed=ed, D=D, S=S))
    return dict(thresholds={str(t): rh.ref_threshold(t) for t in range(0, 5)}, pairs=out)


def clustered_reads(rng, n_cells, reads, perr, lowc=0):
    cells = [rand_bc(rng) for _ in range(n_cells)] + [low_complexity(rng) for _ in range(lowc)]
    out = []
    for _ in range(reads):
        c = cells[int(rng.integers(0, len(cells)))]
        k = int(rng.binomial(16, perr))
        out.append(mutate(c, rng, k) if k else c)
    return out


def gold_graphs():
    m = rh.load()
    rk = m["common"].rank
    graphs = []
    toy = ["ACGTACGTACGTACGT", "ACGTACGTACGTACGA", "ACGTACGTACGTACGT", "TCGTACGTACGTACGT", "ACGTACGAACGTACGT"]
    specs = [("toy_t1", toy, 1)]
    rng = synth.rng_for(11)
    reads_a = clustered_reads(rng, 60, 2500, 0.06)
    reads_b = clustered_reads(rng, 25, 1500, 0.08, lowc=12)
    # 17-mers (stripped), wrong-length strings (skipped): barcode_graph.py:195-197
    reads_c = clustered_reads(rng, 30, 800, 0.05)
    reads_c = [r + "A" if i % 7 == 0 else (r[:12] if i % 11 == 0 else r) for i, r in enumerate(reads_c)]
    for t in (0, 1, 2, 3):
        specs.append(("clustered_t%d" % t, reads_a, t))
    for t in (1, 2):
        specs.append(("lowcomplexity_t%d" % t, reads_b, t))
    specs.append(("ragged_t1", reads_c, 1))
    specs.append(("empty_t1", [], 1))
    specs.append(("single_t2", ["ACGTACGTACGTACGT"] * 3, 2))
    for name, reads, t in specs:
        g, counts, edges = rh.ref_graph(reads, t)
        if len(reads) and len(counts) > 200 and t in (1, 2):    # the process-pool variant must agree
            _, counts4, edges4 = rh.ref_graph(reads, t, threads=2)
            assert counts4 == counts and edges4 == edges, name
        sample_close = []
        ranks = [c[0] for c in counts]
        for q in ranks[:: max(1, len(ranks) // 8)][:8]:
            sample_close.append(dict(query=q, close=rh.ref_get_close(ranks, q, t)))
        graphs.append(dict(name=name, t=t, reads=reads, counts=counts, edges=edges, get_close=sample_close))
        print(name, "N=%d edges=%d" % (len(counts), len(edges)))
    return graphs


def gold_kmer():
    rng = synth.rng_for(13)
    cases = []
    R1 = "CTACACGACGCTCTTCCGATCT"
    for ci in range(160):
        style = ci % 4
        if style == 0:      # whitelist-like: 16-mers vs a mutated member
            known = [rand_bc(rng) for _ in range(int(rng.integers(1, 40)))]
            q = mutate(known[int(rng.integers(0, len(known)))], rng, int(rng.integers(0, 4)))
        elif style == 1:    # low complexity / duplicates in the known list
            known = [low_complexity(rng) for _ in range(int(rng.integers(2, 12)))]
            known += known[:2]
            q = mutate(known[0], rng, int(rng.integers(0, 3)))
        elif style == 2:    # the live caller: adapter vs a read window (barcode_callers.py:162,188-199)
            known = [R1]
            pre = "".join(ALPH[int(x)] for x in rng.integers(0, 4, int(rng.integers(5, 40))))
            q = pre + mutate(R1 + "ACGTAC", rng, int(rng.integers(0, 4))) + rand_bc(rng)
        else:               # short / degenerate
            known = [rand_bc(rng)[: int(rng.integers(3, 17))] for _ in range(5)]
            q = rand_bc(rng)[: int(rng.integers(3, 17))]
        kw = dict(max_hits=int(rng.integers(0, 4)), min_kmers=int(rng.integers(1, 4)),
                  hits_delta=int(rng.integers(0, 3)), ignore_equal=bool(rng.integers(0, 2)))
        res = rh.ref_get_occurrences(known, q, 6, **kw)
        res_arr = rh.ref_get_occurrences(known, q, 6, array=True, **kw)
        assert res == res_arr
        cases.append(dict(known=known, query=q, k=6, kw=kw, result=res))
    return cases


def make_tricky(path, rng):
    """Rewrite an extraction TSV with everything pandas.read_csv treats specially that such files can plausibly hold:
    blank lines, lines of spaces, repeated header lines, short rows, empty / NA-string / '*' barcodes, barcodes of other
    lengths, read ids with spaces and '#', no newline at the end of the file (badger.py:91-111 is the code under test)."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    header, rows = lines[0], [ln for ln in lines[1:] if ln]
    out = [header]
    for i, ln in enumerate(rows):
        f = ln.split("\t")
        r = rng.random()
        if r < 0.03: f[1] = ""
        elif r < 0.06: f[1] = str(rng.choice(["NA", "nan", "NULL", "N/A", "None", "<NA>", "n/a", "#N/A"]))
        elif r < 0.08: f[1] = f[1][:int(rng.integers(1, 16))]
        elif r < 0.10 and f[1] != "*": f[1] = f[1] + "ACGTA"[:int(rng.integers(2, 5))]
        elif r < 0.11: f[1] = "barcode"
        r = rng.random()
        if r < 0.02: f[0] = str(rng.choice(["#read_id", " lead_%d" % i, "trail_%d " % i, "a b_%d" % i, "x#y_%d" % i, "0a1f-%d" % i]))
        if rng.random() < 0.03: f = f[:int(rng.integers(1, 8))]
        out.append("\t".join(f))
        if rng.random() < 0.02: out.append(str(rng.choice(["", "   ", header])))
    with open(path, "w") as fh:
        fh.write("\n".join(out))                        # no trailing newline


def gold_pipeline(name, reads, n_cells, W, perr, t, seed, interval=25, extra17=0.1, tricky=False):
    """Full badger.py run on files (badger.py:62-175) + graph internals, with and without --high_sens."""
    rng = synth.rng_for(seed)
    wl = synth.make_whitelist(W, rng)
    cells = synth.pick_cells(wl, n_cells, rng)
    obs, valid = synth.simulate_reads(cells, reads, perr, rng)
    d = os.path.join(GOLD, name)
    os.makedirs(d, exist_ok=True)
    tsv, wlf = os.path.join(d, "reads.tsv"), os.path.join(d, "whitelist.txt")
    synth.write_extraction_tsv(tsv, obs, valid, rng, extra17_frac=extra17)
    if tricky:
        make_tricky(tsv, rng)
    synth.write_whitelist(wlf, wl)
    meta = dict(name=name, t=t, n_cells=n_cells, interval=interval)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "OUT")
        stdout = rh.ref_main(["-r", tsv, "-l", wlf, "-d", "tenX_v3", "-t", str(t), "--n_cells", str(n_cells),
                              "-i", str(interval), "-o", out])
        with open(out + "_output_file.tsv", "rb") as fh:
            data = fh.read()
        with open(os.path.join(d, "expected_output_file.tsv"), "wb") as fh:
            fh.write(data)
        meta["stdout_tail"] = [ln for ln in stdout.splitlines() if ln.strip().lstrip("-").isdigit()]
    # internals: re-run the graph steps in-process the way badger.main does
    import pandas as pd
    df = pd.read_csv(tsv, sep="\t")
    barcodes = df["barcode"].dropna()
    barcodes = barcodes[(barcodes != "*") & (barcodes != "barcode")].tolist()
    g, counts, edges = rh.ref_graph(barcodes, t)
    with open(wlf) as fh:
        barcode_list = set(fh.read().split("\n"))
    centres, clustering = rh.ref_cluster(g, None, barcode_list, n_cells, interval)
    assignments = g.assign_by_cluster(16)
    base_assign = dict(assignments)
    order = list(set(assignments.values()))            # the iteration order postprocessing() will see
    post = g.postprocessing(assignments, 16)
    meta.update(counts=counts, edges=edges, centres=centres,
                clustering=[[k, v[0], v[1]] for k, v in clustering.items()],
                assignments=base_assign, hs_centre_order=order,
                hs_assignments={k: v for k, v in post.items() if v not in ("", "*")})
    with open(os.path.join(d, "golden.json"), "w") as fh:
        json.dump(meta, fh)
    print(name, "N=%d edges=%d centres=%d assigned=%d hs=%d" % (
        len(counts), len(edges), len(centres), len(base_assign), len(meta["hs_assignments"])))


def gold_c1():
    """BASELINE.json config 1 in full; inputs are regenerated from the seed by the tests."""
    wl, cells, obs, valid, cfg = synth.make_dataset("C1")
    strs = [s.decode() for s in synth.unrank_many(obs[valid]).tolist()]
    g, counts, edges = rh.ref_graph(strs, cfg["threshold"])
    wl_set = set(s.decode() for s in synth.unrank_many(wl).tolist()) | {""}
    centres, clustering = rh.ref_cluster(g, None, wl_set, cfg["n_cells"], 25)
    assign = g.assign_by_cluster(16)
    h = hashlib.sha256()
    for k in sorted(assign):
        h.update(("%s\t%s\n" % (k, assign[k])).encode())
    return dict(config="C1", n_reads=int(valid.sum()), n_distinct=len(counts),
                counts_sha256=hashlib.sha256(np.asarray(counts, dtype=np.uint64).tobytes()).hexdigest(),
                edges=edges, n_centres=len(centres),
                centres_sha256=hashlib.sha256(np.asarray(centres, dtype=np.uint64).tobytes()).hexdigest(),
                n_assigned=len(assign), assignments_sha256=h.hexdigest())


def main():
    os.makedirs(GOLD, exist_ok=True)
    if sys.argv[1:] == ["tricky"]:                      # only the fixture added later (the others stay as committed)
        gold_pipeline("pipeline_tricky", reads=3000, n_cells=150, W=3000, perr=0.05, t=1, seed=23, tricky=True)
        return
    with open(os.path.join(GOLD, "pairs.json"), "w") as fh:
        json.dump(gold_pairs(), fh)
    with open(os.path.join(GOLD, "graphs.json"), "w") as fh:
        json.dump(gold_graphs(), fh)
    with open(os.path.join(GOLD, "kmer_occurrences.json"), "w") as fh:
        json.dump(gold_kmer(), fh)
    gold_pipeline("pipeline_t1", reads=3000, n_cells=200, W=4000, perr=0.05, t=1, seed=21)
    gold_pipeline("pipeline_t2", reads=2500, n_cells=120, W=3000, perr=0.07, t=2, seed=22)
    gold_pipeline("pipeline_tricky", reads=3000, n_cells=150, W=3000, perr=0.05, t=1, seed=23, tricky=True)
    with open(os.path.join(GOLD, "c1.json"), "w") as fh:
        json.dump(gold_c1(), fh)
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
