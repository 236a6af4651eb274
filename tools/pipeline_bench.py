#!/usr/bin/env python3
"""Where does the time of the drop-in CLI go?  Writes a config-shaped extraction TSV + whitelist, runs the stages of
badger.py one by one (same calls, same order) and prints seconds per stage and reads/s.  Development aid."""
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import badger_b200  # noqa: E402
from badger_b200 import BarcodeGraph, synth  # noqa: E402


def main():
    cfg_name = os.environ.get("PIPE_CONFIG", "C2")
    reads = int(os.environ.get("PIPE_READS", "0")) or None
    high_sens = os.environ.get("PIPE_HS", "0") == "1"
    wl, cells, obs, valid, cfg = synth.make_dataset(cfg_name, reads=reads)
    tmp = tempfile.mkdtemp(prefix="bdg_pipe_")
    tsv, wlf = os.path.join(tmp, "reads.tsv"), os.path.join(tmp, "wl.txt")
    synth.write_whitelist(wlf, wl)
    synth.write_extraction_tsv(tsv, obs, valid, synth.rng_for(5))
    R = obs.size
    T = {}
    t0 = time.perf_counter()
    with open(wlf) as f:
        barcode_list = set(f.read().split("\n"))
    T["whitelist file -> set"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    df = pd.read_csv(tsv, sep="\t")
    ids = df["#read_id"].tolist()
    observed = df["barcode"].fillna('*').tolist()
    barcodes = df["barcode"].dropna()
    barcodes = barcodes[barcodes != "*"]
    barcodes = barcodes[barcodes != "barcode"].tolist()
    T["read_csv + column lists"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    read_assignment = []
    for i in range(len(ids)):
        if ids[i] != "#read_id":
            o = observed[i]
            if o != "barcode":
                if len(o) == 17:
                    o = o[:-1]
                read_assignment.append((ids[i], o))
    T["read_assignment loop (badger.py:103-110)"] = time.perf_counter() - t0
    badger_b200.init()
    g = BarcodeGraph(cfg["threshold"])
    g.graph_construction(barcodes[:1000], 16, 1)      # warm the library (workspaces, module load)
    g = BarcodeGraph(cfg["threshold"])
    t0 = time.perf_counter()
    g.graph_construction(barcodes, 16, 1)
    T["graph_construction"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.cluster(None, barcode_list, cfg["n_cells"], 16, 25)
    T["cluster"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.output_file(read_assignment, os.path.join(tmp, "OUT"), None, 16, high_sens)
    T["output_file"] = time.perf_counter() - t0
    tot = sum(T.values())
    print("config %s: %d reads, %d distinct, %d edges, threshold %d" % (cfg_name, R, len(g.counts), g.edge_arrays()[0].size, cfg["threshold"]))
    for k, v in T.items():
        print("  %-45s %8.3f s" % (k, v))
    for k, v in sorted(g.timings.items()):
        print("      . %-39s %8.3f s" % (k, v))
    print("  %-45s %8.3f s  -> %.0f reads/s" % ("total", tot, R / tot))
    native_route(tsv, wlf, cfg, tmp, R, high_sens, os.path.join(tmp, "OUT_output_file.tsv"))


def native_route(tsv, wlf, cfg, tmp, R, high_sens, other_output):
    """The route badger.py takes by default (run_native): native reader, GPU pack, array pipeline, native writer."""
    from badger_b200 import ops, pipeline, tsvio
    for rep in range(2):                                   # first pass warms the page cache / workspaces
        T, P = {}, {}
        t0 = time.perf_counter()
        t = tsvio.ExtractionTsv(tsv)
        T["extraction TSV -> rows (native reader)"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        has = t.has_barcode
        ranks = np.zeros(t.rows, np.uint32)
        packed, ok = ops.pack16(t.seqs16[has])
        assert ok.all()
        ranks[has] = packed
        T["pack16 of the reads (GPU)"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        w, wok = ops.pack16(tsvio.whitelist_records(wlf))
        whitelist = ops.sorted_unique(w[wok])          # (badger.py itself packs and sorts the records on the device: ops.pack16_sorted)
        T["whitelist file -> sorted uint32 (native + GPU)"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        centre, info = pipeline.assign_packed(ranks, has, threshold=cfg["threshold"], n_cells=cfg["n_cells"], interval=25,
                                              whitelist_sorted=whitelist, high_sens=high_sens, centre_order="set", timings=P)
        T["assign_packed"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        out = os.path.join(tmp, "NATIVE_output_file.tsv")
        t.write(out, centre)
        T["output TSV (native writer)"] = time.perf_counter() - t0
        t.close()
    tot = sum(T.values())
    print("native route (badger.py default): %d rows, %d distinct, %d edges, disconnected %d" % (R, info["distinct"], info["edges"], info["disconnected"]))
    for k, v in T.items():
        print("  %-45s %8.3f s" % (k, v))
    for k, v in P.items():
        print("      . %-39s %8.3f s" % (k, v))
    print("  %-45s %8.3f s  -> %.0f reads/s" % ("total", tot, R / tot))
    with open(out, "rb") as a, open(other_output, "rb") as b:
        same = a.read() == b.read()
    print("  output files of the two routes identical: %s" % same)
    assert same


if __name__ == "__main__":
    main()
