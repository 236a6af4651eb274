#!/usr/bin/env python3
"""The five BASELINE.json configs through the array pipeline (badger_b200.pipeline.assign_packed) on one GPU: one JSON
line per config with reads, distinct barcodes, edges, centres, assigned reads, seconds per stage and reads/s.
Synthetic inputs per SURVEY.md 8(d) (badger_b200.synth).  Development / reporting aid, not part of bench.py."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import badger_b200  # noqa: E402
from badger_b200 import pipeline, synth  # noqa: E402


def main():
    names = os.environ.get("CONFIGS", "C1,C2,C3,C4,C5").split(",")
    badger_b200.init([0])
    for name in names:
        t0 = time.perf_counter()
        wl, cells, obs, valid, cfg = synth.make_dataset(name)
        wls = np.sort(wl)
        gen_s = time.perf_counter() - t0
        if name == names[0]:
            pipeline.assign_packed(obs[:20000], valid[:20000], threshold=cfg["threshold"], n_cells=100, whitelist_sorted=wls)   # warm
        T = {}
        t0 = time.perf_counter()
        out, info = pipeline.assign_packed(obs, valid, threshold=cfg["threshold"], n_cells=cfg["n_cells"], whitelist_sorted=wls, timings=T)
        dt = time.perf_counter() - t0
        n = info["distinct"]
        print(json.dumps({"config": name, "threshold": cfg["threshold"], **info, "seconds": round(dt, 4), "reads_per_s": obs.size / dt,
                          "pairs_decided_per_s_edges_stage": n * (n - 1) / 2 / max(T.get("edges", 1e-9), 1e-9),
                          "stages_s": {k: round(v, 4) for k, v in T.items()}, "synthesis_s": round(gen_s, 1)}), flush=True)


if __name__ == "__main__":
    main()
