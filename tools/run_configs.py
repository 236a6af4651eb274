#!/usr/bin/env python3
"""The five BASELINE.json configs through the array pipeline (badger_b200.pipeline.assign_packed) on one or several GPUs of
one box (DEVICE_SETS): one JSON line per config and device set with reads, distinct barcodes, edges, centres, assigned reads, seconds per stage and reads/s.
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
    # DEVICE_SETS="1,8": every config once on the first GPU and once on the first eight (same synthetic input, results compared)
    sets = [int(x) for x in os.environ.get("DEVICE_SETS", "1").split(",")]
    workers = int(os.environ.get("SYNTH_WORKERS", "1"))      # > 1: threaded synthesis (another stream than the serial one, see synth)
    warmed = set()
    for name in names:
        t0 = time.perf_counter()
        wl, cells, obs, valid, cfg = synth.make_dataset(name, workers=workers)
        wls = np.sort(wl)
        gen_s = time.perf_counter() - t0
        first = None
        for k in sets:
            badger_b200.init(list(range(k)))
            if k not in warmed:
                warmed.add(k)
                pipeline.assign_packed(obs[:200000], valid[:200000], threshold=cfg["threshold"], n_cells=100, whitelist_sorted=wls)   # warm
            first_s = None
            for rep in range(2):                      # the first full-size call grows every workspace (cudaMalloc, page-locked pools); the second is the steady state
                T = {}
                t0 = time.perf_counter()
                out, info = pipeline.assign_packed(obs, valid, threshold=cfg["threshold"], n_cells=cfg["n_cells"], whitelist_sorted=wls, timings=T, form="u32")
                dt = time.perf_counter() - t0
                if rep == 0:
                    first_s = dt
            out = out[0].astype(np.uint64) | (np.uint64(1) << np.uint64(32)) * (out[1] == 0)        # one array for the comparison below
            n = info["distinct"]
            same = None
            if first is None:
                first = (out, info)
            else:
                same = bool(np.array_equal(first[0], out) and first[1] == info)
            print(json.dumps({"config": name, "n_gpus": k, "threshold": cfg["threshold"], **info, "seconds": round(dt, 4), "reads_per_s": obs.size / dt,
                              "pairs_decided_per_s_edges_stage": n * (n - 1) / 2 / max(T.get("edges", 1e-9), 1e-9),
                              "stages_s": {k2: round(v, 4) for k2, v in T.items()}, "first_call_seconds": round(first_s, 4), "synthesis_s": round(gen_s, 1), "synthesis_workers": workers,
                              "identical_to_first_device_set": same}), flush=True)
            del out


if __name__ == "__main__":
    main()
