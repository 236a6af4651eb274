#!/usr/bin/env python3
"""Development sweep: time the edge kernel (device-resident input, CUDA events) over sizes, thresholds and the
BDG_EDGE_* experiment switches.  Not part of the product or of bench.py."""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import badger_b200  # noqa: E402
from badger_b200 import synth  # noqa: E402


def dataset(reads):
    wl, cells, obs, valid, cfg = synth.make_dataset("C2", reads=reads, workers=min(32, len(os.sched_getaffinity(0))))
    return synth.sorted_unique(obs[valid])


def time_edges(s, t, reps=3):
    L = badger_b200.lib()
    dev = torch.device("cuda", 0)
    n = int(s.size)
    d_sorted = torch.from_numpy(s.view(np.int32)).to(dev)
    cap = 64 * n
    d_a = torch.empty(cap, dtype=torch.int32, device=dev); d_b = torch.empty(cap, dtype=torch.int32, device=dev)
    d_d = torch.empty(cap, dtype=torch.uint8, device=dev); d_c = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()

    def go():
        badger_b200._lib.check(L.bdg_dev_edges_build(d_sorted.data_ptr(), n, t, 0, 1, d_a.data_ptr(), d_b.data_ptr(), d_d.data_ptr(),
                                                     cap, d_c.data_ptr(), st.cuda_stream))
    go(); go()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(st); go(); e1.record(st); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    import ctypes as C
    v = (C.c_ulonglong * 5)()
    L.bdg_dev_edges_stats(v, st.cuda_stream)
    raw = (C.c_ulonglong * 8)()
    L.bdg_dev_edges_stats_raw(raw, st.cuda_stream)
    global RAW
    RAW = list(raw)
    bal = (C.c_ulonglong * 6)()
    L.bdg_dev_edges_balance(bal, st.cuda_stream)
    nw = 148 * 3 * 8
    global BALANCE
    BALANCE = " ".join("p%d: mean %.0f us max %.0f us" % (p, bal[2 * p] / nw / 1e3, bal[2 * p + 1] / 1e3) for p in range(3) if bal[2 * p + 1])
    return best, int(d_c.item()), list(v)


def main():
    badger_b200.init([0])
    sizes = [int(x) for x in os.environ.get("SWEEP_READS", "400000,1000000").split(",")]
    ts = [int(x) for x in os.environ.get("SWEEP_T", "1,2").split(",")]
    knobs = {k: os.environ.get("SWEEP_" + k, d).split(",") for k, d in (("ITEMS", "16"), ("MODE", "2,1"))}
    data = {r: dataset(r) for r in sizes}
    for r, t in itertools.product(sizes, ts):
        s = data[r]
        n = s.size
        for items, mode in itertools.product(knobs["ITEMS"], knobs["MODE"]):
            os.environ["BDG_EDGE_ITEMS"] = items
            badger_b200.lib().bdg_set_edge_mode(int(mode))
            ms, edges, (subs, fulls, scored, cand, nS) = time_edges(s, t)
            print("reads=%8d N=%8d t=%d items=%-3s mode=%s  %9.3f ms  %.3e pairs/s  edges=%d  sub-tiles=%d full=%.2f%% scored=%.3e cand=%.3e" % (
                r, n, t, items, {"0": "dense ", "1": "sparse", "2": "join  "}[mode], ms, n * (n - 1) / 2 / (ms * 1e-3), edges, subs, 100.0 * fulls / max(subs, 1), scored, cand), flush=True)
            print("      balance (warp exit times): " + BALANCE, flush=True)
            if mode == "2" and t == 2:
                print("      join: units=%d pairs tested=%.3e candidates=%.3e D<=2=%.3e scored=%.3e warp busy mean %.0f us max %.0f us" % (
                    RAW[0], RAW[2], RAW[3], RAW[6], RAW[7], RAW[4] / (148 * 5 * 8) / 1e3, RAW[5] / 1e3), flush=True)


if __name__ == "__main__":
    main()
