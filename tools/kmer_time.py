"""a-5 timing on the GPU box: the resident index queried with Q observed barcodes, posting lists against the scan.
Prints the kernel's device time (bdg_kmer_index_info) beside the wall clock of the host-buffer call."""
import os
import time

import numpy as np
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from badger_b200 import ops, synth

wl, cells, obs, valid, cfg = synth.make_dataset("C2")
s = np.unique(obs[valid])
print("W", wl.size)
for Q in (16, 256, 4096, 65536):
    qk = s[:Q]
    for form, mw in (("postings", "0"), ("scan", str(1 << 40))):
        if form == "scan" and Q > 4096:
            continue
        os.environ["BDG_KMER_POST_MIN_W"] = mw
        t0 = time.perf_counter(); ix = ops.KmerIndex(wl); b = time.perf_counter() - t0
        h = ix.query(qk, min_kmers=4)[0]
        t0 = time.perf_counter(); ix.query(qk, min_kmers=4); dt = time.perf_counter() - t0
        print("Q=%6d %-8s build %7.1f ms  query wall %8.2f ms  kernel %8.3f ms  hits %d" % (Q, form, b * 1e3, dt * 1e3, ix.info()["kernel_ms"], h.size))
        ix.free()
