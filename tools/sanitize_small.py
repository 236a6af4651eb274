#!/usr/bin/env python3
"""Small run of every kernel for compute-sanitizer (memcheck / racecheck): edges in both modes and thresholds,
dedup, membership, both scorers, k-mer scoring, packing.  Sizes are tiny: the tools slow kernels down 10-100x."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import badger_b200  # noqa: E402
from badger_b200 import ops, synth  # noqa: E402

badger_b200.init([0])
L = badger_b200.lib()
rng = synth.rng_for(99)
cells = rng.integers(0, 1 << 32, 60, dtype=np.uint64).astype(np.uint32)
obs, _ = synth.simulate_reads(cells, 6000, 0.06, rng)
s = np.unique(np.concatenate([obs, np.arange(1000, 1400, dtype=np.uint32)]))
for mode in (1, 0):
    L.bdg_set_edge_mode(mode)
    for t in (1, 2, 3):
        sub = s if t < 3 else s[:700]
        a, b, d = ops.edges_build(sub, t)
        print("mode", mode, "t", t, "edges", a.size, flush=True)
L.bdg_set_edge_mode(-1)
d_, c_, m_ = ops.dedup_first_seen(obs, want_map=True)
strs = synth.unrank_many(obs[:2000])
r, ok = ops.pack16(b"".join(strs.tolist()))
hit = ops.member_sorted(np.sort(cells), s)
am, dist = ops.nearest_bounded(s[:3000], cells, 2)
os.environ["BDG_NEAREST_DENSE"] = "1"
am2, dist2 = ops.nearest_bounded(s[:3000], cells, 2)
big_t = rng.integers(0, 1 << 32, 6000, dtype=np.uint64).astype(np.uint32)
os.environ.pop("BDG_NEAREST_DENSE")
am3, _ = ops.nearest_bounded(np.concatenate([s, big_t[:500]]), big_t, 2)      # 3.7e7 pairs: the tiled scorer
hq, hw, cnt, mult = ops.kmer_score(s[:8], cells, min_kmers=2)
print("ok", d_.size, int(ok.all()), int(hit.sum()), int((am >= 0).sum()), int(np.array_equal(am, am2)), int((am3 >= 0).sum()), hq.size)
