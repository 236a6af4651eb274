import os, sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import badger_b200
from badger_b200 import ops, synth
badger_b200.init([0])
wl, cells, obs, valid, cfg = synth.make_dataset("C2")
s = synth.sorted_unique(obs[valid])
sp = torch.from_numpy(s.view(np.int32)).pin_memory().numpy().view(np.uint32)
for _ in range(3): ops.edges_build_part(sp, 1, 0, 1)
os.environ["BDG_TRACE"] = "1"
for _ in range(3):
    t0 = time.perf_counter(); a, b, d = ops.edges_build_part(sp, 1, 0, 1); print("python total %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
