import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import badger_b200
from badger_b200 import synth
badger_b200.init([0])
L = badger_b200.lib()
wl, cells, obs, valid, cfg = synth.make_dataset("C2", reads=1000000)
s = np.unique(obs[valid]); n = s.size
dev = torch.device("cuda", 0)
d_sorted = torch.from_numpy(s.view(np.int32)).to(dev)
cap = 64 * n
d_a = torch.empty(cap, dtype=torch.int32, device=dev); d_b = torch.empty(cap, dtype=torch.int32, device=dev)
d_d = torch.empty(cap, dtype=torch.uint8, device=dev); d_c = torch.zeros(1, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream()
t = int(os.environ.get("T", "1"))
for rep in range(3):
    badger_b200._lib.check(L.bdg_dev_edges_build(d_sorted.data_ptr(), n, t, 0, 1, d_a.data_ptr(), d_b.data_ptr(), d_d.data_ptr(), cap, d_c.data_ptr(), st.cuda_stream))
    torch.cuda.synchronize()
    print("=== launch", rep, "done", flush=True)
