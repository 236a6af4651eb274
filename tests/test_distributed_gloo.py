"""world_size-2 run of the one-process-per-GPU sharding on CPU (gloo): the host-side partition / merge logic
of badger_b200.parallel, with the per-part device operator replaced by the oracle-backed stand-in."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import cpu_ops
    from badger_b200 import synth
    from badger_b200.parallel import edges_build_distributed, part_pairs
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = synth.rng_for(123)
    cells = rng.integers(0, 1 << 32, 150, dtype=np.uint64).astype(np.uint32)
    obs, _ = synth.simulate_reads(cells, 12000, 0.06, rng)
    s = np.unique(obs)
    a, b, d = edges_build_distributed(s, 2, build_part=cpu_ops.edges_build_part)
    la, lb, ld = edges_build_distributed(s, 2, build_part=cpu_ops.edges_build_part, gather=False)
    fa, fb, fd = cpu_ops.edges_build(s, 2)
    o = np.lexsort((b, a))
    ok = np.array_equal(a[o], fa) and np.array_equal(b[o], fb) and np.array_equal(d[o], fd)
    pairs = sum(part_pairs(s.size, p, world) for p in range(world))
    ok = ok and pairs == s.size * (s.size - 1) // 2 and 0 < la.size < fa.size
    q.put((rank, bool(ok), int(la.size), int(fa.size)))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == res[0][3]          # the parts tile the edge set exactly
