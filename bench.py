#!/usr/bin/env python3
"""bench.py -- barcode pairs scored per second on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input: edge construction
(index.py:77-93 + barcode_graph.py:224-249 of the reference) over every unordered pair of the distinct
barcodes of BASELINE.json's config 4 (20 M simulated ONT reads, 10 k cells, ~4.6 M distinct noisy barcodes,
threshold 2) - the configuration the metric "at 1/2/4/8 B200" is quoted on.

  value   whole-job pairs/s with the sorted distinct-barcode array already resident in HBM
          (bdg_dev_edges_build on torch's current stream, CUDA events around every step, max over ranks)
  e2e     the same metric through the public host-buffer call (ops.edges_build_part -> bdg_edges_build_part):
          pinned host input -> H2D -> kernels -> D2H of the edge list, wall clock, max over ranks
  roofline    the step's dominant kernel against the MEASURED integer issue rate of this GPU (probed in the same run)
  cpu_baseline  the oracle's restatement of the reference algorithm on the box's host cores (rank 0, N=1)

N > 1 (torchrun): STRONG scaling - every rank holds the same array (replicated) and takes every N-th
batch of work units of the join (rows of the sorted orders), no data-path collective; the per-rank edge lists are
disjoint and their union is the edge set (checked: the counts add up to the one-GPU count).

Secondary keys at N = 1: `stages` (the other rows of the path), `pipeline` (reads/s of the whole correction step at
C4 and C2), `cli` (the drop-in command file to file at C2), `c2` (the round-1 headline: C2 at t = 1).

`--impl reference` times the reference's own algorithm (oracle port, all host threads) on a bounded sample
of the same workload; /root/reference (pure Python) cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from badger_b200 import synth  # noqa: E402

METRIC = "barcode_pairs_scored_per_s"
UNIT = "pairs/s"
# Algorithmic integer instructions per unit of work of the edge kernels (DESIGN.md "edge construction"; counted
# from the SASS of the code blocks, profiles/*_blocks.txt; the numbers of units come from the kernels' own counters,
# bdg_dev_edges_stats / bdg_dev_edges_stats_raw):
#   join:   per work unit (cursor, slab lookup, staging), per pair tested (quick test + key equality), per candidate
#           (queue + exact distance), per candidate with D <= 2 (hand-over flags + table), per pair scored (6-mer score)
#   sparse: one key-interval test (pass_possible) per column sub-tile, 32 more per surviving sub-tile; one quick
#           test per pair of a surviving 32x32 block; one exact stage (D, then S) per candidate
#   dense:  one light-loop step per pair (1.25 instr at t=1, 6 at t=2), exact stage per candidate
A_TILE = 40
A_QUICK = {1: 14.5, 2: 16.5}       # 58 / 66 SASS instructions per 4 pairs
A_LIGHT = {1: 1.25, 2: 6.0}
A_EXACT = 110                      # loads, un-rotation, pass predicates, dist_small
A_SCORE = 220                      # qgram_score, only for pairs with D <= t
A_JOIN = {"unit": 160, "pair": 22.0, "cand": 85, "d2": 50, "score": 150}
# Share of those instructions that issue on the ALU pipe (LOP3, SHF, ISETP, SEL, VOTE ...; the rest are IMAD on the FMA pipe):
# sm__inst_executed_pipe_alu / (pipe_alu + pipe_fma) of the join kernel in the committed ncu capture (profiles/).  The kernel is
# bound by that pipe (ncu: math-pipe throttle and not-selected are its top stalls), so its roofline is the ALU pipe's issue rate.
ALU_SHARE = {"join": 0.845, "sparse": 0.81, "dense": 0.6}
A_PAIR_SURVEY = {1: 15, 2: 25}   # SURVEY.md §8(d) nominal per-pair figure of a plain all-pairs kernel, reported alongside
# DRAM bytes per step of the dominant kernel from the committed ncu capture of the same command (profiles/r2_join_c4_ncu_full.txt:
# dram__bytes_read.sum + dram__bytes_write.sum of a join launch, mean of launches 1-3 = 38.21 MB + 0.14 MB, times the 25 launches
# of a step).  Not measured in the timed run - ncu cannot run inside it - hence the label.
NCU_TRAFFIC = {"bytes": int(25 * (38.21e6 + 0.14e6)), "workload": ("C4", None, 2, "join", 1),
               "label": "25 join launches x (dram read + write of one launch), ncu --set full capture profiles/r2_join_c4_ncu_full.txt of this build"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", default="C4")
    ap.add_argument("--reads", type=int, default=None, help="override the read count (testing)")
    ap.add_argument("--threshold", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the stages / pipeline / cli / c2 keys")
    ap.add_argument("--mode", choices=["auto", "join", "sparse", "dense"], default="auto", help="edge search strategy (bdg_set_edge_mode)")
    return ap.parse_args()


def dataset(config, reads=None, threshold=None):
    """(cfg, whitelist, observed, valid, sorted distinct) of a BASELINE config, synthesised once per box: the first caller
    writes the arrays to a cache file, the other ranks (and the other arm of the bench) load it."""
    cfg = dict(synth.CONFIGS[config])
    if reads is not None:
        cfg["reads"] = reads
    if threshold is not None:
        cfg["threshold"] = threshold
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(base, "bdg_bench_%s_%d_%g_%d_%d_%d.npz" % (config, cfg["reads"], cfg["perr"], cfg["n_cells"], cfg["whitelist"], cfg["seed"]))
    rank = int(os.environ.get("RANK", "0"))
    if not os.path.exists(path):
        if rank == 0:
            workers = max(1, min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
            wl, cells, obs, valid, _ = synth.make_dataset(cfg, workers=workers)
            s = synth.sorted_unique(obs[valid])
            tmp = path + ".%d.tmp.npz" % os.getpid()
            np.savez(tmp, wl=wl, obs=obs, valid=valid, s=s)
            os.replace(tmp, path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 900:
                    raise RuntimeError("rank 0 did not write %s" % path)
                time.sleep(0.2)
    z = np.load(path)
    return cfg, z["wl"], z["obs"], z["valid"], z["s"]


def workload_name(config, cfg, n):
    return "%s: %d simulated ONT reads, %d cells, %d-entry whitelist, %.0f%% error, %d distinct barcodes, threshold %d" % (
        config, cfg["reads"], cfg["n_cells"], cfg["whitelist"], 100 * cfg["perr"], n, cfg["threshold"])


def config_dict(config, cfg, n):
    """The `config` key: identical in both arms."""
    return {"workload": workload_name(config, cfg, n), "reads": int(cfg["reads"]), "distinct": int(n), "threshold": int(cfg["threshold"]),
            "pairs_per_step": int(n) * (int(n) - 1) // 2}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(s, t, seconds, seed=7):
    """Oracle port of the reference algorithm (6-mer index walk + 3-way verify) on a bounded row sample."""
    from oracle import oracle as orc
    # all the host cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which is not a property of the box)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or orc.num_threads())
    ix = cpu_sample.index.get(id(s))
    if ix is None:
        ix = cpu_sample.index[id(s)] = orc.Index(s)
    rng = np.random.default_rng(seed)
    n = s.size
    k = min(n, 256)
    rows = np.sort(rng.choice(n, k, replace=False)).astype(np.uint32)
    t0 = time.perf_counter()
    ix.edges(t, rows=rows, threads=threads)
    dt = time.perf_counter() - t0
    k2 = int(min(n, max(k, k * seconds / max(dt, 1e-6))))
    rows = np.sort(rng.choice(n, k2, replace=False)).astype(np.uint32)
    t0 = time.perf_counter()
    _, _, _, verified = ix.edges(t, rows=rows, threads=threads)
    dt = time.perf_counter() - t0
    pairs = int(((n - 1) - rows.astype(np.int64)).sum())
    return dict(value=pairs / dt, unit=UNIT, cores=threads, kind="port",
                sample="%d of %d query rows (uniform), each against the full 6-mer index of %d barcodes; %.1f s; "
                       "oracle/badger_oracle.c restating index.py:77-93 + barcode_graph.py:224-249" % (k2, n, n, dt)), dt


cpu_sample.index = {}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    cfg, _, _, _, s = dataset(args.config, args.reads, args.threshold)
    t = cfg["threshold"]
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        cb, dt = cpu_sample(s, t, per_step, seed=7 + i)
        if i >= args.warmup:
            vals.append(cb["value"]); secs.append(dt)
    v = float(np.mean(vals)) if vals else 0.0
    cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * float(np.mean(secs)) if secs else None, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_dict(args.config, cfg, s.size),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [ln.split(", ") for ln in open(self.f.name).read().strip().splitlines() if ln.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


MODE_ID = {"auto": -1, "dense": 0, "sparse": 1, "join": 2}


def int_probe(torch, L, dev, stream):
    """Integer-pipe issue rates of this GPU (the roofline denominator; not in MEASURED_PEAKS.json)."""
    import ctypes as C
    import badger_b200
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    probe = {}
    for kind, label in ((0, "lop3"), (1, "imad"), (2, "lop3_imad_mix"), (3, "popc")):
        best = 0.0
        iters = 4000 if kind != 3 else 1000
        for rep in range(4):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            ops_pt = C.c_ulonglong(0)
            e0.record(stream)
            badger_b200._lib.check(L.bdg_dev_pipe_probe(kind, sms * 8, iters, sink.data_ptr(), C.byref(ops_pt), stream.cuda_stream))
            e1.record(stream)
            torch.cuda.synchronize()
            rate = ops_pt.value * 256 * sms * 8 / (e0.elapsed_time(e1) * 1e-3)
            if rep:
                best = max(best, rate)
        probe[label] = best / 1e12       # T thread-instructions / s
    return probe


class EdgeStep:
    """Device-resident edge construction of one part (bdg_dev_edges_build on torch's stream) with CUDA-event timing."""

    def __init__(self, torch, L, dev, stream, s, t, part, nparts):
        self.torch, self.L, self.dev, self.stream = torch, L, dev, stream
        self.n, self.t, self.part, self.nparts = int(s.size), t, part, nparts
        self.d_sorted = torch.from_numpy(s.view(np.int32)).to(dev)
        self.cap = 0
        self.d_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
        self.alloc(max(1 << 16, (16 if t <= 1 else 32) * self.n // nparts + 1024))

    def alloc(self, cap):
        torch = self.torch
        self.cap = cap
        self.d_a = torch.empty(cap, dtype=torch.int32, device=self.dev)
        self.d_b = torch.empty(cap, dtype=torch.int32, device=self.dev)
        self.d_d = torch.empty(cap, dtype=torch.uint8, device=self.dev)

    def launch(self):
        import badger_b200
        badger_b200._lib.check(self.L.bdg_dev_edges_build(self.d_sorted.data_ptr(), self.n, self.t, self.part, self.nparts, self.d_a.data_ptr(),
                                                          self.d_b.data_ptr(), self.d_d.data_ptr(), self.cap, self.d_count.data_ptr(),
                                                          self.stream.cuda_stream))

    def settle(self, warm):
        """First run sizes the edge buffers; a poisoned or oversized count is an error, never a timed run."""
        self.launch()
        self.torch.cuda.synchronize()
        count = int(self.d_count.item())
        assert count >= 0, "a work list of the edge kernels overflowed (bit 63 of the count): call bdg_edges_build_part once to grow it"
        if count > self.cap:
            self.alloc(count + 1024)
        for _ in range(warm):
            self.launch()
        self.torch.cuda.synchronize()
        count = int(self.d_count.item())
        assert 0 <= count <= self.cap, "edge buffer too small after sizing: count %d, capacity %d" % (count, self.cap)
        return count

    def timed(self, k):
        ms = 0.0
        for _ in range(k):
            self.flush.fill_(1)                      # evict L2 between timed iterations
            e0, e1 = self.torch.cuda.Event(True), self.torch.cuda.Event(True)
            e0.record(self.stream)
            self.launch()
            e1.record(self.stream)
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms

    def stats(self):
        import ctypes as C
        import badger_b200
        v = (C.c_ulonglong * 5)()
        badger_b200._lib.check(self.L.bdg_dev_edges_stats(v, self.stream.cuda_stream))
        raw = (C.c_ulonglong * 8)()
        badger_b200._lib.check(self.L.bdg_dev_edges_stats_raw(raw, self.stream.cuda_stream))
        d = dict(zip(("sub_tiles", "sub_tiles_scored", "pairs_scored", "candidates", "pairs_S_scored"), (int(x) for x in v)))
        d["candidates_within_t"] = int(raw[6])
        d["warp_busy_ns_sum"], d["warp_busy_ns_max"] = int(raw[4]), int(raw[5])
        return d


def algorithmic(mode, t, st, n_edges_part):
    """Algorithmic integer instructions of one step from the kernels' own unit counters."""
    if mode == "join":
        return (A_JOIN["unit"] * st["sub_tiles"] + A_JOIN["pair"] * st["pairs_scored"] + A_JOIN["cand"] * st["candidates"]
                + A_JOIN["d2"] * st["candidates_within_t"] + A_JOIN["score"] * st["pairs_S_scored"])
    if t not in A_QUICK:
        return None
    per_pair = A_QUICK[t] if mode == "sparse" else A_LIGHT[t]
    tiles = (st["sub_tiles"] + 32 * st["sub_tiles_scored"]) if mode == "sparse" else 0
    score = st["pairs_S_scored"] if mode == "sparse" else n_edges_part     # dense counts S inside the exact stage: at least the edges
    return A_TILE * tiles + per_pair * st["pairs_scored"] + A_EXACT * st["candidates"] + A_SCORE * score


def effective_mode(mode, t, n):
    if mode != "auto":
        return mode if (mode != "join" or t == 2) else "sparse"
    if t not in (1, 2):
        return "dense"
    return "join" if (t == 2 and n >= int(os.environ.get("BDG_JOIN_MIN_N", "150000"))) else "sparse"


def run_b200(args):
    import torch
    import torch.distributed as dist
    import badger_b200
    from badger_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    badger_b200.init([local])
    L = badger_b200.lib()

    cfg, wl, obs, valid, s = dataset(args.config, args.reads, args.threshold)
    t = cfg["threshold"]
    n = int(s.size)
    total_pairs = n * (n - 1) // 2
    stream = torch.cuda.current_stream()
    mode = effective_mode(args.mode, t, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def sum_over_ranks(x):
        if world == 1:
            return int(x)
        tt = torch.tensor([int(x)], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        return int(tt.item())

    probe = int_probe(torch, L, dev, stream)
    sampler = ClockSampler(local)
    sampler.start()                        # covers warm-up, the timed steps and the e2e loop
    badger_b200._lib.check(L.bdg_set_edge_mode(MODE_ID[args.mode]))
    step = EdgeStep(torch, L, dev, stream, s, t, rank, world)
    warm = max(args.warmup, 3)
    n_edges_part = step.settle(warm)

    launches0 = L.bdg_launch_count()
    barrier()
    ms = step.timed(args.steps)
    barrier()
    launches = L.bdg_launch_count() - launches0
    stats = step.stats()
    ms_total = max_over_ranks(ms)
    step_ms = ms / args.steps             # this rank's average step (all launches of the step, CUDA events)
    ms_ranks = [step_ms]
    if world > 1:
        tt = torch.zeros(world, dtype=torch.float64, device=dev)
        tt[rank] = step_ms
        dist.all_reduce(tt)
        ms_ranks = [float(x) for x in tt.tolist()]
    value = total_pairs * args.steps / (ms_total * 1e-3)
    edges_total = sum_over_ranks(n_edges_part)

    # ---- e2e: public host-buffer API, pinned input, edge list back on the host
    s_pinned = torch.from_numpy(s.view(np.int32)).pin_memory()
    s_host = s_pinned.numpy().view(np.uint32)
    r1 = ops.edges_build_part(s_host, t, rank, world)      # warm; two result sets alive at once, as in the timed loop below,
    r2 = ops.edges_build_part(s_host, t, rank, world)      # so that the operator's pinned output pool holds both of them
    assert r1[0].size == n_edges_part, "the host-buffer call and the device call disagree on the edge count"
    del r1, r2
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ea, eb, ed = ops.edges_build_part(s_host, t, rank, world)
    torch.cuda.synchronize()
    e2e_mine = time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_mine)

    def by_rank(x):                                        # every rank's own figure, for the record
        if world == 1:
            return [float(x)]
        tt = torch.zeros(world, dtype=torch.float64, device=dev)
        tt[rank] = float(x)
        dist.all_reduce(tt)
        return [float(v) for v in tt.tolist()]
    e2e_ms_ranks = by_rank(1000 * e2e_mine / args.steps)
    edges_ranks = [int(v) for v in by_rank(n_edges_part)]
    barrier()
    clocks = sampler.stop()
    e2e = {"value": total_pairs * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(4 * n),
           "d2h_bytes_per_step": int(9 * ea.size + 8), "ms_per_step": 1000 * e2e_s / args.steps,
           "api": "badger_b200.ops.edges_build_part -> bdg_edges_build_into (host buffers; every rank uploads the array, the edges cross PCIe while later seed conditions are still joined)"}
    del ea, eb, ed

    # ---- the other search strategy on the same input, for the record (one GPU only: it is several times slower)
    other = None
    if world == 1 and t in (1, 2) and not args.no_secondary:
        other_mode = "sparse" if mode == "join" else ("dense" if n < 600000 else None)
        if other_mode:
            badger_b200._lib.check(L.bdg_set_edge_mode(MODE_ID[other_mode]))
            n_other = step.settle(1)
            assert n_other == n_edges_part, "the two edge modes disagree on the edge count"
            k = 2
            ms_other = step.timed(k) / k
            st_o = step.stats()
            alg_o = algorithmic(other_mode, t, st_o, n_other)
            other = {"mode": other_mode, "ms_per_step": ms_other, "pairs_per_s": total_pairs / (ms_other * 1e-3), "work": st_o,
                     "achieved": alg_o / (ms_other * 1e-3) / 1e12 if alg_o else None,
                     "frac": alg_o / (ms_other * 1e-3) / 1e12 / probe["lop3_imad_mix"] if alg_o else None}
            badger_b200._lib.check(L.bdg_set_edge_mode(MODE_ID[args.mode]))

    out = None
    if rank == 0:
        peak = probe["lop3_imad_mix"]
        alg = algorithmic(mode, t, stats, n_edges_part)
        roof = None
        if alg is not None:
            achieved = alg / (step_ms * 1e-3) / 1e12
            passes = {"join": "join_kernel (one persistent launch per seed condition, 25 per step) behind counting sorts by seed key",
                      "sparse": "sparse_scan_kernel + sparse_tile_kernel<%d,p>, %d passes per step" % (t, 2 if t == 1 else 3),
                      "dense": "edges_kernel<%d>" % t}[mode]
            hbm_bytes = int(9 * n_edges_part + 4 * n * ({"join": 2 * 25 + 25, "sparse": 2 if t == 1 else 3, "dense": 1}[mode]))
            try:
                hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                hbm = 6451.8
            alu_peak = probe["lop3"]
            alu_achieved = achieved * ALU_SHARE[mode]
            roof = {"bound": "int_alu_pipe", "kernel": passes, "achieved": alu_achieved, "peak": alu_peak, "unit": "Tinst/s",
                    "frac": alu_achieved / alu_peak if alu_peak else None,
                    "issue_model": {"achieved": achieved, "peak": peak, "frac": achieved / peak if peak else None,
                                    "note": "all algorithmic instructions against the LOP3+IMAD 1:1 dual-pipe issue rate (the figure round 1 reported)"},
                    "traffic": NCU_TRAFFIC["bytes"] if NCU_TRAFFIC and NCU_TRAFFIC["workload"] == (args.config, args.reads, t, mode, world) else None,
                    "traffic_note": (NCU_TRAFFIC["label"] if NCU_TRAFFIC else "no ncu capture of this build yet") +
                                    "; the algorithmic bytes are hbm.algorithmic_bytes_per_step",
                    "how": "achieved = ALU-pipe share (%.3f, ncu) of the algorithmic integer instructions of rank 0's step (unit counts from the "
                           "kernels' own counters x per-unit costs counted from the SASS, DESIGN.md: %s) / %.3f ms (CUDA events, this run); peak = "
                           "measured issue rate of an independent LOP3 stream on this GPU (bdg_dev_pipe_probe, this run)" % (ALU_SHARE[mode],
                               json.dumps(A_JOIN if mode == "join" else {"tile": A_TILE, "pair": A_QUICK.get(t) if mode == "sparse" else A_LIGHT.get(t),
                                                                         "cand": A_EXACT, "score": A_SCORE}), step_ms),
                    "work": stats,
                    "pairs_decided_per_pair_tested": (total_pairs / world) / max(stats["pairs_scored"], 1),
                    "survey_nominal": {"ops_per_pair": A_PAIR_SURVEY.get(t), "achieved": A_PAIR_SURVEY.get(t, 0) * (total_pairs / world) / (step_ms * 1e-3) / 1e12,
                                       "note": "SURVEY.md 8(d) costs every pair 5(2t+1) instructions; these kernels decide most pairs by key "
                                               "exclusion (sort order), so this figure exceeds the peak by design"},
                    "other_mode": other, "probe_Tinst_per_s": probe,
                    "hbm": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_GBps": hbm_bytes / (step_ms * 1e-3) / 1e9, "peak_GBps": hbm,
                            "frac": hbm_bytes / (step_ms * 1e-3) / 1e9 / hbm,
                            "note": "not the bound: operands live in registers / shared memory, the kernels are integer-issue bound "
                                    "(north_star asks for the INT-pipe roofline; MEASURED_PEAKS.json has no integer figure, so it is probed live)"}}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "u32", "data": "synthetic",
               "config": config_dict(args.config, cfg, n),
               "details": {"edges": edges_total, "edges_rank0": n_edges_part, "edges_by_rank": edges_ranks, "ms_per_step_by_rank": ms_ranks,
                           "e2e_ms_per_step_by_rank": e2e_ms_ranks, "edge_mode": mode, "l2": "flushed between timed iterations (256 MB write)",
                           "partition": "array replicated; work units of the sorted orders dealt round-robin to ranks; no data-path collective"},
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
               "reads_per_s": cfg["reads"] * args.steps / (ms_total * 1e-3)}
    if rank == 0 and world == 1 and not args.no_secondary:
        out["pipeline"] = whole_pipeline(t, wl, obs, valid, cfg)
        out["stages"] = other_stages(torch, dev, stream, L, s, args)
        out["c2"] = c2_secondary(torch, L, dev, stream, probe)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_sample(s, t, args.cpu_seconds)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def c2_secondary(torch, L, dev, stream, probe):
    """The round-1 headline for continuity: C2 (1 M reads, t = 1) edge construction, the whole correction step and the
    drop-in command file to file."""
    import badger_b200
    cfg, wl, obs, valid, s = dataset("C2")
    t = cfg["threshold"]
    n = int(s.size)
    badger_b200._lib.check(L.bdg_set_edge_mode(-1))
    step = EdgeStep(torch, L, dev, stream, s, t, 0, 1)
    edges = step.settle(3)
    k = 10
    ms = step.timed(k) / k
    st = step.stats()
    alg = algorithmic("sparse", t, st, edges)
    return {"config": config_dict("C2", cfg, n), "ms_per_step": ms, "value": n * (n - 1) // 2 / (ms * 1e-3), "unit": UNIT, "edges": edges,
            "edge_mode": "sparse", "roofline_frac": alg / (ms * 1e-3) / 1e12 / probe["lop3_imad_mix"], "work": st,
            "pipeline": whole_pipeline(t, wl, obs, valid, cfg), "cli": cli_file_to_file(t, wl, obs, valid, cfg)}


def whole_pipeline(t, wl, obs, valid, cfg):
    """reads/s of the array form of the whole correction step (badger_b200.pipeline.assign_packed: dedup, edges, centre
    selection with whitelist membership, clustering rounds, per-read gather), host arrays in and out, wall clock."""
    from badger_b200 import pipeline
    wls = np.sort(wl)
    pipeline.assign_packed(obs, valid, threshold=t, n_cells=cfg["n_cells"], whitelist_sorted=wls, form="u32")      # warm
    reps, T = 3, {}
    t0 = time.perf_counter()
    for _ in range(reps):
        out, info = pipeline.assign_packed(obs, valid, threshold=t, n_cells=cfg["n_cells"], whitelist_sorted=wls, timings=T, form="u32")
    dt = (time.perf_counter() - t0) / reps
    return {"reads_per_s": obs.size / dt, "ms": dt * 1e3, "stages_ms": {k: 1e3 * v / reps for k, v in T.items()}, **info,
            "api": "badger_b200.pipeline.assign_packed(form='u32') (packed barcode + valid byte per read in, centre + flag byte per read out)"}


def cli_file_to_file(t, wl, obs, valid, cfg):
    """reads/s of the drop-in command itself, file to file: `badger.py -r reads.tsv -l whitelist.txt -d 10x -t T --n_cells C
    -o OUT` run in this process on the workload's reads written as an extraction TSV (wall clock; the files are written
    before the clock starts and sit in the page cache).  Default route = native TSV reader -> GPU pack16 -> array pipeline
    -> native TSV writer; `--no_native_io` = pandas + the dict-shaped BarcodeGraph, the reference's own host structure."""
    import contextlib
    import importlib.util
    import io
    import logging
    import shutil
    tmp = tempfile.mkdtemp(prefix="bdg_cli_")
    try:
        tsv, wlf = os.path.join(tmp, "reads.tsv"), os.path.join(tmp, "wl.txt")
        synth.write_whitelist(wlf, wl)
        synth.write_extraction_tsv(tsv, obs, valid, synth.rng_for(5))
        spec = importlib.util.spec_from_file_location("badger_cli", os.path.join(ROOT, "badger.py"))
        cli = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(cli)
        cli.init = lambda *a, **k: 1                      # the bench already holds the device; keep its context
        argv = ["-r", tsv, "-l", wlf, "-d", "10x", "-t", str(t), "--n_cells", str(cfg["n_cells"])]

        def run(extra, out):
            sink = io.StringIO()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sink):
                cli.main(argv + ["-o", os.path.join(tmp, out)] + extra)
            dt = time.perf_counter() - t0
            for h in list(logging.getLogger("BarcodeGraph").handlers):
                logging.getLogger("BarcodeGraph").removeHandler(h)
            return dt, "reading through pandas" in sink.getvalue()

        run([], "WARM")
        reps = 3
        dts = [run([], "NATIVE") for _ in range(reps)]
        assert not any(fell_back for _, fell_back in dts), "the native TSV reader declined the bench's own file"
        dt = sum(d for d, _ in dts) / reps
        dt_pandas, _ = run(["--no_native_io"], "PANDAS")
        with open(os.path.join(tmp, "NATIVE_output_file.tsv"), "rb") as a, open(os.path.join(tmp, "PANDAS_output_file.tsv"), "rb") as b:
            same = a.read() == b.read()
        assert same, "the two routes of badger.py wrote different output files"
        return {"reads_per_s": obs.size / dt, "ms": dt * 1e3, "reads": int(obs.size),
                "bytes_in": os.path.getsize(tsv) + os.path.getsize(wlf), "bytes_out": os.path.getsize(os.path.join(tmp, "NATIVE_output_file.tsv")),
                "command": "badger.py -r reads.tsv -l whitelist.txt -d 10x -t %d --n_cells %d -o OUT (in-process, files in the page cache)" % (t, cfg["n_cells"]),
                "pandas_route": {"reads_per_s": obs.size / dt_pandas, "ms": dt_pandas * 1e3, "flag": "--no_native_io"},
                "outputs_identical": same}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def other_stages(torch, dev, stream, L, s, args):
    """The other rows of the hot path, each timed on its own: device-resident inputs and CUDA events
    for the kernels that have a device entry point, wall clock through the host-buffer call otherwise."""
    import ctypes as C
    import badger_b200
    from badger_b200 import ops
    chk = badger_b200._lib.check
    hbm = 6451.8
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_src = "MEASURED_PEAKS.json value recorded in BASELINE.md (file absent on this box)"
    rng = np.random.default_rng(11)
    res = {}

    def dev_time(fn, reps=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    # a-1 pack16: 16 B in + 5 B out per read
    R = 8_000_000
    seqs = torch.from_numpy(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, R * 16)].copy()).to(dev)
    d_out = torch.empty(R, dtype=torch.int32, device=dev); d_ok = torch.empty(R, dtype=torch.uint8, device=dev)
    dt = dev_time(lambda: chk(L.bdg_dev_pack16(seqs.data_ptr(), R, d_out.data_ptr(), d_ok.data_ptr(), stream.cuda_stream)))
    res["pack16"] = {"reads_per_s": R / dt, "GBps": 21 * R / dt / 1e9, "frac_of_hbm": 21 * R / dt / 1e9 / hbm, "n": R, "bytes_per_read": 21}
    # a-6 membership: 4 B in + 1 B out per query, 3 M-entry whitelist resident
    wl = synth.sorted_unique(rng.integers(0, 1 << 32, 3_000_000, dtype=np.uint64).astype(np.uint32))
    Q = 8_000_000
    q = rng.integers(0, 1 << 32, Q, dtype=np.uint64).astype(np.uint32)
    q[::3] = wl[rng.integers(0, wl.size, q[::3].size)]
    d_wl = torch.from_numpy(wl.view(np.int32)).to(dev); d_q = torch.from_numpy(q.view(np.int32)).to(dev)
    d_hit = torch.empty(Q, dtype=torch.uint8, device=dev)
    dt = dev_time(lambda: chk(L.bdg_dev_member_sorted(d_wl.data_ptr(), wl.size, d_q.data_ptr(), Q, d_hit.data_ptr(), stream.cuda_stream)))
    res["member_sorted"] = {"queries_per_s": Q / dt, "GBps": 5 * Q / dt / 1e9, "frac_of_hbm": 5 * Q / dt / 1e9 / hbm, "n": Q, "whitelist": int(wl.size),
                            "note": "random probes of a 12 MB sorted table: L2-latency bound, not HBM bound"}
    # a-7 post-processing: Q unassigned x W centres, first minimum of the plain edit distance, bound 2
    W7, Q7 = 10_000, 400_000
    tg = rng.integers(0, 1 << 32, W7, dtype=np.uint64).astype(np.uint32)
    q7 = s[rng.integers(0, s.size, Q7)]
    d_t = torch.from_numpy(tg.view(np.int32)).to(dev); d_q7 = torch.from_numpy(q7.view(np.int32)).to(dev)
    d_keys = torch.empty(Q7, dtype=torch.int32, device=dev); d_am = torch.empty(Q7, dtype=torch.int32, device=dev)
    d_di = torch.empty(Q7, dtype=torch.uint8, device=dev)
    dt = dev_time(lambda: chk(L.bdg_dev_nearest_bounded(d_q7.data_ptr(), Q7, d_t.data_ptr(), W7, 2, d_keys.data_ptr(), d_am.data_ptr(),
                                                        d_di.data_ptr(), stream.cuda_stream)))
    res["nearest_bounded"] = {"pairs_per_s": Q7 * W7 / dt, "queries": Q7, "targets": W7, "ms": dt * 1e3,
                              "note": "sorted/tiled form (same scan + tile kernels as the edges, bipartite)"}
    os.environ["BDG_NEAREST_DENSE"] = "1"
    dt = dev_time(lambda: chk(L.bdg_dev_nearest_bounded(d_q7.data_ptr(), Q7, d_t.data_ptr(), W7, 2, d_keys.data_ptr(), d_am.data_ptr(),
                                                        d_di.data_ptr(), stream.cuda_stream)))
    os.environ.pop("BDG_NEAREST_DENSE", None)
    res["nearest_bounded"]["brute_force_kernel"] = {"pairs_per_s": Q7 * W7 / dt, "ms": dt * 1e3, "int_Tinst_per_s": 11 * Q7 * W7 / dt / 1e12,
                                                    "note": "11 prefilter instructions per pair"}
    # a-2 dedup in first-seen order and a-5 k-mer scoring: host-buffer calls, wall clock (copies included)
    reads = s[rng.integers(0, s.size, 2_000_000)]
    ops.dedup_first_seen(reads, want_map=True)
    t0 = time.perf_counter()
    for _ in range(3):
        ops.dedup_first_seen(reads, want_map=True)
    dt = (time.perf_counter() - t0) / 3
    res["dedup_first_seen"] = {"reads_per_s": reads.size / dt, "n": int(reads.size), "ms": dt * 1e3, "timing": "host-buffer call, wall clock"}
    # a-5: the strings resident (ops.KmerIndex), 256 observed barcodes scored against the whitelist; both forms of the operator
    qk = s[:256]
    res["kmer_score"] = {"queries": int(qk.size), "whitelist": int(wl.size), "min_kmers": 4, "timing": "host-buffer call, wall clock"}
    for form, min_w in (("postings", "0"), ("scan", str(1 << 40))):
        os.environ["BDG_KMER_POST_MIN_W"] = min_w
        t0 = time.perf_counter()
        ix = ops.KmerIndex(wl)
        build = time.perf_counter() - t0
        hq = ix.query(qk, min_kmers=4)[0]
        t0 = time.perf_counter()
        for _ in range(3):
            ix.query(qk, min_kmers=4)
        dt = (time.perf_counter() - t0) / 3
        info = ix.info()
        ix.free()
        assert info["postings"] == (form == "postings")
        res["kmer_score"][form] = {"kernel_ms": info["kernel_ms"], "ms": dt * 1e3, "index_build_ms": build * 1e3, "queries_per_s": qk.size / dt,
                                   "pairs_covered_per_s": qk.size * wl.size / dt, "hits": int(hq.size)}
    os.environ.pop("BDG_KMER_POST_MIN_W", None)
    res["hbm_peak_GBps"] = hbm
    res["hbm_peak_source"] = hbm_src
    return res


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
