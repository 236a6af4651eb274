#!/usr/bin/env python3
"""bench.py -- barcode pairs scored per second on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input: edge construction
(index.py:77-93 + barcode_graph.py:224-249 of the reference) over every unordered pair of the distinct
barcodes of BASELINE.json's config 2 (1 M simulated ONT reads, 10 k cells, 3 M whitelist, threshold 1).

  value   whole-job pairs/s with the sorted distinct-barcode array already resident in HBM
          (bdg_dev_edges_build on torch's current stream, CUDA events around every step, max over ranks)
  e2e     the same metric through the public host-buffer call (ops.edges_build_part -> bdg_edges_build_part):
          pinned host input -> H2D -> kernel -> D2H of the edge list, wall clock, max over ranks
  roofline    dominant kernel (the edge kernel's passes) against the MEASURED integer issue rate of this GPU
  cpu_baseline  the oracle's restatement of the reference algorithm on the box's host cores (rank 0, N=1)

N > 1 (torchrun): weak scaling - the read count is raised until the distinct barcodes are sqrt(N) times the one-GPU
count, so that the pairs per GPU stay fixed; rows are dealt to ranks in 2048-row tiles, no data-path collective
(SURVEY.md §8e).

`--impl reference` times the reference's own algorithm (oracle port, all host threads) on a bounded sample
of the same workload; /root/reference (pure Python) cannot travel to the GPU box.
"""
from __future__ import annotations

import argparse
import json
import math
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
# from the SASS of the inner loops, the numbers of units come from the kernel's own counters, bdg_dev_edges_stats):
#   sparse: one key-interval test (pass_possible) per column sub-tile, 32 more per surviving sub-tile; one quick
#           test per pair of a surviving 32x32 block; one exact stage (D, then S) per candidate
#   dense:  one light-loop step per pair (1.25 instr at t=1, 6 at t=2), exact stage per candidate
A_TILE = 40
A_QUICK = {1: 14.5, 2: 16.5}       # 58 / 66 SASS instructions per 4 pairs
A_LIGHT = {1: 1.25, 2: 6.0}
A_EXACT = 110                      # loads, un-rotation, pass predicates, dist_small
A_SCORE = 220                      # qgram_score, only for pairs with D <= t
# DRAM traffic of one step's dominant kernels from the ncu --set full capture of this round (profiles/r1e_sparse_t*_ncu_full.txt:
# dram__bytes_read.sum + dram__bytes_write.sum of the scan + tile kernels of all passes, C2, N = 492 093).  Writes stay in L2.
NCU_TRAFFIC_BYTES = {1: 536064 + 2174976 + 535040 + 2056704, 2: 3 * 535808 + 2710528 + 2709760 + 2153728 + 1280}
A_PAIR_SURVEY = {1: 15, 2: 25}   # SURVEY.md §8(d) nominal per-pair figure of a plain all-pairs kernel, reported alongside


# read counts at which the distinct barcodes of a config reach sqrt(N) times the one-GPU count (deterministic synthesis:
# found once by workload()'s search; used as its first guess and re-verified there, so that N > 1 runs start quickly)
WEAK_SCALING_READS = {("C2", 1_000_000, 0.05, 10_000): {"n1": 492093, 2: 1490288, 4: 2248893, 8: 3527307}}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reads", type=int, default=None, help="override the read count (testing)")
    ap.add_argument("--threshold", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", choices=["sparse", "dense"], default="sparse", help="edge search strategy (bdg_set_edge_mode)")
    return ap.parse_args()


def workload(args, world):
    """C2 at one GPU.  For N GPUs (weak scaling) the read count is raised until the number of DISTINCT barcodes is
    sqrt(N) times the one-GPU count, so that the pairs per GPU stay fixed (distinct barcodes grow slower than reads:
    sqrt(N) times the reads alone would hand every rank less work than the one-GPU run has)."""
    cfg = dict(synth.CONFIGS[args.config])
    if args.threshold is not None:
        cfg["threshold"] = args.threshold

    def distinct_of(reads):
        wl, cells, obs, valid, _ = synth.make_dataset(cfg, reads=reads)
        return synth.sorted_unique(obs[valid]), (wl, obs, valid)

    base_reads = args.reads if args.reads is not None else cfg["reads"]
    reads = base_reads
    known = WEAK_SCALING_READS.get((args.config, base_reads, cfg["perr"], cfg["n_cells"]))   # found once by the search below
    if world == 1 or not known:
        s, data = distinct_of(reads)
    if world > 1:
        n1 = known["n1"] if known else s.size
        target = n1 * math.sqrt(world)
        r_prev, n_prev = base_reads, n1
        reads = known.get(world, 0) if known else 0
        reads = reads or int(round(base_reads * math.sqrt(world)))
        for _ in range(3):
            s, data = distinct_of(reads)
            if abs(s.size - target) <= 0.012 * target:
                break
            alpha = math.log(s.size / n_prev) / math.log(reads / r_prev) if reads != r_prev and s.size != n_prev else 0.8
            alpha = min(max(alpha, 0.3), 1.0)
            r_prev, n_prev = reads, s.size
            reads = int(round(reads * (target / s.size) ** (1.0 / alpha)))
    name = "%s: %d simulated ONT reads, %d cells, %d-entry whitelist, %.0f%% error, threshold %d" % (
        args.config, reads, cfg["n_cells"], cfg["whitelist"], 100 * cfg["perr"], cfg["threshold"])
    workload.dataset = (data[0], data[1], data[2], cfg)          # for the whole-pipeline reads/s figure
    return s, cfg["threshold"], reads, name


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(s, t, seconds, seed=7):
    """Oracle port of the reference algorithm (6-mer index walk + 3-way verify) on a bounded row sample."""
    from oracle import oracle as orc
    # all the host cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which is not a property of the box)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or orc.num_threads())
    ix = orc.Index(s)
    rng = np.random.default_rng(seed)
    n = s.size
    k = min(n, 2000)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    s, t, reads, name = workload(args, world)
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
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": name, "reads": reads, "distinct": int(s.size), "threshold": t},
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

    s, t, reads, name = workload(args, world)
    n = int(s.size)
    total_pairs = n * (n - 1) // 2
    my_pairs = int(L.bdg_part_pairs(n, rank, world))
    stream = torch.cuda.current_stream()

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

    # ---- integer-pipe probe: the roofline denominator (not in MEASURED_PEAKS.json)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    import ctypes as C
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

    # ---- device-resident buffers (torch owns memory and stream; the library only launches)
    d_sorted = torch.from_numpy(s.view(np.int32)).to(dev)
    cap = max(1 << 16, 16 * n // world + 1024)
    d_a = torch.empty(cap, dtype=torch.int32, device=dev)
    d_b = torch.empty(cap, dtype=torch.int32, device=dev)
    d_d = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_dev():
        badger_b200._lib.check(L.bdg_dev_edges_build(d_sorted.data_ptr(), n, t, rank, world, d_a.data_ptr(), d_b.data_ptr(),
                                                     d_d.data_ptr(), cap, d_count.data_ptr(), stream.cuda_stream))

    def timed_steps(k):
        ms = 0.0
        for _ in range(k):
            flush.fill_(1)                      # evict L2 between timed iterations
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            step_dev()
            e1.record(stream)
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms

    def work_stats():
        v = (C.c_ulonglong * 5)()
        badger_b200._lib.check(L.bdg_dev_edges_stats(v, stream.cuda_stream))
        return dict(zip(("sub_tiles", "sub_tiles_scored", "pairs_scored", "candidates", "pairs_S_scored"), (int(x) for x in v)))

    sampler = ClockSampler(local)
    sampler.start()                        # covers warm-up, the timed steps and the e2e loop (a step is ~1 ms)
    badger_b200._lib.check(L.bdg_set_edge_mode(1 if args.mode == "sparse" else 0))
    step_dev()
    torch.cuda.synchronize()
    if int(d_count.item()) > cap:          # dense data (t >= 2): size the edge buffers from the first count
        cap = int(d_count.item()) + 1024
        d_a = torch.empty(cap, dtype=torch.int32, device=dev)
        d_b = torch.empty(cap, dtype=torch.int32, device=dev)
        d_d = torch.empty(cap, dtype=torch.uint8, device=dev)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    torch.cuda.synchronize()
    n_edges_part = int(d_count.item())
    assert 0 <= n_edges_part <= cap, "edge buffer or tile list too small: count %d, capacity %d" % (n_edges_part, cap)

    launches0 = L.bdg_launch_count()
    barrier()
    ms = timed_steps(args.steps)
    barrier()
    launches = L.bdg_launch_count() - launches0
    stats = work_stats()
    ms_total = max_over_ranks(ms)
    step_ms = ms / args.steps             # this rank's average step (all launches of the step, CUDA events)
    value = total_pairs * args.steps / (ms_total * 1e-3)

    # ---- the other search strategy, for the record (same buffers, same timing rules)
    other = "dense" if args.mode == "sparse" else "sparse"
    badger_b200._lib.check(L.bdg_set_edge_mode(0 if args.mode == "sparse" else 1))
    for _ in range(3):
        step_dev()
    torch.cuda.synchronize()
    n_other = int(d_count.item())          # per-part counts differ between the modes (different row orders), totals must not
    if world > 1:
        tt = torch.tensor([n_edges_part, n_other], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        assert int(tt[0].item()) == int(tt[1].item()), "the two edge modes disagree on the total edge count"
    else:
        assert n_other == n_edges_part, "the two edge modes disagree on the edge count"
    ms_other = timed_steps(max(2, min(args.steps, 3))) / max(2, min(args.steps, 3))
    stats_other = work_stats()
    badger_b200._lib.check(L.bdg_set_edge_mode(1 if args.mode == "sparse" else 0))

    # ---- e2e: public host-buffer API, pinned input, edge list back on the host
    s_pinned = torch.from_numpy(s.view(np.int32)).pin_memory()
    s_host = s_pinned.numpy().view(np.uint32)
    r1 = ops.edges_build_part(s_host, t, rank, world)      # warm; two result sets alive at once, as in the timed loop below,
    r2 = ops.edges_build_part(s_host, t, rank, world)      # so that the operator's pinned output pool holds both of them
    del r1, r2
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ea, eb, ed = ops.edges_build_part(s_host, t, rank, world)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop()
    e2e = {"value": total_pairs * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(4 * n),
           "d2h_bytes_per_step": int(9 * ea.size + 8), "ms_per_step": 1000 * e2e_s / args.steps,
           "api": "badger_b200.ops.edges_build_part -> bdg_edges_build_part (host buffers)"}

    edges_total = n_edges_part
    if world > 1:
        tt = torch.tensor([n_edges_part], dtype=torch.int64, device=dev)
        dist.all_reduce(tt)
        edges_total = int(tt.item())

    def algorithmic(mode, st):
        if t not in A_QUICK:
            return None
        per_pair = A_QUICK[t] if mode == "sparse" else A_LIGHT[t]
        tiles = (st["sub_tiles"] + 32 * st["sub_tiles_scored"]) if mode == "sparse" else 0
        score = st["pairs_S_scored"] if mode == "sparse" else n_edges_part     # dense counts S inside the exact stage: at least the edges
        return A_TILE * tiles + per_pair * st["pairs_scored"] + A_EXACT * st["candidates"] + A_SCORE * score

    out = None
    if rank == 0:
        peak = probe["lop3_imad_mix"]
        roof = None
        alg = algorithmic(args.mode, stats)
        if alg is not None:
            achieved = alg / (step_ms * 1e-3) / 1e12
            alg_o = algorithmic(other, stats_other)
            roof = {"bound": "int_issue",
                    "kernel": ("sparse_scan_kernel + sparse_tile_kernel<%d,p>, %d passes per step" % (t, 2 if t == 1 else 3)) if args.mode == "sparse" else "edges_kernel<%d>" % t,
                    "achieved": achieved, "peak": peak, "unit": "Tinst/s", "frac": achieved / peak if peak else None,
                    "traffic": NCU_TRAFFIC_BYTES.get(t) if (args.mode == "sparse" and args.config == "C2" and args.reads is None and world == 1) else None,
                    "traffic_note": "bytes per step, dram read+write of the scan and tile kernels from the committed ncu capture (profiles/); the "
                                    "algorithmic bytes are hbm.algorithmic_bytes_per_launch",
                    "how": "achieved = algorithmic integer instructions of rank 0's step (%d interval tests x %d + %d pairs scored x %s + "
                           "%d candidates x %d + pairs S-scored x 220; unit counts from the kernel's own counters, per-unit costs from its SASS, DESIGN.md) / "
                           "%.3f ms (CUDA events, this run); peak = measured issue rate of an independent LOP3+IMAD 1:1 stream on this GPU "
                           "(bdg_dev_pipe_probe, this run)" % (stats["sub_tiles"] + 32 * stats["sub_tiles_scored"] if args.mode == "sparse" else 0,
                                                               A_TILE, stats["pairs_scored"], A_QUICK[t] if args.mode == "sparse" else A_LIGHT[t],
                                                               stats["candidates"], A_EXACT, step_ms),
                    "work": stats,
                    "pairs_decided_per_pair_scored": my_pairs / max(stats["pairs_scored"], 1),
                    "survey_nominal": {"ops_per_pair": A_PAIR_SURVEY[t], "achieved": A_PAIR_SURVEY[t] * my_pairs / (step_ms * 1e-3) / 1e12,
                                       "note": "SURVEY.md 8(d) costs every pair 5(2t+1) instructions; this kernel decides most pairs by key-interval "
                                               "exclusion, so this figure exceeds the peak by design"},
                    "other_mode": {"mode": other, "ms_per_step": ms_other, "pairs_per_s": my_pairs / (ms_other * 1e-3), "work": stats_other,
                                   "achieved": alg_o / (ms_other * 1e-3) / 1e12, "frac": alg_o / (ms_other * 1e-3) / 1e12 / peak if peak else None},
                    "probe_Tinst_per_s": probe,
                    "hbm": {"algorithmic_bytes_per_launch": int(4 * n * (2 if t == 1 else 3) + 9 * n_edges_part),
                            "achieved_GBps": (4 * n * (2 if t == 1 else 3) + 9 * n_edges_part) / (step_ms * 1e-3) / 1e9,
                            "peak_GBps": 6451.8, "frac": (4 * n * (2 if t == 1 else 3) + 9 * n_edges_part) / (step_ms * 1e-3) / 1e9 / 6451.8,
                            "note": "not the bound: operands live in registers / shared memory, the kernels are integer-issue bound "
                                    "(north_star asks for the INT-pipe roofline; MEASURED_PEAKS.json has no integer figure, so it is probed live)"}}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u32", "data": "synthetic",
               "config": {"workload": name, "reads": reads, "distinct": n, "threshold": t, "pairs_per_step": total_pairs,
                          "edges": edges_total, "edge_mode": args.mode, "l2": "flushed between timed iterations (256 MB write)",
                          "partition": "2048-row tiles dealt boustrophedon to ranks; no data-path collective"},
               "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
               "reads_per_s": reads * args.steps / (ms_total * 1e-3)}
    if rank == 0 and world == 1:
        out["stages"] = other_stages(torch, dev, stream, L, s, args)
        out["pipeline"] = whole_pipeline(t)
        out["cli"] = cli_file_to_file(t)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_sample(s, t, args.cpu_seconds)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def whole_pipeline(t):
    """reads/s of the array form of the whole correction step (badger_b200.pipeline.assign_packed: dedup, edges, centre
    selection with whitelist membership, clustering rounds, per-read gather), host arrays in and out, wall clock."""
    from badger_b200 import pipeline
    wl, obs, valid, cfg = workload.dataset
    wls = np.sort(wl)
    pipeline.assign_packed(obs, valid, threshold=t, n_cells=cfg["n_cells"], whitelist_sorted=wls)      # warm
    reps, T = 3, {}
    t0 = time.perf_counter()
    for _ in range(reps):
        out, info = pipeline.assign_packed(obs, valid, threshold=t, n_cells=cfg["n_cells"], whitelist_sorted=wls, timings=T)
    dt = (time.perf_counter() - t0) / reps
    return {"reads_per_s": obs.size / dt, "ms": dt * 1e3, "stages_ms": {k: 1e3 * v / reps for k, v in T.items()}, **info,
            "api": "badger_b200.pipeline.assign_packed (packed barcodes per read in, centre per read out)"}


def cli_file_to_file(t):
    """reads/s of the drop-in command itself, file to file: `badger.py -r reads.tsv -l whitelist.txt -d 10x -t T --n_cells C
    -o OUT` run in this process on the workload's reads written as an extraction TSV (wall clock; the files are written
    before the clock starts and sit in the page cache).  Default route = native TSV reader -> GPU pack16 -> array pipeline
    -> native TSV writer; `--no_native_io` = pandas + the dict-shaped BarcodeGraph, the reference's own host structure."""
    import contextlib
    import importlib.util
    import io
    import logging
    import shutil
    wl, obs, valid, cfg = workload.dataset
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
    qk = s[:64]
    ops.kmer_score(qk, wl, min_kmers=4)
    t0 = time.perf_counter()
    ops.kmer_score(qk, wl, min_kmers=4)
    dt = time.perf_counter() - t0
    res["kmer_score"] = {"pairs_per_s": qk.size * wl.size / dt, "queries": int(qk.size), "whitelist": int(wl.size), "ms": dt * 1e3,
                         "timing": "host-buffer call, wall clock"}
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
