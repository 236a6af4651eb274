#!/usr/bin/env python3
"""Summarise ncu output into the small text files kept under profiles/.

  ncu_summary.py rep  <file.ncu-rep> [kernel-regex]   key metrics of every profiled launch (--set full capture)
  ncu_summary.py list <launches.csv> [first count]    per-kernel launch count, total and share of device time
                                                      (--metrics gpu__time_duration.sum capture); first / count keep
                                                      that window of the launches, e.g. the timed steps of bench.py
Reads only files; runs `ncu -i` locally (no GPU needed).
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def rep(path, pattern=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pattern and not re.search(pattern, d["Kernel Name"]):
            continue
        print("== %s  grid %s block %s  (launch id %s)" % (d["Kernel Name"], d.get("Grid Size"), d.get("Block Size"), d.get("ID")))
        for k in KEYS:
            if k in d:
                print("  %-92s %s %s" % (k, d[k], units[hdr.index(k)]))


def launches(path, first=0, count=None):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r.get("Metric Name") == "gpu__time_duration.sum"]
    if first or count is not None:
        print("launches %d .. %d of %d" % (first, first + (count if count is not None else len(rows) - first) - 1, len(rows)))
        rows = rows[first:first + count] if count is not None else rows[first:]
    agg = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        ns = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print("%-70s %7s %12s %12s %7s  %s" % ("kernel", "launches", "total_ms", "avg_ms", "share", "grid x block"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %7d %12.3f %12.4f %6.1f%%  %s x %s" % (name[:70], a[0], a[1] / 1e6, a[1] / a[0] / 1e6, 100 * a[1] / total, a[2], a[3]))
    print("%-70s %7d %12.3f" % ("TOTAL", sum(a[0] for a in agg.values()), total / 1e6))


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("rep", "list"):
        sys.exit(__doc__)
    if sys.argv[1] == "rep":
        rep(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0, int(sys.argv[4]) if len(sys.argv) > 4 else None)
