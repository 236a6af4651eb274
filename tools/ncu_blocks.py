#!/usr/bin/env python3
"""Basic-block view of an ncu source page: consecutive SASS instructions with the same execution count are
folded into one line (first instruction, length, warp-level executions, share of all issued instructions).
usage: ncu_blocks.py <file.ncu-rep> [min_share_percent] [launch index in the report]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
start = starts[which]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(lines)
print(lines[start - 1][:120])
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
tot = sum(int(r["Instructions Executed"]) for r in rows)
ttot = sum(int(r["Thread Instructions Executed"]) for r in rows)
print("total warp instructions %d, thread instructions %d (avg %.1f threads)" % (tot, ttot, ttot / max(tot, 1)))
blocks = []
cur = None
for i, r in enumerate(rows):
    n = int(r["Instructions Executed"]); th = int(r["Thread Instructions Executed"]); smp = int(r["# Samples"])
    if cur and cur["n"] == n:
        cur["len"] += 1; cur["th"] += th; cur["smp"] += smp; cur["ops"].append(r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1])
    else:
        cur = {"i": i, "n": n, "len": 1, "th": th, "smp": smp, "first": r["Source"].strip(), "ops": [r["Source"].split()[0]]}
        blocks.append(cur)
ssum = sum(b["smp"] for b in blocks)
for b in blocks:
    share = 100.0 * b["n"] * b["len"] / tot
    if share >= minshare:
        from collections import Counter
        top = ", ".join("%s x%d" % kv for kv in Counter(b["ops"]).most_common(4))
        print("#%5d len %4d  exec %12d  share %5.1f%%  samples %5.1f%%  thr/inst %4.1f  | %s" % (
            b["i"], b["len"], b["n"], share, 100.0 * b["smp"] / max(ssum, 1), b["th"] / max(b["n"] * b["len"], 1), top))
