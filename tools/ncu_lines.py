#!/usr/bin/env python3
"""Per source line of a profiled kernel: warp instructions executed and stall samples (ncu --set full --import-source on report).

  ncu_lines.py <file.ncu-rep> [top N]      prints the N heaviest source lines of the first profiled launch, by file
Reads only files; runs `ncu -i` locally (no GPU needed)."""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    cur_file, per_line, per_file = None, defaultdict(lambda: [0, 0, ""]), defaultdict(lambda: [0, 0])
    launches = 0
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            launches += 1
            if launches > 1 and cur_file is None:
                break
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            line = int(r[0])
            inst = int(r[hdr.index("Instructions Executed")])
            samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        key = (cur_file, line)
        per_line[key][0] += inst
        per_line[key][1] += samp
        per_line[key][2] = r[1].strip()
        per_file[cur_file][0] += inst
        per_file[cur_file][1] += samp
    total = sum(v[0] for v in per_file.values()) or 1
    tot_s = sum(v[1] for v in per_file.values()) or 1
    print("warp instructions executed %d, stall samples %d" % (total, tot_s))
    for f, (i, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-22s %5.1f%% of instructions  %5.1f%% of samples" % (f, 100.0 * i / total, 100.0 * s / tot_s))
    print("heaviest lines:")
    for (f, line), (i, s, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("  %-16s %4d  %5.1f%% inst %5.1f%% samp  %s" % (f, line, 100.0 * i / total, 100.0 * s / tot_s, src[:110]))


if __name__ == "__main__":
    main()
