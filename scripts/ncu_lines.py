#!/usr/bin/env python
"""Per-source-line instruction / stall / shared-memory-wavefront totals from an ncu report:
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv ; python scripts/ncu_lines.py /tmp/src.csv"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
per_line = collections.defaultdict(lambda: [0, 0, 0])
cur, hdr = None, None
for row in csv.reader(open(sys.argv[1])):
    if not row:
        continue
    if row[0] == "File Path":
        cur, hdr = row[1], None
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None or row[0] == "Function Name":
        continue
    d = dict(zip(hdr, row))
    if d.get("Address") != "-" or not d["Line No"].isdigit():
        continue
    num = lambda k: int(d[k]) if d.get(k, "").isdigit() else 0  # noqa: E731
    v = per_line[(cur, int(d["Line No"]))]
    v[0] += num("Instructions Executed")
    v[1] += num("# Samples")
    v[2] += num("L1 Wavefronts Shared")
tot = sum(v[0] for v in per_line.values())
smp = sum(v[1] for v in per_line.values())
print("warp instructions %d, stall samples %d" % (tot, smp))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for (fn, ln), v in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
    try:
        text = open(fn).read().split("\n")[ln - 1].strip()[:84]
    except OSError:
        text = ""
    print("%5.1f%% inst %5.1f%% smp %9d smem-wf  %s:%d  %s" % (100.0 * v[0] / tot, 100.0 * v[1] / max(smp, 1), v[2],
                                                             os.path.basename(fn), ln, text))
