#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/<name>.txt
  python scripts/summarize_ncu.py report   gpurun_out/prof.ncu-rep  > profiles/<name>.txt"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sm__inst_executed.avg.per_cycle_active", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
            rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("aicam::<unnamed>::", ""), r["Grid Size"], v))
    tot = sum(r[2] for r in rows)
    print("launches %d, summed kernel time %.1f us (cold-cache, serialised: compare shares)" % (len(rows), tot))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, g, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-34s n=%3d %10.1f us %6.1f%%" % (k[:34], n, v, 100 * v / tot))
    print("\nper launch, in order:")
    for i, (k, g, v) in enumerate(rows):
        print("%3d %-28s grid %-16s %9.1f us" % (i, k[:28], g, v))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")][:100], "grid", vals[hdr.index("Grid Size")],
              "block", vals[hdr.index("Block Size")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("  %-75s %s %s" % (k, vals[i], units[i]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr, data = rows[1], rows[2:]
        isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        tot = sum(int(r[iex]) for r in data if r[iex].isdigit())
        print("\nwarp instructions executed: %d; top SASS lines by stall samples:" % tot)
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tsm = sum(int(r[ismp]) for r in data if r[ismp].isdigit())
        for r in sorted(data, key=lambda r: -int(r[ismp]) if r[ismp].isdigit() else 0)[:14]:
            st = sorted(((int(r[i]), hdr[i]) for i in stall if r[i].isdigit()), reverse=True)[:2]
            print("  %5.1f%%  executed %9s  %-46s %s" % (100 * int(r[ismp]) / max(1, tsm), r[iex], r[isrc].strip()[:46], st))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
