#!/usr/bin/env python
"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libaicam.so (cuobjdump -sass):
UTCHMMA (tcgen05.mma kind::f16 / tf32), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store),
UBLKCP (bulk copy), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus HMMA / FFMA for the kernels that do not use tcgen05.
   python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ai-camera_b200", "libaicam.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "FFMA", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "").replace("aicam::", "").replace("void ", "")
            cur = re.sub(r"\(.*", "", cur)
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            per[cur]["_total"] += 1
            if op in OPS:
                per[cur][op] += 1
    # template instantiations of one kernel are summed
    agg = collections.OrderedDict()
    for name, c in per.items():
        base = re.sub(r"<.*", "", name)
        a = agg.setdefault(base, [0, collections.Counter()])
        a[0] += 1
        a[1].update(c)
    print("SASS opcode histogram of ai-camera_b200/libaicam.so (sm_100a), per kernel (template instantiations summed)\n")
    print("%-34s %5s %9s " % ("kernel", "inst.", "SASS") + " ".join("%8s" % o for o in OPS))
    tot = collections.Counter()
    for name, (n, c) in agg.items():
        print("%-34s %5d %9d " % (name[:34], n, c["_total"]) + " ".join("%8d" % c[o] for o in OPS))
        tot.update(c)
    print("%-34s %5s %9d " % ("total", "", tot["_total"]) + " ".join("%8d" % tot[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
