#!/usr/bin/env python
"""Per-chain CUDA-event timings of the fused convolution chains (csrc/conv_chain.cu) at the shapes of one 64-stream
YOLOv8n step.  AICAM_CONV_TRACE=1 adds the event trace of CTA 0; AICAM_CHAIN_DEBUG=1 the chosen tiling.
   python scripts/chain_bench.py [name-substring ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ai_camera_b200 import _lib  # noqa: E402
from ai_camera_b200._lib import ChainDesc, check  # noqa: E402
import test_gpu_chain as T  # noqa: E402

B = int(os.environ.get("BATCH", "64"))
CASES = [
    # name, H, W, chain, unfused time of the same layers in profiles/r2_launches_step64_v1.txt (us, 64 frames)
    ("model2_full_c16", 160, 160, T.c2f_full(32, 16, 32), 225.4),
    ("model2_tail_c16", 160, 160, T.c2f_tail(16, 32), 181.8),
    ("model2_pair_c16", 160, 160, T.pair(16), 127.9),
    ("model4_pair_c32", 80, 80, T.pair(32), 60.9),
    ("model15_tail_c32", 80, 80, T.c2f_tail(32, 64, res=False), 84.2),
    ("model6_pair_c64", 40, 40, T.pair(64), 39.9),
    ("model12_pair_c64_nores", 40, 40, T.pair(64, res=False), 39.2),
    ("model8_pair_c128", 20, 20, T.pair(128), 40.6),
    ("head0_box_64", 80, 80, T.head(64, 64, 64), 131.0),
    ("head0_cls_80", 80, 80, T.head(64, 80, 80), 156.7),
    ("head1_box_64", 40, 40, T.head(128, 64, 64), 70.1),
    ("head1_cls_80", 40, 40, T.head(128, 80, 80), 80.5),
    ("head2_box_64", 20, 20, T.head(256, 64, 64), 45.3),
    ("head2_cls_80", 20, 20, T.head(256, 80, 80), 48.6),
]


def main():
    lib = _lib.load()
    sel = sys.argv[1:]
    for name, H, W, chain, unfused in CASES:
        if sel and not any(s in name for s in sel):
            continue
        d = ChainDesc()
        d.batch, d.h, d.w, d.in_c = B, H, W, chain["in_c"]
        d.nstages = len(chain["stages"])
        d.out_f32 = 1 if chain.get("out_f32") else 0
        for s, st in enumerate(chain["stages"]):
            e = d.st[s]
            e.cin, e.cout, e.ksize, e.act, e.nsrc = st["cin"], st["cout"], st["k"], st["act"], len(st["src"])
            for j, (b, co, cc) in enumerate(st["src"]):
                e.src_buf[j], e.src_coff[j], e.src_c[j] = b, co, cc
            if st.get("res"):
                e.res_buf, e.res_coff, e.res_mode = st["res"]
            else:
                e.res_buf, e.res_coff, e.res_mode = -1, 0, 0
        ms = C.c_double()
        try:
            check(lib.aicam_conv_chain_bench(C.byref(d), 20, C.byref(ms), None))
            print("%-26s %4dx%-4d x%d  %8.1f us   (single layers: %6.1f us)" % (name, H, W, B, ms.value * 1e3, unfused), flush=True)
        except _lib.AicamError as ex:
            print("%-26s %s" % (name, ex), flush=True)


if __name__ == "__main__":
    main()
