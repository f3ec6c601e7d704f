#!/usr/bin/env python
"""Per-layer micro-benchmark of the tcgen05 convolution on the layer shapes of the 64-stream
step (YOLOv8n at batch 64, ReID at ~1024 crops).  Prints time, TFLOP/s and minimum-traffic GB/s."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_camera_b200 import _lib  # noqa: E402
from ai_camera_b200._lib import ConvDesc  # noqa: E402

SHAPES = [
    # name, batch, h, w, cin, cout, k, s, act, res, f32
    ("yolo.stem 3->16 s2 @640", 64, 640, 640, 3, 16, 3, 2, 1, 0, 0),
    ("yolo.1 16->32 s2 @320", 64, 320, 320, 16, 32, 3, 2, 1, 0, 0),
    ("yolo.2.m 16->16 @160", 64, 160, 160, 16, 16, 3, 1, 1, 1, 0),
    ("yolo.2.cv1 1x1 32->32 @160", 64, 160, 160, 32, 32, 1, 1, 1, 0, 0),
    ("yolo.4.m 32->32 @80", 64, 80, 80, 32, 32, 3, 1, 1, 1, 0),
    ("yolo.6.m 64->64 @40", 64, 40, 40, 64, 64, 3, 1, 1, 1, 0),
    ("yolo.8.m 128->128 @20", 64, 20, 20, 128, 128, 3, 1, 1, 1, 0),
    ("yolo.22.cv2 64->64 @80", 64, 80, 80, 64, 64, 3, 1, 1, 0, 0),
    ("yolo.22.cv3 64->80 @80", 64, 80, 80, 64, 80, 3, 1, 1, 0, 0),
    ("yolo.22.cv3.1 80->80 @80", 64, 80, 80, 80, 80, 3, 1, 1, 0, 0),
    ("yolo.22.cv3 128->80 @40", 64, 40, 40, 128, 80, 3, 1, 1, 0, 0),
    ("yolo.15.cv2 1x1 96->64 @80", 64, 80, 80, 96, 64, 1, 1, 1, 0, 0),
    ("yolo.12.cv1 1x1 384->128 @40", 64, 40, 40, 384, 128, 1, 1, 1, 0, 0),
    ("yolo.22.cv3.2 1x1 80->80 f32 @80", 64, 80, 80, 80, 80, 1, 1, 0, 0, 1),
    ("reid.stem 3->64 @128x64", 1024, 128, 64, 3, 64, 3, 1, 2, 0, 0),
    ("reid.l1 64->64 @64x32", 1024, 64, 32, 64, 64, 3, 1, 2, 2, 0),
    ("reid.l2 128->128 @32x16", 1024, 32, 16, 128, 128, 3, 1, 2, 2, 0),
    ("reid.l3 256->256 @16x8", 1024, 16, 8, 256, 256, 3, 1, 2, 2, 0),
    ("reid.l4 512->512 @8x4", 1024, 8, 4, 512, 512, 3, 1, 2, 2, 0),
    ("reid.s2.l2 64->128 s2 @64x32", 1024, 64, 32, 64, 128, 3, 2, 2, 0, 0),
    ("reid.s2.l3 128->256 s2 @32x16", 1024, 32, 16, 128, 256, 3, 2, 2, 0, 0),
    ("reid.s2.l4 256->512 s2 @16x8", 1024, 16, 8, 256, 512, 3, 2, 2, 0, 0),
    ("reid.ds.l2 1x1 64->128 s2 @64x32", 1024, 64, 32, 64, 128, 1, 2, 0, 0, 0),
]

lib = _lib.load()
only = sys.argv[1] if len(sys.argv) > 1 else None
iters = int(os.environ.get("ITERS", "10"))
print("%-34s %9s %9s %9s" % ("layer", "ms", "TFLOP/s", "GB/s(min)"))
for (name, b, h, w, cin, cout, k, s, act, res, f32) in SHAPES:
    if only and only not in name:
        continue
    d = ConvDesc(b, h, w, cin, cout, k, s, act, res, f32)
    ms = C.c_double()
    _lib.check(lib.aicam_conv2d_bench(C.byref(d), iters, C.byref(ms), None))
    ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    flops = 2.0 * b * ho * wo * cout * cin * k * k
    cpad = 4 if cin <= 4 else cin
    byts = b * h * w * cpad * 2 + b * ho * wo * cout * (4 if f32 else 2) * (2 if res else 1)
    print("%-34s %9.3f %9.1f %9.0f" % (name, ms.value, flops / ms.value / 1e9, byts / ms.value / 1e6))
