#!/usr/bin/env python
"""K1 (letterbox / preprocess) and K5 (ReID crops) micro-benchmark against the HBM roofline.
64 x 1080p frames per launch (398 MB BGR / 199 MB NV12: larger than L2), CUDA events around each launch, L2 flushed
between launches by a 512 MB memset; reports the algorithmic bytes over the median time and the fraction
of MEASURED_PEAKS.json's copy bandwidth.
    python scripts/k1_bench.py [frames] [iters]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ai_camera_b200 import _lib  # noqa: E402
from ai_camera_b200._lib import check, ptr  # noqa: E402
from ai_camera_b200.config import tracked_class_mask  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
H, W = 1080, 1920
dev = torch.device("cuda:0")
lib = _lib.load()
peak = 6547.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
gen = torch.Generator(device=dev).manual_seed(1)
bgr = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev, generator=gen)
nv12 = torch.randint(0, 256, (B, H * 3 // 2, W), dtype=torch.uint8, device=dev, generator=gen)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[3:]))


print("K1 / K5 micro-benchmark: %d frames %dx%d, %d launches each, peak %.0f GB/s (MEASURED_PEAKS.json hbm_gbs)" % (B, H, W, iters, peak))
for fmt, name, out_bytes in ((3, "4x4 pixel blocks bf16 (engine input)", 640 * 640 * 8), (2, "2x2 pixel blocks bf16", 640 * 640 * 8), (1, "NHWC4 bf16", 640 * 640 * 8),
                             (0, "NCHW fp32 (reference layout)", 640 * 640 * 12)):
    out = torch.empty(B * out_bytes, dtype=torch.uint8, device=dev)
    for src, frames, fn, row_bytes in (("BGR", bgr, lib.aicam_preprocess, W * 3), ("NV12", nv12, lib.aicam_preprocess_nv12, W * 2)):
        ms = timed(lambda: check(fn(ptr(frames), B, H, W, fmt, ptr(out), None)))
        alg = B * (360 * row_bytes + out_bytes)
        print("K1 %-4s -> %-36s %7.1f us  %6.0f GB/s algorithmic = %4.1f %% of peak" % (src, name, ms * 1e3, alg / ms / 1e6, 100 * alg / ms / 1e6 / peak))

# K5: ~16 person-sized boxes per frame
K = 100
n_per = 16
boxes = torch.zeros((B, K, 4), dtype=torch.float32, device=dev)
cx = torch.rand((B, n_per), device=dev, generator=gen) * 1500 + 200
cy = torch.rand((B, n_per), device=dev, generator=gen) * 700 + 200
hh = torch.rand((B, n_per), device=dev, generator=gen) * 200 + 150
ww = hh * 0.4
boxes[:, :n_per] = torch.stack([cx - ww / 2, cy - hh / 2, cx + ww / 2, cy + hh / 2], dim=-1)
scores = torch.full((B, K), 0.9, dtype=torch.float32, device=dev)
labels = torch.zeros((B, K), dtype=torch.int32, device=dev)
num = torch.full((B,), n_per, dtype=torch.int32, device=dev)
max_crops = B * n_per
det_index = torch.zeros((B, K), dtype=torch.int32, device=dev)
det_count = torch.zeros((B,), dtype=torch.int32, device=dev)
crop_slot = torch.zeros((B, K), dtype=torch.int32, device=dev)
crop_rect = torch.zeros((max_crops, 5), dtype=torch.int32, device=dev)
crops = torch.empty((max_crops, 128, 64, 8), dtype=torch.bfloat16, device=dev)
crop_count = torch.zeros((2,), dtype=torch.int32, device=dev)
lo, hi = tracked_class_mask()
src_bytes = float((ww * hh).sum().item()) * 3  # BGR bytes under the boxes
for src, frames, fn in (("BGR", bgr, lib.aicam_reid_crops), ("NV12", nv12, lib.aicam_reid_crops_nv12)):
    ms = timed(lambda: check(fn(ptr(frames), B, H, W, ptr(boxes), ptr(scores), ptr(labels), ptr(num), K, 0.3, lo, hi, 2, max_crops,
                                ptr(det_index), ptr(det_count), ptr(crop_slot), ptr(crop_rect), ptr(crops), ptr(crop_count), None)))
    alg = src_bytes * (1.0 if src == "BGR" else 0.5) + max_crops * 128 * 64 * 16
    print("K5 %-4s filter + %d crops -> NHWC8 bf16            %7.1f us  %6.0f GB/s algorithmic = %4.1f %% of peak (two launches)" %
          (src, max_crops, ms * 1e3, alg / ms / 1e6, 100 * alg / ms / 1e6 / peak))
