#!/usr/bin/env python
"""BASELINE.json configs[3]: YOLOv8 n/s/m detection-only throughput over batch sizes (preprocess excluded:
the engine alone, NHWC4 bf16 input resident in HBM), against the tensor-pipe roofline.
    python scripts/yolo_sweep.py [scales] [max_batch]      e.g.  python scripts/yolo_sweep.py nsm 256"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_camera_b200 import _lib, synth, weights as W  # noqa: E402

scales = sys.argv[1] if len(sys.argv) > 1 else "nsm"
max_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lib = _lib.load()
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0)
except Exception:
    peak = 1400.0
dev = torch.device("cuda:0")
print("%-8s %6s %10s %12s %10s %8s" % ("model", "batch", "ms", "frames/s", "TFLOP/s", "of peak"))
for sc in scales:
    path = os.path.join(synth.blob_dir(), "yolov8%s_sweep.aicw" % sc)
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        kind, params, tensors = W.synth_yolov8_weights(sc, seed=0)
        W.write_blob(path, kind, params, tensors)
    e = C.c_void_p()
    _lib.check(lib.aicam_engine_create(path.encode(), 0, max_batch, C.byref(e)))
    flops = lib.aicam_engine_flops_per_item(e)
    b = 1
    while b <= max_batch:
        x = (torch.rand((b, 640, 640, 4), device=dev) * 0.5).to(torch.bfloat16)
        head = torch.empty((b, 8400, 144), dtype=torch.float32, device=dev)
        st = _lib.stream_ptr(dev)
        for _ in range(3):
            _lib.check(lib.aicam_yolo_forward(e, _lib.ptr(x), b, _lib.ptr(head), st))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10 if b <= 32 else 4
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            _lib.check(lib.aicam_yolo_forward(e, _lib.ptr(x), b, _lib.ptr(head), st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = flops * b / ms / 1e9
        print("yolov8%-2s %6d %10.3f %12.0f %10.1f %7.1f%%" % (sc, b, ms, b / ms * 1e3, tf, 100 * tf / peak))
        del x, head
        b *= 2
    lib.aicam_engine_destroy(e)
