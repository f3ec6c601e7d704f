#!/usr/bin/env python
"""BASELINE.json configs[4]: crowded scene - D persons per frame (default 300) per stream, so T -> D live tracks, galleries
filling to the budget (G = 100), a D x D association per stream.  Times the device tracker step (K5 filter + K7-K12:
normalize, appearance = gallery cosine distance, assoc = Kalman / gating / cascade LSAP / lifecycle) on planted boxes and
features (the ReID net is timed elsewhere), steady state after the galleries are full.
    python scripts/crowded_bench.py [streams] [persons] [steps]"""
import ctypes as C
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
import gpu_util as G  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
D = int(sys.argv[2]) if len(sys.argv) > 2 else 300
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 140
F, GAL = 512, 100
lib = _lib.load()
dev = G.DEV
K = D
trk = G.Tracker(S, max_tracks=D + 84, max_dets=D, feature_dim=F, stride_k=K)
gen = torch.Generator(device=dev).manual_seed(7)
cx = torch.rand((S, D), device=dev, generator=gen) * 1700 + 100
cy = torch.rand((S, D), device=dev, generator=gen) * 900 + 90
vx = torch.randn((S, D), device=dev, generator=gen) * 0.8
vy = torch.randn((S, D), device=dev, generator=gen) * 0.4
h = torch.rand((S, D), device=dev, generator=gen) * 60 + 60
w = h * 0.4
base = torch.nn.functional.normalize(torch.randn((S, D, F), device=dev, generator=gen), dim=-1)
trk.scores.fill_(0.9)
trk.labels.zero_()
trk.num.fill_(D)
lo, hi = tracked_class_mask()
st = None
times = []
for t in range(steps):
    x, y = cx + vx * t, cy + vy * t
    jit = torch.randn((S, D, 4), device=dev, generator=gen)
    trk.boxes.copy_(torch.stack([x - w / 2, y - h / 2, x + w / 2, y + h / 2], dim=-1) + jit)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # planted features land in the crop rows the filter assigns: here every detection is kept in order, row = s * D + d
    feats = base + 0.02 * torch.randn((S, D, F), device=dev, generator=gen)
    trk.feats.copy_(feats.reshape(S * D, F))
    torch.cuda.synchronize()
    e0.record()
    check(lib.aicam_reid_crops(None, S, 1080, 1920, ptr(trk.boxes), ptr(trk.scores), ptr(trk.labels), ptr(trk.num), K, 0.3, lo, hi,
                               1, S * K, ptr(trk.det_index), ptr(trk.det_count), ptr(trk.crop_slot), ptr(trk.crop_rect), None,
                               ptr(trk.crop_count), st))
    check(lib.aicam_tracker_step(trk.h, ptr(trk.boxes), ptr(trk.scores), ptr(trk.labels), K, ptr(trk.det_index), ptr(trk.det_count),
                                 ptr(trk.crop_slot), ptr(trk.feats), ptr(trk.out_tracks), ptr(trk.out_conf), ptr(trk.out_count), st))
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    if os.environ.get("PROFILE_LAST") and t == steps - 2:
        torch.cuda.cudart().cudaProfilerStart()
    if os.environ.get("PROFILE_LAST") and t == steps - 1:
        torch.cuda.cudart().cudaProfilerStop()
assert not trk.overflow().any(), "tracker capacity exceeded"
reported = trk.out_count.float().mean().item()
steady = float(np.median(times[-20:]))
gal_bytes = S * reported * GAL * F * 4
print("crowded scene: %d streams x %d persons, %d steps; tracks reported per stream %.1f" % (S, D, steps, reported))
print("tracker step (filter + normalize + appearance + assoc), steady state: %.3f ms  (first frames %.3f ms)" % (steady, times[1]))
print("gallery read per step %.2f GB -> >= %.0f GB/s if it were the whole step; association problems %d x %d per stream" %
      (gal_bytes / 1e9, gal_bytes / steady / 1e6, int(reported), D))
print("per-stream-frame: %.1f us" % (1e3 * steady / S))
trk.close()
