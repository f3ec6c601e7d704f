#!/usr/bin/env python
"""In-graph timeline of one 64-stream step: which layer runs when INSIDE the CUDA graph (ncu serialises launches and
flushes caches, per-launch events break programmatic dependent launch; this does neither).  The step is captured with
aicam_debug_timeline on, so CTA 0 of every window-convolution launch stamps the GPU's global timer at entry, when its
dependency on the previous kernel has resolved and at exit; a replay of the graph then fills the records.
    python scripts/step_timeline.py [streams]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_camera_b200 import _lib, synth  # noqa: E402
from ai_camera_b200.pipeline import TrackingPipeline  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shift = float(os.environ.get("LOGIT_SHIFT", "-0.98"))
lib = _lib.load()
yolo, reid = synth.make_blobs(synth.blob_dir())
video = synth.SynthVideo(S, (1080, 1920), n_frames=4, device="cuda:0")
pipe = TrackingPipeline(yolo, reid, S, "cuda:0", max_tracks=128, max_crops=S * 40)
for n in synth.CLS_LAYERS:
    pipe.detector.engine.set_bias(n, pipe.detector.engine.get_bias(n).numpy() + np.float32(shift))
for t in range(6):
    pipe.step(video.frames(t % 4))
torch.cuda.synchronize()
CAP = 256
_lib.check(lib.aicam_debug_timeline(1, None, CAP))
stream = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
frames = video.frames(2)
with torch.cuda.stream(stream):
    with torch.cuda.graph(g, stream=stream):
        pipe.step(frames)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
buf = (C.c_longlong * (4 * CAP))()
n = lib.aicam_debug_timeline(2, buf, CAP)
rec = np.frombuffer(buf, dtype=np.int64).reshape(CAP, 4)[:n].copy()
lib.aicam_debug_timeline(0, None, 0)
print("one graph replay: %.3f ms event-timed, %d crops; %d window-convolution launches recorded" %
      (e0.elapsed_time(e1), pipe.tracker.crop_count[0].item(), n))
print("times in us relative to the first entry; wait = entry -> dependency resolved (programmatic dependent launch: the CTA is")
print("resident, its prologue done, the previous kernel still draining); run = dependency -> exit of CTA 0; gap = previous exit -> this")
print("dependency (other kernels - pools, upsample, decode, crops, stem, pair kernel - and launch latency sit in the gaps)")
t0 = rec[0, 0]
prev_exit = None
tot_run = tot_gap = 0.0
print("%3s %22s %9s %8s %8s %8s" % ("#", "cout<-cin @height", "entry", "wait", "run", "gap"))
for i in range(n):
    ent, dep, ex, tag = rec[i]
    cout, cin, hh = tag // 1000000, (tag // 1000) % 1000, tag % 1000
    gap = (dep - prev_exit) / 1e3 if prev_exit is not None else 0.0
    run = (ex - dep) / 1e3
    tot_run += run
    tot_gap += max(gap, 0.0)
    print("%3d %10d<-%-4d @%-4d %9.1f %8.1f %8.1f %8.1f" % (i, cout, cin, hh, (ent - t0) / 1e3, (dep - ent) / 1e3, run, gap))
    prev_exit = ex
print("sum of run %.1f us, sum of positive gaps %.1f us, first entry -> last exit %.1f us" % (tot_run, tot_gap, (rec[n - 1, 2] - t0) / 1e3))
