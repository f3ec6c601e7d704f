#!/usr/bin/env python
"""One step of the 64-stream pipeline between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/launches.csv python scripts/profile_step.py
Prints the step's event-timed duration when run without ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ai_camera_b200 import synth  # noqa: E402
from ai_camera_b200.pipeline import TrackingPipeline  # noqa: E402

S = int(os.environ.get("STREAMS", "64"))
shift = float(os.environ.get("LOGIT_SHIFT", "-0.93"))
yolo, reid = synth.make_blobs(synth.blob_dir())
video = synth.SynthVideo(S, (1080, 1920), n_frames=4, device="cuda:0")
pipe = TrackingPipeline(yolo, reid, S, "cuda:0", max_tracks=128, max_crops=S * 40)
for n in synth.CLS_LAYERS:
    pipe.detector.engine.set_bias(n, pipe.detector.engine.get_bias(n).numpy() + np.float32(shift))
for t in range(3):
    pipe.step(video.frames(t))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
pipe.step(video.frames(3))
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("step ms %.3f crops %d tracks %d" % (e0.elapsed_time(e1), pipe.tracker.crop_count[0].item(),
                                           pipe.tracker.out_count.sum().item()))
