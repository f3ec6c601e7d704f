#!/usr/bin/env python
"""Would an FP8 (E4M3, tcgen05 kind::f8f6f4) detector hold the 1e-2 box bar?  CPU simulation, no GPU needed.

The YOLOv8 oracle net (oracle/nets.py) is run three ways on the same seeded blob and images:
  fp32      : the oracle as is (the reference the GPU tests compare with);
  bf16      : activations rounded to bf16 between layers (what the shipped tcgen05 kernels do);
  fp8       : every convolution's input activation quantised to E4M3 with a per-tensor scale (amax / 448, the best case:
              scales taken from THIS input) and its weights to E4M3 with per-output-channel scales, fp32 accumulate;
  fp8-first : the same with the two Detect-head 1x1 layers and the stem kept in bf16 (a common mixed recipe).
Reported: head-logit error and the box / score error over the top-200 anchors, measured like tests/test_gpu_engine.py.
   python scripts/fp8_study.py > profiles/r2_fp8_study.txt"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ai_camera_b200 import weights as W  # noqa: E402
from oracle import detect_post, image_ops, nets  # noqa: E402
from scenarios import synth_image  # noqa: E402

E4M3_MAX = 448.0


def q_e4m3(x, scale):
    return (x / scale).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).float() * scale


class QuantNet(nets.YoloV8):
    def __init__(self, params, tensors, mode):
        super().__init__(params, tensors)
        self.mode = mode

    def conv(self, x, name, k, s, act=True):
        w, b = self.w[name + ".weight"], self.w[name + ".bias"]
        keep_bf16 = self.mode == "fp8-first" and (name.endswith(".2") or name.startswith("model.0"))
        if self.mode == "bf16" or keep_bf16:
            x = x.to(torch.bfloat16).float()
        elif self.mode.startswith("fp8"):
            x = q_e4m3(x, x.abs().max().clamp_min(1e-12) / E4M3_MAX)
            ws = w.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-12) / E4M3_MAX
            w = q_e4m3(w, ws)
        y = F.conv2d(x, w, b, stride=s, padding=k // 2)
        return F.silu(y) if act else y


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    frames = [synth_image(rng, 540, 960), rng.integers(0, 256, (540, 960, 3), dtype=np.uint8)]
    xin = torch.from_numpy(np.concatenate([image_ops.preprocess_yolo_input(f)[0] for f in frames]))
    print("FP8 feasibility of the detector (CPU simulation, seeded synthetic weights; bar: boxes <= 1e-2 of the box size)\n")
    print("%-8s %-10s %14s %14s %16s %14s" % ("model", "mode", "head max err", "head mean err", "box err / size", "score err"))
    for scale in "nsm":
        kind, params, tensors = W.synth_yolov8_weights(scale, seed=0)
        ref = nets.YoloV8(params, tensors).head_flat(xin).numpy()
        for mode in ("bf16", "fp8", "fp8-first"):
            got = QuantNet(params, tensors, mode).head_flat(xin).numpy()
            err = np.abs(got - ref)
            box_err, score_err = 0.0, 0.0
            for b in range(len(frames)):
                gb, gs, _ = detect_post.decode(got[b])
                wb, ws, _ = detect_post.decode(ref[b])
                top = np.argsort(-ws)[:200]
                size = np.maximum(wb[top, 2] - wb[top, 0], wb[top, 3] - wb[top, 1])[:, None]
                box_err = max(box_err, float((np.abs(gb[top] - wb[top]) / size).max()))
                score_err = max(score_err, float(np.abs(gs[top] - ws[top]).max()))
            print("yolov8%-2s %-10s %14.4f %14.5f %16.4f %14.4f   %s" % (scale, mode, err.max(), err.mean(), box_err, score_err,
                                                                    "ok" if box_err < 1e-2 else "FAILS the 1e-2 bar"))


if __name__ == "__main__":
    main()
