"""Overlay (N3) on the GPU: the kernel against the numpy oracle (bit-exact) and the module's reference-shaped functions
against the reference's own drawing (src/utils/visualization.py through oracle/_ref)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
from ai_camera_b200 import config, visualization as V  # noqa: E402
from oracle import overlay as O  # noqa: E402
import test_overlay as T  # noqa: E402


def test_kernel_matches_oracle_on_random_items():
    rng = np.random.default_rng(5)
    S, H, W = 3, 270, 480
    ov = V.Overlay(device="cuda:0", slots=32)
    frames = rng.integers(0, 256, (S, H, W, 3), dtype=np.uint8)
    per_frame = []
    for s in range(S):
        items = []
        for (x1, y1, x2, y2) in T.random_rects(rng, 25, H, W):
            items.append((int(rng.integers(0, 2)), x1, y1, x2, y2, int(rng.integers(0, 1 << 24)), 0))
        items += ov.track_items(T.tracks_case(rng, H, W, 8), (H, W))
        order = rng.permutation(len(items))
        per_frame.append([items[i] for i in order])  # any order: later items overwrite earlier ones
    dev = torch.from_numpy(frames.copy()).cuda()
    ov.draw(dev, per_frame)
    torch.cuda.synchronize()
    atlas = ov.atlas.cpu().numpy()
    for s in range(S):
        want = O.draw_items(frames[s].copy(), per_frame[s], atlas)
        assert np.array_equal(dev[s].cpu().numpy(), want), "frame %d" % s


def test_reference_shaped_functions_match_reference_drawing():
    rv = T.reference_visualization()
    rng = np.random.default_rng(21)
    H, W = 1080, 1920
    frame = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    objs = T.tracks_case(rng, H, W, 20)
    lines = ["AICamera: YOLOv8 + DeepSORT", "Input: aicamera_test_clip.mp4", "FPS: 1234.56"]
    ref = rv.draw_info_panel(rv.draw_tracks(frame.copy(), objs), lines)
    got = frame.copy()
    out = V.draw_info_panel(V.draw_tracks(got, objs), lines)
    assert out is got  # drawn in place, like the reference
    mx, frac = T.check_close(got, ref, max_diff=6)
    print("draw_tracks + draw_info_panel on a 1080p frame vs the reference: max difference %d, %.4f %% of the pixels differ" % (mx, frac))
    boxes = np.array([[30.7, 40.2, 200.9, 300.1], [300, 20, 420, 200], [1500, 100, 1915, 1075]], np.float32)
    scores = np.array([0.91, 0.456, 0.3], np.float32)
    cls = np.array([0, 2, 99])
    ref = rv.draw_fps(rv.draw_detections(frame.copy(), boxes, scores, cls, config.CLASSES), 59.94)
    dev = torch.from_numpy(frame.copy()).cuda()  # device tensors are drawn in place without any copy
    V.draw_fps(V.draw_detections(dev, boxes, scores, cls, config.CLASSES), 59.94)
    T.check_close(dev.cpu().numpy(), ref, max_diff=6)


def test_batched_overlay_from_track_tables():
    rng = np.random.default_rng(8)
    S, H, W = 4, 360, 640
    ov = V.Overlay(device="cuda:0", slots=64)
    frames = rng.integers(0, 256, (S, H, W, 3), dtype=np.uint8)
    tracks = [T.tracks_case(rng, H, W, 6) for _ in range(S)]
    dev = torch.from_numpy(frames.copy()).cuda()
    ov.draw(dev, [ov.track_items(t, (H, W)) for t in tracks],
            panels=[(lambda img, s=s: V.Overlay._draw_fps(img, 10.0 * s)) for s in range(S)])
    rv = T.reference_visualization()
    for s in range(S):
        ref = rv.draw_fps(rv.draw_tracks(frames[s].copy(), tracks[s]), 10.0 * s)
        T.check_close(dev[s].cpu().numpy(), ref, max_diff=6)
