"""BASELINE.json configs[0] (C1): the reference's own clip, one stream, batch 1, through the reference-shaped
facades (YOLODetector.detect -> DeepSORT.update, the loop of /root/reference/src/aicamera_tracker.py:169-207),
against the CPU oracle pipeline on the same decoded frames; plus the single-upload path of the frame loop."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLIP = os.path.join(GOLDEN, "aicamera_test_clip.mp4")


def _clip_frames(n):
    from ai_camera_b200.aicamera_tracker import video_frames
    return list(video_frames(CLIP, n))


@pytest.fixture(scope="module")
def blobs(tmp_path_factory):
    from ai_camera_b200 import synth
    return synth.make_blobs(str(tmp_path_factory.mktemp("blobs")))


def test_clip_decodes_as_recorded():
    import hashlib
    meta = json.load(open(os.path.join(GOLDEN, "clip_meta.json")))
    frames = _clip_frames(8)
    assert list(frames[0].shape) == meta["frame_shape"]
    h = hashlib.sha256()
    for f in frames:
        h.update(f.tobytes())
    assert h.hexdigest() == meta["decoded_sha256_first8"], "cv2 decodes the clip differently from the box the fixture was made on"


def test_clip_single_stream_matches_oracle(blobs):
    from ai_camera_b200 import synth
    from ai_camera_b200.aicamera_tracker import run_single_stream
    from ai_camera_b200.deepsort_tracker import DeepSORT
    from ai_camera_b200.yolo_detector import YOLODetector
    from e2e_compare import compare_runs, oracle_margins
    from oracle.pipeline import Pipeline
    yolo, reid = blobs
    bias = synth.shifted_class_bias(yolo, synth.CLIP_LOGIT_SHIFT)
    frames = _clip_frames(64)
    det = YOLODetector(yolo)
    synth.apply_class_bias(det.trt_engine, bias)
    trk = DeepSORT(reid)
    got, dets_got = [], []
    stats = run_single_stream(frames, det, trk, on_frame=lambda i, f, d, tr: (got.append(tr), dets_got.append(len(d[0]))))
    ora = Pipeline(yolo, reid, yolo_bias=bias)
    ora.tracker.tracker_core.cost_log = []
    want, scores = [], []
    for f in frames:
        b, s, c, _ = ora.detector.detect(f)
        scores.append(s)
        want.append(ora.tracker.update(b, s, c, f))
    r = compare_runs(got, want)
    m = oracle_margins(ora.tracker.tracker_core.cost_log, scores)
    print("clip C1: %s" % r)
    print("clip C1: detections/frame device %.1f oracle %.1f; margins %s; loop %s" % (
        np.mean(dets_got), np.mean([len(s) for s in scores]), m, stats.summary()))
    # Two INDEPENDENT runs (bf16 tensor-core nets vs fp32 nets): the seeded random-weight detector has detections
    # within 1e-4 of the 0.3 score threshold in most frames (printed above), so a detection can exist on one side only
    # and shift the ids after it; what holds is that the reported tracks coincide and mostly keep one relabelling.
    # The strict statement (same detections in -> identical tuples out) is the next test.
    assert r["oracle_tracks"] > 100, "the clip run reported too few tracks to compare"
    assert r["matched"] >= 0.85 * r["oracle_tracks"]
    assert r["consistent"] >= 0.7 * r["matched"]


def test_clip_tracker_matches_oracle_given_oracle_detections(blobs):
    """C1 on the real clip frames, strict: the oracle detector's detections of every frame go into the device
    DeepSORT (bf16 ReID net on the real frame) and into the oracle DeepSORT (fp32 ReID net): ids, classes and
    integer boxes of every returned tuple are equal over the first 64 frames; the threshold margins the sequence
    came within are logged."""
    from ai_camera_b200 import synth
    from ai_camera_b200.deepsort_tracker import DeepSORT
    from e2e_compare import oracle_margins
    from oracle.pipeline import Detector, ReID
    from oracle.tracker import DeepSORT as OracleDeepSORT
    yolo, reid = blobs
    bias = synth.shifted_class_bias(yolo, synth.CLIP_LOGIT_SHIFT)
    frames = _clip_frames(64)
    ora_det = Detector(yolo, bias_overrides=bias)
    gpu = DeepSORT(reid)
    ora = OracleDeepSORT(reid_fn=ReID(reid))
    ora.tracker_core.cost_log = []
    scores, n_out, gate = [], 0, 1e9
    for t, frame in enumerate(frames):
        b, s, c, _ = ora_det.detect(frame)
        scores.append(s)
        _, app, d2 = ora.probe_costs(b, s, c, frame)
        if d2.size:
            gate = min(gate, float(np.abs(d2 - 9.487729036781154).min()))
        got = gpu.update(b, s, c, frame.copy())
        want = ora.update(b, s, c, frame)
        assert [g[:6] for g in got] == [w[:6] for w in want], "frame %d" % t
        n_out += len(want)
    m = oracle_margins(ora.tracker_core.cost_log, scores)
    m["gate"] = gate
    print("clip C1 (oracle detections): %d frames, %d track tuples identical; minimum margins %s" % (len(frames), n_out, m))
    assert n_out > 300


def test_frame_loop_single_upload_equals_double_upload(blobs):
    """detect() leaves the frame in HBM; update() given that device tensor returns exactly what update() given
    the numpy frame returns (one H2D copy per frame instead of two)."""
    from ai_camera_b200 import synth
    from ai_camera_b200.aicamera_tracker import run_single_stream
    from ai_camera_b200.deepsort_tracker import DeepSORT
    from ai_camera_b200.yolo_detector import YOLODetector
    yolo, reid = blobs
    bias = synth.shifted_class_bias(yolo, synth.CLIP_LOGIT_SHIFT)
    frames = _clip_frames(12)
    outs = []
    for share in (True, False):
        det = YOLODetector(yolo)
        synth.apply_class_bias(det.trt_engine, bias)
        trk = DeepSORT(reid, n_init=2)
        res = []
        run_single_stream(frames, det, trk, on_frame=lambda i, f, d, tr: res.append(tr), share_upload=share)
        outs.append(res)
    assert outs[0] == outs[1] and sum(len(r) for r in outs[0]) > 0


def test_frame_loop_writes_device_drawn_video(blobs, tmp_path):
    """The loop's save step (src/aicamera_tracker.py:211-236): the frames the AnnotatedWriter hands cv2.VideoWriter are
    the reference's draw_tracks + draw_info_panel of the same frame and tracks (drawn on the device copy of the frame the
    detector uploaded), and the file it writes decodes to as many frames."""
    import cv2
    from ai_camera_b200 import synth
    from ai_camera_b200.aicamera_tracker import AnnotatedWriter, run_single_stream
    from ai_camera_b200.deepsort_tracker import DeepSORT
    from ai_camera_b200.yolo_detector import YOLODetector
    import test_overlay as T
    rv = T.reference_visualization()
    yolo, reid = blobs
    frames = _clip_frames(10)
    det = YOLODetector(yolo)
    synth.apply_class_bias(det.trt_engine, synth.shifted_class_bias(yolo, synth.CLIP_LOGIT_SHIFT))
    trk = DeepSORT(reid, n_init=2)
    out = AnnotatedWriter(tmp_path / "clip_tracked.mp4", det.device)
    written, drawn = [], 0
    real_write = None

    def on_frame(idx, frame, dets, tracks):
        nonlocal real_write, drawn
        lines = ["AICamera: YOLOv8 + DeepSORT", "Input: clip", "FPS: %.2f" % (10.0 + idx)]
        out.write(det.device_frame, tracks, lines)
        if real_write is None:  # (after the first call the cv2 writer exists: tap what it is given)
            real_write = out.writer.write
            out.writer = type("Tap", (), {"write": lambda self, img: (written.append(img.copy()), real_write(img))[1],
                                          "release": out.writer.release})()
        else:
            ref = rv.draw_info_panel(rv.draw_tracks(frame.copy(), tracks), lines)
            T.check_close(written[-1], ref, max_diff=6)
            drawn += len(tracks)
    run_single_stream(frames, det, trk, on_frame=on_frame)
    out.close()
    assert drawn > 0, "no track was drawn on the clip frames"
    cap = cv2.VideoCapture(str(tmp_path / "clip_tracked.mp4"))
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == len(frames)
