"""Reference-shaped facades (YOLODetector.detect, DeepSORT.update, ReIDModel, TRTEngine) and the
batched pipeline, end to end on synthetic video, against the CPU oracle pipeline (-m gpu)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    from ai_camera_b200 import synth
    from ai_camera_b200.pipeline import BatchDetector
    d = tmp_path_factory.mktemp("blobs")
    yolo, reid = synth.make_blobs(str(d))
    video = synth.SynthVideo(3, (540, 960), n_frames=10, seed=99)
    det = BatchDetector(yolo, 3)
    delta, bias = synth.calibrate_detector(det, video.frames(0), target_tracked=8.0)
    del det
    torch.cuda.empty_cache()
    return dict(yolo=yolo, reid=reid, video=video, bias=bias, delta=delta)


def _iou(a, b):
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    u = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - iw * ih
    return iw * ih / u if u > 0 else 0.0


def test_detect_matches_oracle(setup):
    from ai_camera_b200.yolo_detector import YOLODetector
    from oracle.pipeline import Detector
    det = YOLODetector(setup["yolo"])
    for n, b in setup["bias"].items():
        det.trt_engine.set_bias(n, b)
    ora = Detector(setup["yolo"], bias_overrides=setup["bias"])
    total = matched = 0
    for t in range(3):
        frame = setup["video"].ring[t, 0].cpu().numpy()
        gb, gs, gc, gi = det.detect(frame)
        ob, os_, oc, oi = ora.detect(frame)
        assert gb.dtype == np.float32 and gs.dtype == np.float32 and gc.dtype == np.int32 and gi.dtype == np.int64
        assert len(gb) > 0 and (np.diff(gs) <= 0).all()
        for k in range(len(ob)):
            if os_[k] < 0.32:
                continue
            total += 1
            cand = [j for j in range(len(gb)) if gc[j] == oc[k] and _iou(gb[j], ob[k]) > 0.9]
            if cand:
                j = max(cand, key=lambda j: _iou(gb[j], ob[k]))
                size = max(ob[k][2] - ob[k][0], ob[k][3] - ob[k][1])
                assert np.abs(gb[j] - ob[k]).max() / size < 1e-2
                assert abs(gs[j] - os_[k]) < 2e-2
                matched += 1
    print("detections matched: %d / %d" % (matched, total))
    assert total > 0 and matched >= 0.9 * total
    # empty result convention (yolo_detector.py:126)
    for n, b in setup["bias"].items():
        det.trt_engine.set_bias(n, b - 30.0)
    e = det.detect(setup["video"].ring[0, 0].cpu().numpy())
    assert e[0].shape == (0, 4) and e[0].dtype == np.float64 and e[3].dtype == np.int64 and len(e[1]) == 0


def test_reid_model_and_trt_engine_match_oracle(setup):
    from ai_camera_b200.reid_model import ReIDModel
    from ai_camera_b200.trt_engine import TRTEngine
    from oracle import image_ops, nets
    frame = setup["video"].ring[0, 1].cpu().numpy()
    rects = [(10, 20, 90, 200), (300, 100, 364, 228), (500, 30, 530, 60), (700, 200, 959, 539)]
    crops = [frame[y1:y2, x1:x2] for (x1, y1, x2, y2) in rects]
    m = ReIDModel(setup["reid"], max_batch=3)
    got = m.extract_features_batched(crops + [np.zeros((0, 5, 3), np.uint8)])
    assert got.shape == (4, 512) and got.dtype == np.float32
    want = nets.load_net(setup["reid"]).forward(torch.from_numpy(image_ops.reid_batch(frame, rects))).numpy()
    assert ((got * want).sum(1) >= 0.999).all()
    assert m.extract_features_batched([]).shape == (0, 512)
    # the TRTEngine-shaped seam: reference tensors in, reference tensors out
    eng = TRTEngine(setup["reid"], max_batch=4)
    assert eng.get_input_details()[0].name == "input" and eng.get_output_details()[0].shape[-1] == 512
    x = torch.from_numpy(image_ops.reid_batch(frame, rects)).cuda()
    out = eng({"input": x})["output"]
    torch.cuda.synchronize()
    assert ((out.cpu().numpy() * want).sum(1) >= 0.999).all()
    with pytest.raises(TypeError):
        eng(x)
    with pytest.raises(FileNotFoundError):
        TRTEngine(setup["reid"] + ".missing")
    yeng = TRTEngine(setup["yolo"])
    assert [i.name for i in yeng.get_output_details()] == ["num_dets", "bboxes", "scores", "labels"]
    xin, _, _ = image_ops.preprocess_yolo_input(frame)
    o = yeng.infer({"images": torch.from_numpy(xin).cuda()})
    assert o["num_dets"].shape == (1, 1) and o["bboxes"].shape == (1, 100, 4) and o["labels"].dtype == torch.int32


def test_deepsort_update_matches_oracle_given_same_detections(setup):
    """Same detections and frames into the device DeepSORT (bf16 ReID net) and the oracle DeepSORT
    (fp32 ReID net): same ids, classes and boxes in every returned tuple."""
    from ai_camera_b200.deepsort_tracker import DeepSORT
    from oracle.pipeline import Detector, ReID
    from oracle.tracker import DeepSORT as OracleDeepSORT
    ora_det = Detector(setup["yolo"], bias_overrides=setup["bias"])
    gpu = DeepSORT(setup["reid"], n_init=2)
    ora = OracleDeepSORT(reid_fn=ReID(setup["reid"]), n_init=2)
    seen = 0
    for t in range(8):
        frame = setup["video"].ring[t, 2].cpu().numpy()
        b, s, c, _ = ora_det.detect(frame)
        got = gpu.update(b, s, c, frame.copy())
        want = ora.update(b, s, c, frame)
        assert [g[:6] for g in got] == [w[:6] for w in want], "frame %d" % t
        assert np.allclose([g[6] for g in got], [w[6] for w in want])
        seen += len(want)
    assert seen > 0
    assert gpu.update(np.array([]), np.array([]), np.array([]), frame) == ora.update(
        np.array([]), np.array([]), np.array([]), frame)


def test_pipeline_equals_single_stream_facades(setup):
    """The batched device-resident pipeline gives, per stream, exactly what the single-stream
    facades give (streams are independent; ids start at 1 in each)."""
    from ai_camera_b200.deepsort_tracker import DeepSORT
    from ai_camera_b200.pipeline import TrackingPipeline
    from ai_camera_b200.yolo_detector import YOLODetector
    from ai_camera_b200 import config
    video = setup["video"]
    pipe = TrackingPipeline(setup["yolo"], setup["reid"], 3, max_tracks=128, n_init=2)
    det = YOLODetector(setup["yolo"])
    for n, b in setup["bias"].items():
        pipe.detector.engine.set_bias(n, b)
        det.trt_engine.set_bias(n, b)
    trks = [DeepSORT(setup["reid"], n_init=2, max_dets=100, max_tracks=128) for _ in range(3)]
    for t in range(6):
        ot, oc, on = pipe.step(video.frames(t))
        torch.cuda.synchronize()
        ot, oc, on = ot.cpu().numpy(), oc.cpu().numpy(), on.cpu().numpy()
        for s in range(3):
            frame = video.ring[t, s].cpu().numpy()
            b, sc, c, _ = det.detect(frame)
            want = trks[s].update(b, sc, c, frame)
            got = [tuple(int(v) for v in ot[s, k, :5]) + (config.CLASSES[ot[s, k, 5]], float(oc[s, k])) for k in range(on[s])]
            assert got == want, (t, s)
    assert not pipe.tracker.overflow().any()
