"""Run the REAL reference tracker (snapshot under ``oracle/_ref/``, see ``oracle/build_ref.py``) on the
scenarios / frames the tests and the CPU baseline use (oracle; test infrastructure).

The reference's ``DeepSORT`` (``src/tracker/deepsort_tracker.py``) is constructed in its own "CPU mock"
ReID mode (``reid_model.py:51-56``; ``torch.cuda.is_available`` is masked during construction because
``DeepSORT.__init__`` passes no device and would otherwise look for an engine file on the GPU box,
SURVEY.md 8c) and its ``reid_model.extract_features_batched`` is replaced by the caller's feature
function - the same seam the golden generator uses.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_mods = None


def available():
    return os.path.isfile(os.path.join(REF_DIR, "src", "tracker", "deepsort_tracker.py"))


def import_reference(ref_root=None):
    """(deepsort_tracker, image_processing, kalman_filter, config) modules of the reference."""
    global _mods
    if _mods is not None:
        return _mods
    root = ref_root or REF_DIR
    if ref_root is None and not available():
        raise RuntimeError("oracle/_ref is not built: run `python oracle/build_ref.py` where /root/reference exists")
    if ref_root is not None or "tensorrt" not in sys.modules:
        stub = os.path.join(REF_DIR, "tensorrt.py")
        if os.path.isfile(stub) and REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)  # provides both `tensorrt` (stub) and, for the snapshot, `src`
    if "tensorrt" not in sys.modules:
        import types
        try:
            import tensorrt  # noqa: F401  (the stub, or a real one)
        except ImportError:
            trt = types.ModuleType("tensorrt")
            trt.Logger = type("Logger", (), {"WARNING": 1, "__init__": lambda self, *a: None})
            for n in ("bool", "int8", "int32", "float16", "float32"):
                setattr(trt, n, n)
            sys.modules["tensorrt"] = trt
    if root not in sys.path:
        sys.path.insert(0, root)
    with contextlib.redirect_stdout(io.StringIO()):  # (src/config.py prints warnings about missing engine files)
        import src.config as cfg
        import src.tracker.core.kalman_filter as kfm
        import src.tracker.deepsort_tracker as ds
        import src.utils.image_processing as ip
    _mods = (ds, ip, kfm, cfg)
    return _mods


def make_deepsort(**tracker_kw):
    """The reference's DeepSORT, constructed without an engine (CPU mock ReID)."""
    import torch
    ds = import_reference()[0]
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            trk = ds.DeepSORT(**tracker_kw)
    finally:
        torch.cuda.is_available = real
    return trk


def dense_to_cov16(P):
    cov = np.zeros(16, np.float32)
    for i in range(4):
        cov[i], cov[4 + i], cov[8 + i], cov[12 + i] = P[i, i], P[i, i + 4], P[i + 4, i], P[i + 4, i + 4]
    Q = P.copy()
    for i in range(4):
        Q[i, i] = Q[i, i + 4] = Q[i + 4, i] = Q[i + 4, i + 4] = 0
    assert not Q.any(), "covariance has entries outside the four 2x2 blocks"
    return cov


def run_tracker_scenario(frames, frame_hw=(1080, 1920), tracker_kw=None):
    """A planted-feature scenario (tests/scenarios.py) through the reference DeepSORT -> dict in the golden layout
    (tests/golden_util.py).  Features are planted by identifying each crop view by its address inside the frame."""
    ds, _, _, cfg = import_reference()
    H, W = frame_hw
    frame = np.zeros((H, W, 3), np.uint8)
    base = frame.__array_interface__["data"][0]
    tracker = make_deepsort(**(tracker_kw or {}))
    state = {}

    def planted(crops):
        out = []
        for c in crops:
            off = c.__array_interface__["data"][0] - base
            y1, x1 = divmod(off // 3, W)
            key = (x1, y1, x1 + c.shape[1], y1 + c.shape[0])
            idx = state["by_rect"][key].pop(0)
            out.append(state["feats"][idx])
        return np.asarray(out, np.float32).reshape(len(out), -1)

    tracker.reid_model.extract_features_batched = planted
    name_to_id = {n: i for i, n in enumerate(cfg.CLASSES)}
    outs, out_conf, out_off = [], [], [0]
    trk_i, trk_f, trk_off = [], [], [0]
    for f in frames:
        by_rect = {}
        for i, b in enumerate(f["boxes"]):
            cid = int(f["classes"][i])
            if not (f["scores"][i] >= tracker.min_detection_confidence and cfg.CLASSES[cid] in cfg.CLASSES_TO_TRACK):
                continue
            x1, y1, x2, y2 = map(int, b)
            key = (max(0, x1), max(0, y1), min(W, x2), min(H, y2))
            by_rect.setdefault(key, []).append(i)
        state["by_rect"], state["feats"] = by_rect, f["feats"]
        res = tracker.update(f["boxes"], f["scores"], f["classes"], frame)
        for (x1, y1, x2, y2, tid, cname, conf) in res:
            outs.append([x1, y1, x2, y2, tid, name_to_id[cname]])
            out_conf.append(conf)
        out_off.append(len(outs))
        for t in tracker.tracker_core.tracks:
            trk_i.append([t.track_id, t.state, t.hits, t.age, t.time_since_update, name_to_id[t.class_name], len(t.features)])
            trk_f.append(np.concatenate([t.mean.astype(np.float32), dense_to_cov16(t.covariance), [np.float32(t.confidence)]]))
        trk_off.append(len(trk_i))
    return dict(out=np.asarray(outs, np.int64).reshape(-1, 6), out_conf=np.asarray(out_conf, np.float64),
                out_off=np.asarray(out_off, np.int64), trk_i=np.asarray(trk_i, np.int64).reshape(-1, 7),
                trk_f=np.asarray(trk_f, np.float32).reshape(-1, 25), trk_off=np.asarray(trk_off, np.int64))


class RefDeepSORT:
    """The reference DeepSORT with features from ``reid_fn(frame_bgr, rects)`` (the oracle ReID net): same call
    signature as ``oracle.tracker.DeepSORT.update`` with a frame.  Used by the CPU baseline of bench.py."""

    def __init__(self, reid_fn, **tracker_kw):
        self.trk = make_deepsort(**tracker_kw)
        self.reid_fn = reid_fn
        self._frame = None
        self.trk.reid_model.extract_features_batched = self._features

    def _features(self, crops):
        frame = self._frame
        base = frame.__array_interface__["data"][0]
        W = frame.shape[1]
        rects = []
        for c in crops:
            off = c.__array_interface__["data"][0] - base
            y1, x1 = divmod(off // 3, W)
            rects.append((x1, y1, x1 + c.shape[1], y1 + c.shape[0]))
        return self.reid_fn(frame, rects)

    def update(self, boxes, scores, class_ids, frame_bgr):
        self._frame = np.ascontiguousarray(frame_bgr)
        return self.trk.update(boxes, scores, class_ids, self._frame)
