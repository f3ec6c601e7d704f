#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Runs only in the build container, where the reference is mounted read-only at
/root/reference (it does not exist on the GPU box, so nothing at test time reads it).
Usage:  python tests/golden/make_golden.py            # writes tests/golden/*.npz

What is recorded
  tracker_<name>.npz   the reference's own ``DeepSORT`` (src/tracker/deepsort_tracker.py,
                       on top of src/tracker/core/*) driven with the seeded scenarios of
                       tests/scenarios.py; ReID features are planted by replacing
                       ``ReIDModel.extract_features_batched`` (the reference's own CPU
                       mock returns np.random.rand there, reid_model.py:104-107).
                       Per frame: the returned tuples and a snapshot of every live track.
  kalman.npz           initiate / predict / update / gating_distance of the reference
                       ``KalmanFilter`` on seeded inputs (N = 1 and N >= 2 gating).
  imageops.npz         ``preprocess_yolo_input`` / ``preprocess_reid_input`` /
                       ``scale_bboxes`` / ``letterbox`` of src/utils/image_processing.py on
                       seeded images (stored as checksums + small samples).
  known_answers.json   the known answers the reference's self-tests assert.

``tensorrt`` is absent here; a stub module with the handful of attributes touched at
import time (trt_engine.py:13,20-26) lets the reference facades import unmodified.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def import_reference():
    trt = types.ModuleType("tensorrt")

    class Logger:
        WARNING = 1

        def __init__(self, *a):
            pass
    trt.Logger = Logger
    for n in ("bool", "int8", "int32", "float16", "float32"):
        setattr(trt, n, n)
    sys.modules["tensorrt"] = trt
    sys.path.insert(0, REF)
    import src.tracker.deepsort_tracker as ds  # noqa
    import src.utils.image_processing as ip  # noqa
    import src.tracker.core.kalman_filter as kfm  # noqa
    import src.config as cfg  # noqa
    return ds, ip, kfm, cfg


def sha(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def scenario_digest(frames):
    h = hashlib.sha1()
    for f in frames:
        h.update(sha(f["boxes"], f["scores"], f["classes"], f["feats"]).encode())
    return h.hexdigest()


def dense_to_cov16(P):
    cov = np.zeros(16, np.float32)
    for i in range(4):
        cov[i], cov[4 + i], cov[8 + i], cov[12 + i] = P[i, i], P[i, i + 4], P[i + 4, i], P[i + 4, i + 4]
    Q = P.copy()
    for i in range(4):
        Q[i, i] = Q[i, i + 4] = Q[i + 4, i] = Q[i + 4, i + 4] = 0
    assert not Q.any(), "covariance has entries outside the four 2x2 blocks"
    return cov


def run_tracker_scenario(ds, cfg, frames, frame_hw, tracker_kw):
    import contextlib
    import io
    H, W = frame_hw
    frame = np.zeros((H, W, 3), np.uint8)
    base = frame.__array_interface__["data"][0]
    with contextlib.redirect_stdout(io.StringIO()):
        tracker = ds.DeepSORT(**tracker_kw)  # CPU mock mode: no engine, no CUDA here
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
            if not (f["scores"][i] >= tracker.min_detection_confidence and
                    cfg.CLASSES[cid] in cfg.CLASSES_TO_TRACK):
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
            trk_i.append([t.track_id, t.state, t.hits, t.age, t.time_since_update,
                          name_to_id[t.class_name], len(t.features)])
            trk_f.append(np.concatenate([t.mean.astype(np.float32), dense_to_cov16(t.covariance),
                                         [np.float32(t.confidence)]]))
            assert t.mean.dtype == np.float32 and t.covariance.dtype == np.float32
        trk_off.append(len(trk_i))
    return dict(
        out=np.asarray(outs, np.int64).reshape(-1, 6), out_conf=np.asarray(out_conf, np.float64),
        out_off=np.asarray(out_off, np.int64),
        trk_i=np.asarray(trk_i, np.int64).reshape(-1, 7),
        trk_f=np.asarray(trk_f, np.float32).reshape(-1, 25),
        trk_off=np.asarray(trk_off, np.int64),
        digest=np.frombuffer(scenario_digest(frames).encode(), np.uint8))


def make_tracker(ds, cfg):
    from scenarios import GOLDEN_SCENARIOS, GOLDEN_TRACKER_KW, make_scenario
    for name, kw in GOLDEN_SCENARIOS.items():
        frames = make_scenario(**kw)
        g = run_tracker_scenario(ds, cfg, frames, (1080, 1920), GOLDEN_TRACKER_KW.get(name, {}))
        np.savez_compressed(os.path.join(HERE, "tracker_%s.npz" % name), **g)
        print("tracker_%s: %d frames, %d outputs, %d track snapshots, max id %d" % (
            name, len(frames), len(g["out"]), len(g["trk_i"]),
            g["trk_i"][:, 0].max() if len(g["trk_i"]) else 0))


def make_kalman(kfm):
    rng = np.random.default_rng(2024)
    kf = kfm.KalmanFilter()
    rec = dict(z0=[], init=[], pred_in=[], pred_out=[], upd_in=[], upd_z=[], upd_out=[],
               gate_in=[], gate_z=[], gate_n=[], gate_out=[])
    for _ in range(60):
        z = np.array([rng.uniform(0, 1920), rng.uniform(0, 1080), rng.uniform(0.2, 1.5),
                      rng.uniform(8, 600)], np.float32)
        m, P = kf.initiate(z)
        rec["z0"].append(z)
        rec["init"].append(np.concatenate([m, dense_to_cov16(P)]))
        for step in range(12):
            rec["pred_in"].append(np.concatenate([m, dense_to_cov16(P)]))
            m, P = kf.predict(m, P)
            rec["pred_out"].append(np.concatenate([m, dense_to_cov16(P)]))
            n = 1 if step % 3 == 0 else int(rng.integers(2, 7))
            Z = (m[:4] + rng.normal(0, 1, (n, 4)) * np.array([6, 6, 0.03, 6])).astype(np.float32)
            g = kf.gating_distance(m, P, Z)
            assert g.dtype == np.float32
            Zp = np.zeros((6, 4), np.float32)
            Zp[:n] = Z
            gp = np.zeros(6, np.float32)
            gp[:n] = g
            rec["gate_in"].append(np.concatenate([m, dense_to_cov16(P)]))
            rec["gate_z"].append(Zp)
            rec["gate_n"].append(n)
            rec["gate_out"].append(gp)
            if rng.random() < 0.7:
                rec["upd_in"].append(np.concatenate([m, dense_to_cov16(P)]))
                rec["upd_z"].append(Z[0])
                m, P = kf.update(m, P, Z[0])
                rec["upd_out"].append(np.concatenate([m, dense_to_cov16(P)]))
    out = {k: np.asarray(v, np.int32 if k == "gate_n" else np.float32) for k, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "kalman.npz"), **out)
    print("kalman: %d initiate, %d predict, %d update, %d gating" % (
        len(out["init"]), len(out["pred_in"]), len(out["upd_in"]), len(out["gate_in"])))


def make_imageops(ip):
    from scenarios import IMAGEOPS_SEED, LETTERBOX_SIZES, REID_CROP_SIZES, synth_image
    rng = np.random.default_rng(IMAGEOPS_SEED)
    rec = {}
    sizes = LETTERBOX_SIZES
    for k, (h, w) in enumerate(sizes):
        img = synth_image(rng, h, w)
        t, ratios, pad = ip.preprocess_yolo_input(img, (640, 640))
        assert t.dtype == np.float32 and t.shape == (1, 3, 640, 640)
        rec["yolo%d_hw" % k] = np.asarray([h, w], np.int64)
        rec["yolo%d_seed_digest" % k] = np.frombuffer(sha(img).encode(), np.uint8)
        rec["yolo%d_meta" % k] = np.asarray([ratios[0], ratios[1], pad[0], pad[1]], np.float64)
        rec["yolo%d_digest" % k] = np.frombuffer(sha(t).encode(), np.uint8)
        # the exact uint8 letterboxed RGB image: the parity target for the fused kernel
        u8 = np.rint(t[0] * 255.0).astype(np.uint8)
        assert np.array_equal((u8.astype(np.float32) / 255.0), t[0])
        rec["yolo%d_rows" % k] = u8[:, ::37, ::41].copy()
    # ReID crops: (crop h, crop w) incl. upscaling, heavy downscaling, 1-pixel extents
    crops = REID_CROP_SIZES
    for k, (h, w) in enumerate(crops):
        img = synth_image(rng, max(h, 16), max(w, 16))[:h, :w]
        t = ip.preprocess_reid_input(np.ascontiguousarray(img), (128, 64))
        assert t.dtype == np.float32 and t.shape == (1, 3, 128, 64)
        rec["reid%d_hw" % k] = np.asarray([h, w], np.int64)
        rec["reid%d_seed_digest" % k] = np.frombuffer(sha(img).encode(), np.uint8)
        rec["reid%d_digest" % k] = np.frombuffer(sha(t).encode(), np.uint8)
        rec["reid%d_sample" % k] = t[0, :, ::13, ::7].copy()
    # scale_bboxes
    b = (rng.uniform(-20, 660, (64, 4))).astype(np.float32)
    for k, (h, w) in enumerate(sizes):
        r = min(640 / h, 640 / w, 1.0)
        nh, nw = int(round(h * r)), int(round(w * r))
        dw, dh = (640 - nw) / 2, (640 - nh) / 2
        rec["scale%d" % k] = ip.scale_bboxes(b, (h, w), (640, 640), (r, r), (dw, dh))
    rec["scale_in"] = b
    np.savez_compressed(os.path.join(HERE, "imageops.npz"), **rec)
    print("imageops: %d letterbox sizes, %d crop sizes" % (len(sizes), len(crops)))


def make_known_answers():
    """Known answers asserted or printed by the reference's own self-tests."""
    ka = {
        "xyah": [  # src/tracker/core/detection.py:53-123
            {"tlwh": [10, 20, 30, 60], "xyah": [25, 50, 0.5, 60]},
            {"tlwh": [10, 20, 30, 0], "xyah": [25, 20, 0, 0]},
        ],
        "iou": [  # src/tracker/core/matching.py:220-333
            {"a": [0, 0, 10, 10], "b": [0, 0, 10, 10], "iou": 1.0},
            {"a": [0, 0, 10, 10], "b": [5, 5, 10, 10], "iou": 25.0 / 175.0},
            {"a": [0, 0, 10, 10], "b": [0, 0, 5, 5], "iou": 0.25},
            {"a": [0, 0, 10, 10], "b": [20, 20, 5, 5], "iou": 0.0},
        ],
        "chi2inv95_4": 9.487729036781154,  # src/tracker/core/kalman_filter.py:16
    }
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)


if __name__ == "__main__":
    ds, ip, kfm, cfg = import_reference()
    make_known_answers()
    make_kalman(kfm)
    make_imageops(ip)
    make_tracker(ds, cfg)
