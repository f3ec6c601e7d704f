"""DeepSORT association core + ``DeepSORT.update`` facade, restated (oracle; test
infrastructure).  Every function cites the reference lines it follows
(paths relative to ``/root/reference``).

Differences from the reference, all deliberate and documented in DESIGN.md:
  * Kalman arithmetic goes through ``oracle.kalman`` (explicit float32 ops instead
    of BLAS/LAPACK calls; bit-identical to the reference, host-CPU independent).
  * The track-id counter belongs to the tracker instance (one counter per video
    stream, starting at 1) instead of being a class-level global
    (``src/tracker/core/track.py:21,42-43``; reset in ``tracker_core.py:42``).
  * Class names are carried as COCO class ids; names are looked up on output.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy.optimize import linear_sum_assignment

from . import kalman
from .constants import (CLASSES, CLASSES_TO_TRACK, DEEPSORT_MAX_AGE, DEEPSORT_MAX_DIST,
                        DEEPSORT_MAX_IOU_DISTANCE, DEEPSORT_MIN_CONFIDENCE, DEEPSORT_N_INIT,
                        DEEPSORT_NN_BUDGET, INFTY_COST)

F32 = np.float32
TENTATIVE, CONFIRMED, DELETED = 1, 2, 3  # src/tracker/core/track.py:10-14


class Det:
    """src/tracker/core/detection.py:15-47."""
    __slots__ = ("tlwh", "confidence", "class_id", "feature")

    def __init__(self, tlwh, confidence, class_id, feature):
        self.tlwh = np.asarray(tlwh, dtype=F32)
        self.confidence = float(confidence)
        self.class_id = int(class_id)
        self.feature = None if feature is None else np.asarray(feature, dtype=F32)

    def to_xyah(self):
        ret = self.tlwh.copy()
        ret[:2] += ret[2:] / 2.0
        if ret[3] > 0:
            ret[2] /= ret[3]
        else:
            ret[2] = 0
        return ret


class Trk:
    """src/tracker/core/track.py:23-151 (state kept as mean[8] + 16 covariance floats)."""
    __slots__ = ("track_id", "mean", "cov", "class_id", "confidence", "hits", "age",
                 "time_since_update", "state", "features", "n_init", "max_age", "budget")

    def __init__(self, track_id, mean, cov, det: Det, n_init, max_age, budget):
        self.track_id = track_id
        self.mean, self.cov = mean, cov
        self.class_id = det.class_id
        self.confidence = det.confidence
        self.hits, self.age, self.time_since_update = 1, 1, 0
        self.state = TENTATIVE
        self.n_init, self.max_age, self.budget = n_init, max_age, budget
        self.features: List[np.ndarray] = []
        if det.feature is not None:
            self._add_feature(det.feature)

    def _add_feature(self, f):  # track.py:70-74
        self.features.append(f)
        if self.budget is not None and len(self.features) > self.budget:
            self.features.pop(0)

    def predict(self):  # track.py:76-80
        m, c = kalman.predict(self.mean, self.cov)
        self.mean, self.cov = m[0], c[0]
        self.age += 1
        self.time_since_update += 1

    def update(self, det: Det):  # track.py:82-104
        self.mean, self.cov = kalman.update(self.mean, self.cov, det.to_xyah())
        if det.feature is not None:
            self._add_feature(det.feature)
        self.hits += 1
        self.time_since_update = 0
        self.confidence = det.confidence
        self.class_id = det.class_id
        if self.state == TENTATIVE and self.hits >= self.n_init:
            self.state = CONFIRMED

    def mark_missed(self):  # track.py:106-119
        if self.state == TENTATIVE:
            self.state = DELETED
        elif self.state == CONFIRMED and self.time_since_update > self.max_age:
            self.state = DELETED

    def to_tlwh(self):  # track.py:133-151
        p = self.mean[:4].copy()
        if p[3] > 0:
            width = p[2] * p[3]
        else:
            width = 0
            p[3] = max(0, p[3])
        tl_x = p[0] - width / 2.0
        tl_y = p[1] - p[3] / 2.0
        return np.array([tl_x, tl_y, width, p[3]], dtype=F32)


def iou(bbox_tlwh, cand_tlwh):
    """src/tracker/core/matching.py:13-54 (float32 elementwise)."""
    if cand_tlwh.size == 0:
        return np.array([], dtype=F32)
    box_tl, box_br = bbox_tlwh[:2], bbox_tlwh[:2] + bbox_tlwh[2:]
    c_tl = cand_tlwh[:, :2]
    c_br = cand_tlwh[:, :2] + cand_tlwh[:, 2:]
    tlx = np.maximum(box_tl[0], c_tl[:, 0])
    tly = np.maximum(box_tl[1], c_tl[:, 1])
    brx = np.minimum(box_br[0], c_br[:, 0])
    bry = np.minimum(box_br[1], c_br[:, 1])
    iw = np.maximum(0., brx - tlx)
    ih = np.maximum(0., bry - tly)
    inter = iw * ih
    area_box = bbox_tlwh[2] * bbox_tlwh[3]
    area_c = cand_tlwh[:, 2] * cand_tlwh[:, 3]
    union = area_box + area_c - inter
    return inter / np.maximum(union, 1e-7)


def iou_cost(tracks, dets, t_idx, d_idx):
    """matching.py:57-106."""
    if len(t_idx) == 0 or len(d_idx) == 0:
        return np.empty((len(t_idx), len(d_idx)), dtype=F32)
    cm = np.full((len(t_idx), len(d_idx)), INFTY_COST, dtype=F32)
    cand = np.asarray([dets[d].tlwh for d in d_idx], dtype=F32)
    for r, t in enumerate(t_idx):
        cm[r, :] = 1.0 - iou(tracks[t].to_tlwh(), cand)
    return cm


def cosine_distance(a, b):
    """matching.py:109-141 (float32; the matrix product goes through BLAS in both
    the reference and here, so this value is NOT bit-reproducible across hosts and
    device parity on it is a tolerance)."""
    if a.size == 0 or b.size == 0:
        return np.empty((a.shape[0], b.shape[0]), dtype=F32)
    na = np.linalg.norm(a, axis=1, keepdims=True)
    nb = np.linalg.norm(b, axis=1, keepdims=True)
    an = a / np.maximum(na, 1e-7)
    bn = b / np.maximum(nb, 1e-7)
    return np.maximum(1.0 - np.dot(an, bn.T), 0.0)


def appearance_cost(tracks, dets, t_idx, d_idx):
    """matching.py:144-217: min over the track gallery of the cosine distance;
    INFTY_COST where the detection or the track has no feature."""
    if len(t_idx) == 0 or len(d_idx) == 0:
        return np.empty((len(t_idx), len(d_idx)), dtype=F32)
    cm = np.full((len(t_idx), len(d_idx)), INFTY_COST, dtype=F32)
    cols = [k for k, d in enumerate(d_idx) if dets[d].feature is not None]
    if not cols:
        return cm
    feats = np.asarray([dets[d_idx[k]].feature for k in cols], dtype=F32)
    for r, t in enumerate(t_idx):
        if not tracks[t].features:
            continue
        gal = np.asarray(tracks[t].features, dtype=F32)
        cm[r, cols] = np.min(cosine_distance(gal, feats), axis=0)
    return cm


def gate_by_mahalanobis(cm, tracks, dets, t_idx, d_idx):
    """linear_assignment.py:160-212 (4 dof, strict >, in place)."""
    meas = np.asarray([dets[d].to_xyah() for d in d_idx])
    for r, t in enumerate(t_idx):
        g = kalman.gating_distance(tracks[t].mean, tracks[t].cov, meas)
        cm[r, g > kalman.CHI2_GATE] = INFTY_COST
    return cm


def min_cost_matching(metric, max_distance, tracks, dets, t_idx, d_idx, log=None):
    """linear_assignment.py:19-88."""
    if not d_idx or not t_idx:
        return [], t_idx, d_idx
    cm = metric(tracks, dets, t_idx, d_idx)
    if log is not None:
        log.append(("raw", list(t_idx), list(d_idx), cm.copy(), max_distance))
    cm[cm > max_distance] = max_distance + 1e-5
    rows, cols = linear_sum_assignment(cm)
    matches, un_t, un_d = [], list(t_idx), list(d_idx)
    for r, c in zip(rows, cols):
        if cm[r, c] <= max_distance:
            matches.append((t_idx[r], d_idx[c]))
            un_t.remove(t_idx[r])
            un_d.remove(d_idx[c])
    return matches, un_t, un_d


def matching_cascade(metric, max_distance, depth, tracks, dets, t_idx, d_idx, log=None):
    """linear_assignment.py:91-157."""
    un_d = list(d_idx)
    matches = []
    for level in range(depth):
        if not un_d:
            break
        lvl = [t for t in t_idx if tracks[t].time_since_update == level + 1]
        if not lvl:
            continue
        m, _, un_d = min_cost_matching(metric, max_distance, tracks, dets, lvl, un_d, log)
        matches.extend(m)
    matched = {t for t, _ in matches}
    return matches, [t for t in t_idx if t not in matched], un_d


class TrackerCore:
    """src/tracker/core/tracker_core.py:12-199."""

    def __init__(self, max_cosine_distance=DEEPSORT_MAX_DIST, nn_budget=DEEPSORT_NN_BUDGET,
                 max_iou_distance=DEEPSORT_MAX_IOU_DISTANCE, max_age=DEEPSORT_MAX_AGE,
                 n_init=DEEPSORT_N_INIT):
        self.max_cosine_distance = max_cosine_distance
        self.nn_budget = nn_budget
        self.max_iou_distance = max_iou_distance
        self.max_age = max_age
        self.n_init = n_init
        self.tracks: List[Trk] = []
        self.next_id = 1
        self.cost_log = None  # set to a list to record the raw cost matrices per frame
        self.last_matches: List[Tuple[int, int]] = []

    def predict(self):  # tracker_core.py:44-49
        for t in self.tracks:
            t.predict()

    def _match(self, dets):  # tracker_core.py:83-177
        def gated_metric(tracks, ds, ti, di):
            return gate_by_mahalanobis(appearance_cost(tracks, ds, ti, di), tracks, ds, ti, di)

        confirmed = [i for i, t in enumerate(self.tracks) if t.state == CONFIRMED]
        unconfirmed = [i for i, t in enumerate(self.tracks) if t.state == TENTATIVE]
        m_app, un_conf, un_d = matching_cascade(
            gated_metric, self.max_cosine_distance, self.max_age, self.tracks, dets,
            confirmed, list(range(len(dets))), self.cost_log)
        iou_cand = unconfirmed + [i for i in un_conf if self.tracks[i].time_since_update == 1]
        remaining = [i for i in un_conf if self.tracks[i].time_since_update > 1]
        if iou_cand and un_d:
            m_iou, un_iou, un_d2 = min_cost_matching(
                iou_cost, self.max_iou_distance, self.tracks, dets, iou_cand, un_d, self.cost_log)
        else:
            m_iou, un_iou, un_d2 = [], iou_cand, un_d
        return m_app + m_iou, remaining + un_iou, un_d2

    def update(self, dets: List[Det]):  # tracker_core.py:51-81
        matches, un_t, un_d = self._match(dets)
        self.last_matches = [(self.tracks[t].track_id, d) for t, d in matches]
        for t, d in matches:
            self.tracks[t].update(dets[d])
        for t in un_t:
            self.tracks[t].mark_missed()
        for d in un_d:
            self._initiate(dets[d])
        self.tracks = [t for t in self.tracks if t.state != DELETED]

    def _initiate(self, det):  # tracker_core.py:180-194, kalman_filter.py:55-83
        mean, cov = kalman.initiate(det.to_xyah())
        self.tracks.append(Trk(self.next_id, mean, cov, det, self.n_init, self.max_age,
                               self.nn_budget))
        self.next_id += 1


def crop_rect(bbox, frame_h, frame_w):
    """src/tracker/deepsort_tracker.py:143-159: int() truncation, clamp; None if empty."""
    x1, y1, x2, y2 = map(int, bbox)
    x1c, y1c = max(0, x1), max(0, y1)
    x2c, y2c = min(frame_w, x2), min(frame_h, y2)
    if x1c < x2c and y1c < y2c:
        return x1c, y1c, x2c, y2c
    return None


class DeepSORT:
    """src/tracker/deepsort_tracker.py:15-199.

    ``reid_fn(frame_bgr, rects) -> (len(rects), F) float32`` stands in for
    ``ReIDModel.extract_features_batched`` (reid_model.py:67-126); rects are the
    clamped integer crop rectangles (x1, y1, x2, y2)."""

    def __init__(self, reid_fn: Optional[Callable] = None,
                 max_cosine_distance=DEEPSORT_MAX_DIST, nn_budget=DEEPSORT_NN_BUDGET,
                 max_iou_distance=DEEPSORT_MAX_IOU_DISTANCE, max_age=DEEPSORT_MAX_AGE,
                 n_init=DEEPSORT_N_INIT, min_detection_confidence=DEEPSORT_MIN_CONFIDENCE):
        self.reid_fn = reid_fn
        self.tracker_core = TrackerCore(max_cosine_distance, nn_budget, max_iou_distance,
                                        max_age, n_init)
        self.min_detection_confidence = min_detection_confidence
        self.frame_count = 0

    def filter_indices(self, confidences, class_ids):
        """deepsort_tracker.py:88-95."""
        keep = []
        for i in range(len(confidences)):
            cid = int(class_ids[i])
            name = CLASSES[cid] if 0 <= cid < len(CLASSES) else "Unknown"
            if confidences[i] >= self.min_detection_confidence and name in CLASSES_TO_TRACK:
                keep.append(i)
        return keep

    def update(self, bboxes_xyxy, confidences, class_ids, frame_bgr=None, *,
               frame_hw: Optional[Sequence[int]] = None, planted_features=None):
        """deepsort_tracker.py:63-141.

        Either ``frame_bgr`` (features come from ``reid_fn`` on the crops) or
        ``frame_hw`` + ``planted_features`` ((N,F), aligned with the *input*
        detections; used where the tracker is tested without the ReID net)."""
        self.frame_count += 1
        self.tracker_core.predict()
        dets = self._detections(bboxes_xyxy, confidences, class_ids, frame_bgr, frame_hw, planted_features)
        self.tracker_core.update(dets)
        out = []
        for t in self.tracker_core.tracks:  # :125-141
            if t.state == CONFIRMED and t.time_since_update == 0:
                x1, y1, w, h = t.to_tlwh()
                w = max(0, w)
                h = max(0, h)
                x2, y2 = x1 + w, y1 + h
                name = CLASSES[t.class_id] if 0 <= t.class_id < len(CLASSES) else "Unknown"
                out.append((int(round(x1)), int(round(y1)), int(round(x2)), int(round(y2)),
                            t.track_id, name, float(t.confidence)))
        return out

    def probe_costs(self, bboxes_xyxy, confidences, class_ids, frame_bgr=None, *, frame_hw=None,
                    planted_features=None):
        """What the matching of the NEXT ``update`` with these arguments would be computed from,
        without changing any state: for every live track (track-list order) against every
        filtered detection, the appearance cost (matching.py:144-217; INFTY_COST rows for
        tentative tracks, which never enter the cascade, tracker_core.py:112-117) and the squared
        Mahalanobis distance to the predicted state (linear_assignment.py:160-212).
        Returns (track_ids [T], app_cost [T, D] float32, gate_d2 [T, D] float32)."""
        import copy
        tracks = copy.deepcopy(self.tracker_core.tracks)
        for t in tracks:
            t.predict()
        dets = self._detections(bboxes_xyxy, confidences, class_ids, frame_bgr, frame_hw, planted_features)
        T, D = len(tracks), len(dets)
        app = np.full((T, D), INFTY_COST, dtype=F32)
        d2 = np.zeros((T, D), dtype=F32)
        if T and D:
            conf = [i for i, t in enumerate(tracks) if t.state == CONFIRMED]
            if conf:
                app[conf] = appearance_cost(tracks, dets, conf, list(range(D)))
            meas = np.asarray([d.to_xyah() for d in dets])
            for r, t in enumerate(tracks):
                d2[r] = kalman.gating_distance(t.mean, t.cov, meas)
        return np.asarray([t.track_id for t in tracks], dtype=np.int64), app, d2

    def _detections(self, bboxes_xyxy, confidences, class_ids, frame_bgr, frame_hw, planted_features):
        """deepsort_tracker.py:82-121, :161-199: filter, crops, features, Detection objects."""
        bboxes_xyxy = np.asarray(bboxes_xyxy)
        keep = self.filter_indices(confidences, class_ids)
        dets: List[Det] = []
        if keep:
            fb = bboxes_xyxy[keep]
            fc = np.asarray(confidences)[keep]
            fk = np.asarray(class_ids)[keep]
            h, w = frame_bgr.shape[:2] if frame_bgr is not None else frame_hw
            rects = [crop_rect(b, h, w) for b in fb]
            valid = [i for i, r in enumerate(rects) if r is not None]
            feats = {}
            if valid:
                if planted_features is not None:
                    pf = np.asarray(planted_features, dtype=F32)
                    feats = {i: pf[keep[i]] for i in valid}
                else:
                    out = self.reid_fn(frame_bgr, [rects[i] for i in valid])
                    if out.ndim == 2 and out.shape[0] == len(valid):  # :174-178
                        feats = {i: out[k] for k, i in enumerate(valid)}
            for i in range(len(fb)):  # :180-198
                x1, y1, x2, y2 = fb[i]
                tlwh = np.array([x1, y1, x2 - x1, y2 - y1], dtype=F32)
                dets.append(Det(tlwh, float(fc[i]), int(fk[i]), feats.get(i)))
        return dets
