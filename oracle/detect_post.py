"""Detect-head decode, in-engine NMS and ``YOLODetector.detect`` post-processing
(oracle; test infrastructure).

PARITY UNPINNED for decode + NMS: in the reference these run inside the TensorRT engine,
which returns four tensors ``num_dets / bboxes / scores / labels``
(``/root/reference/src/detector/yolo_detector.py:49-54,108-112``); the plugin and its
parameters live in an un-vendored ONNX file.  This build DEFINES them:
  decode   Ultralytics Detect: DFL = softmax over 16 bins per side -> expectation,
           xyxy = (ax - l, ay - t, ax + r, ay + b) * stride with anchor centres
           (ix + 0.5, iy + 0.5); class score = sigmoid(logit); per anchor the best class
           (lowest index on ties).
  select   score >= score_thr; at most ``max_candidates`` highest-scoring anchors
           (ties: lower anchor index first).
  NMS      class-aware greedy: visit candidates in (score desc, anchor asc) order; a
           candidate is dropped if an already kept candidate of the SAME class has
           IoU > iou_thr (strict); stop after ``topk`` keeps.
  IoU      float32, each operation rounded: inter / ((area_a + area_b) - inter).
The part after the engine IS reference code and is restated exactly:
  ``detect`` post-processing   yolo_detector.py:107-149
  ``scale_bboxes``             src/utils/image_processing.py:141-183 (oracle.image_ops)
"""
import numpy as np

from . import image_ops
from .constants import YOLO_CONF_THRESHOLD, YOLO_NMS_THRESHOLD, YOLO_TOPK

F32 = np.float32
REG_MAX = 16
STRIDES = (8, 16, 32)


def make_anchors(input_hw=(640, 640)):
    """(A,2) anchor centres in grid units and (A,) strides, level-major then row-major."""
    pts, strides = [], []
    for s in STRIDES:
        h, w = input_hw[0] // s, input_hw[1] // s
        ys, xs = np.meshgrid(np.arange(h, dtype=F32) + F32(0.5), np.arange(w, dtype=F32) + F32(0.5),
                             indexing="ij")
        pts.append(np.stack([xs.ravel(), ys.ravel()], 1))
        strides.append(np.full(h * w, s, F32))
    return np.concatenate(pts).astype(F32), np.concatenate(strides)


def decode(head, input_hw=(640, 640)):
    """head (A, 64+nc) float32 -> boxes (A,4) xyxy letterbox px, scores (A,), labels (A,) int32."""
    head = np.asarray(head, F32)
    anchors, strides = make_anchors(input_hw)
    d = head[:, :4 * REG_MAX].reshape(-1, 4, REG_MAX)
    d = d - d.max(axis=2, keepdims=True)
    e = np.exp(d.astype(F32))
    p = e / e.sum(axis=2, keepdims=True)
    dist = (p * np.arange(REG_MAX, dtype=F32)).sum(axis=2).astype(F32)  # l, t, r, b
    x1 = (anchors[:, 0] - dist[:, 0]) * strides
    y1 = (anchors[:, 1] - dist[:, 1]) * strides
    x2 = (anchors[:, 0] + dist[:, 2]) * strides
    y2 = (anchors[:, 1] + dist[:, 3]) * strides
    logits = head[:, 4 * REG_MAX:]
    labels = np.argmax(logits, axis=1).astype(np.int32)  # first index on ties
    best = logits[np.arange(len(logits)), labels]
    scores = (F32(1.0) / (F32(1.0) + np.exp(-best))).astype(F32)
    return np.stack([x1, y1, x2, y2], 1).astype(F32), scores, labels


def iou_one_to_many(b, others):
    """float32 IoU with the operation order the device kernel uses."""
    b = b.astype(F32)
    o = others.astype(F32)
    iw = np.maximum(F32(0), np.minimum(b[2], o[:, 2]) - np.maximum(b[0], o[:, 0]))
    ih = np.maximum(F32(0), np.minimum(b[3], o[:, 3]) - np.maximum(b[1], o[:, 1]))
    inter = iw * ih
    area_b = (b[2] - b[0]) * (b[3] - b[1])
    area_o = (o[:, 2] - o[:, 0]) * (o[:, 3] - o[:, 1])
    union = (area_b + area_o) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        v = inter / union
    return np.where(union > 0, v, F32(0)).astype(F32)


def select_and_nms(boxes, scores, labels, score_thr=YOLO_CONF_THRESHOLD, iou_thr=YOLO_NMS_THRESHOLD,
                   topk=YOLO_TOPK, max_candidates=1024):
    """-> (keep_anchor_indices int64 in output order, candidate_indices int64 in visit order)."""
    scores = np.asarray(scores, F32)
    cand = np.nonzero(scores >= F32(score_thr))[0]
    order = cand[np.lexsort((cand, -scores[cand].astype(np.float64)))]  # score desc, anchor asc
    order = order[:max_candidates]
    keep = []
    thr = F32(iou_thr)
    for i in order:
        if len(keep) >= topk:
            break
        if keep:
            k = np.asarray(keep)
            same = labels[k] == labels[i]
            if same.any() and (iou_one_to_many(boxes[i], boxes[k[same]]) > thr).any():
                continue
        keep.append(int(i))
    return np.asarray(keep, np.int64), order.astype(np.int64)


def engine_outputs(head, score_thr=YOLO_CONF_THRESHOLD, iou_thr=YOLO_NMS_THRESHOLD, topk=YOLO_TOPK,
                   max_candidates=1024, input_hw=(640, 640)):
    """What the engine hands to ``YOLODetector.detect``: (num_dets, bboxes (topk,4), scores (topk,),
    labels (topk,) int32), padded with zeros, boxes in letterbox pixels, score-descending."""
    boxes, scores, labels = decode(head, input_hw)
    keep, _ = select_and_nms(boxes, scores, labels, score_thr, iou_thr, topk, max_candidates)
    n = len(keep)
    ob = np.zeros((topk, 4), F32)
    os_ = np.zeros(topk, F32)
    ol = np.zeros(topk, np.int32)
    ob[:n], os_[:n], ol[:n] = boxes[keep], scores[keep], labels[keep]
    return n, ob, os_, ol


def detect_postprocess(num_dets, bboxes, scores, labels, original_shape, conf_threshold=YOLO_CONF_THRESHOLD,
                       input_shape=(640, 640)):
    """yolo_detector.py:107-149 -> (bboxes_xyxy, scores, class_ids, filtered_indices)."""
    if num_dets == 0:
        return np.empty((0, 4)), np.empty(0), np.empty(0), np.empty(0, dtype=int)
    b = np.asarray(bboxes)[:num_dets]
    s = np.asarray(scores)[:num_dets]
    c = np.asarray(labels)[:num_dets].astype(np.int32)
    conf = s >= conf_threshold
    b, s, c = b[conf], s[conf], c[conf]
    if b.shape[0] == 0:
        return np.empty((0, 4)), np.empty(0), np.empty(0), np.empty(0, dtype=int)
    p = image_ops.letterbox_params(original_shape[0], original_shape[1], input_shape)
    out = image_ops.scale_bboxes(b, original_shape, (p["r"], p["r"]), (p["dw"], p["dh"]))
    return out, s, c, np.where(conf)[0]
