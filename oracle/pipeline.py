"""CPU restatement of the whole per-frame hot path: ``YOLODetector.detect`` followed by
``DeepSORT.update`` (oracle; test infrastructure and the timed CPU baseline of bench.py).

  Detector.detect      /root/reference/src/detector/yolo_detector.py:68-149
                       (preprocess_yolo_input -> engine -> num_dets/bboxes/scores/labels ->
                        confidence filter -> scale_bboxes)
  ReID.__call__        /root/reference/src/tracker/reid_model.py:67-126
                       (per-crop preprocess_reid_input -> concat -> engine)
  DeepSORT             oracle.tracker (src/tracker/deepsort_tracker.py:63-141)
The two engines are the PyTorch-CPU fp32 nets of oracle.nets (the reference's TensorRT
engines and their ONNX sources are not available: PARITY UNPINNED for the NN arithmetic).
"""
import numpy as np
import torch

from . import detect_post, image_ops, nets
from .constants import (YOLO_CONF_THRESHOLD, YOLO_INPUT_SHAPE, YOLO_NMS_THRESHOLD, YOLO_TOPK)
from .tracker import DeepSORT


class Detector:
    def __init__(self, blob_path, conf_threshold=YOLO_CONF_THRESHOLD, nms_threshold=YOLO_NMS_THRESHOLD,
                 topk=YOLO_TOPK, max_candidates=1024, bias_overrides=None):
        self.net = nets.load_net(blob_path)
        for name, b in (bias_overrides or {}).items():
            self.net.w[name + ".bias"] = torch.as_tensor(np.asarray(b, np.float32))
        self.conf_threshold, self.nms_threshold = conf_threshold, nms_threshold
        self.topk, self.max_candidates = topk, max_candidates

    def head(self, frame_bgr):
        x, _, _ = image_ops.preprocess_yolo_input(frame_bgr, YOLO_INPUT_SHAPE)
        return self.net.head_flat(torch.from_numpy(x))[0].numpy()

    def engine_outputs(self, frame_bgr):
        return detect_post.engine_outputs(self.head(frame_bgr), min(self.conf_threshold, YOLO_CONF_THRESHOLD),
                                          self.nms_threshold, self.topk, self.max_candidates)

    def detect(self, frame_bgr):
        n, b, s, l = self.engine_outputs(frame_bgr)
        return detect_post.detect_postprocess(n, b, s, l, frame_bgr.shape[:2], self.conf_threshold)


class ReID:
    def __init__(self, blob_path, bias_overrides=None):
        self.net = nets.load_net(blob_path)
        for name, b in (bias_overrides or {}).items():
            self.net.w[name + ".bias"] = torch.as_tensor(np.asarray(b, np.float32))

    def __call__(self, frame_bgr, rects):
        x = image_ops.reid_batch(frame_bgr, rects)
        if x.shape[0] == 0:
            return np.empty((0, 512), np.float32)
        return self.net.forward(torch.from_numpy(x)).numpy()


class Pipeline:
    """One stream: detect(frame) then update(dets, frame), as src/aicamera_tracker.py:180,193."""

    def __init__(self, yolo_blob, reid_blob, yolo_bias=None, **tracker_kw):
        self.detector = Detector(yolo_blob, bias_overrides=yolo_bias)
        self.tracker = DeepSORT(reid_fn=ReID(reid_blob), **tracker_kw)

    def step(self, frame_bgr):
        boxes, scores, cls, _ = self.detector.detect(frame_bgr)
        return self.tracker.update(boxes, scores, cls, frame_bgr)
