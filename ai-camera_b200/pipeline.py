"""Batched, device-resident detect -> track path over many video streams.

The reference processes one frame of one stream per call, on the CPU except for the two
TensorRT engines (``/root/reference/src/aicamera_tracker.py:169-240``).  Here a batch of
frames (one per stream) goes through K1-K12 without leaving the GPU and without a host
synchronisation: detections, crop counts and track tables stay in fixed-capacity device
buffers.  ``YOLODetector`` / ``DeepSORT`` (the reference-shaped facades) are thin single-stream
wrappers around ``BatchDetector`` / ``BatchTracker``.
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib, config
from .trt_engine import TRTEngine


def frame_geometry(frames: torch.Tensor):
    """(n, h, w, is_nv12) of a batch of frames: [n, H, W, 3] packed BGR or [n, H*3/2, W] NV12."""
    if frames.dtype != torch.uint8 or not frames.is_contiguous():
        raise _lib.AicamError(-1, "frames must be a contiguous uint8 tensor")
    if frames.dim() == 4 and frames.shape[3] == 3:
        return frames.shape[0], frames.shape[1], frames.shape[2], False
    if frames.dim() == 3 and frames.shape[1] % 3 == 0:
        return frames.shape[0], frames.shape[1] * 2 // 3, frames.shape[2], True
    raise _lib.AicamError(-1, "frames must be [n, H, W, 3] (BGR) or [n, H*3/2, W] (NV12)")


class BatchDetector:
    """K1-K4 + detect() post-processing for ``batch`` frames of one size."""

    def __init__(self, engine_path, batch: int, device=None, conf_threshold=config.YOLO_CONF_THRESHOLD,
                 nms_threshold=config.YOLO_NMS_THRESHOLD, topk=config.YOLO_TOPK,
                 max_candidates=config.YOLO_MAX_CANDIDATES):
        self.engine = TRTEngine(engine_path, device, max_batch=batch, topk=topk,
                                score_threshold=min(conf_threshold, config.YOLO_CONF_THRESHOLD),
                                nms_threshold=nms_threshold, max_candidates=max_candidates)
        if self.engine.kind != _lib.KIND_YOLOV8:
            raise RuntimeError("BatchDetector needs a yolov8 weight blob")
        self.lib = _lib.load()
        self.device = self.engine.device
        self.batch, self.topk = batch, topk
        self.conf_threshold = conf_threshold
        e, dev = self.engine, self.device
        self.num_dets = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.boxes_lb = torch.zeros((batch, topk, 4), dtype=torch.float32, device=dev)
        self.boxes = torch.zeros((batch, topk, 4), dtype=torch.float32, device=dev)
        self.scores = torch.zeros((batch, topk), dtype=torch.float32, device=dev)
        self.labels = torch.zeros((batch, topk), dtype=torch.int32, device=dev)
        self._nms = _lib.NmsParams(e.score_threshold, e.nms_threshold, topk, e.max_candidates, 0, 0)
        # 0: NHWC4 input (aicam_preprocess format 1), 1: 2x2 pixel blocks (format 2), 2: 4x4 pixel blocks (format 3, yolov8n)
        self._s2d = int(self.lib.aicam_engine_accepts_s2d(e.handle))

    def detect(self, frames: torch.Tensor):
        """frames: uint8 cuda [n<=batch, H, W, 3] BGR, or [n, H*3/2, W] NV12 (Y plane + interleaved UV plane, the
        surface format of hardware decoders: half the bytes).  Returns device tensors (num_dets [n],
        boxes [n,topk,4] in frame pixels, scores [n,topk], labels [n,topk]); views of internal
        buffers, valid until the next call; asynchronous on the current stream."""
        n, h, w, nv12 = frame_geometry(frames)
        pre = self.lib.aicam_preprocess_nv12 if nv12 else self.lib.aicam_preprocess
        if n > self.batch:
            raise _lib.AicamError(-4, "detect: %d frames exceed the detector's batch %d" % (n, self.batch))
        e, st = self.engine, _lib.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            # same bytes in 2x2 / 4x4 pixel blocks (formats 2 / 3): the stride-2 stem runs as a stride-1 window over the blocks
            _lib.check(pre(_lib.ptr(frames), n, h, w, 1 + self._s2d, _lib.ptr(e._nhwc), st))
            self._nms.frame_h, self._nms.frame_w = h, w
            # network + decode (fused into the Detect-head epilogues) + NMS + un-letterboxing: one entry
            _lib.check(self.lib.aicam_yolo_detect(
                e.handle, _lib.ptr(e._nhwc), self._s2d, n, C.byref(self._nms),
                _lib.ptr(e._head) if e._head is not None else None, _lib.ptr(self.num_dets),
                _lib.ptr(self.boxes_lb), _lib.ptr(self.boxes), _lib.ptr(self.scores), _lib.ptr(self.labels),
                _lib.ptr(e._ws), e._ws.numel(), st))
        return self.num_dets[:n], self.boxes[:n], self.scores[:n], self.labels[:n]

    def launches_per_step(self):
        # K1 + network + NMS (+ the decode kernel when the engine cannot decode inside its Detect-head epilogues)
        return 1 + self.engine.launches_per_forward() + (1 if self.engine.fused_decode else 2)


class BatchTracker:
    """K5-K12 for ``n_streams`` independent streams (one DeepSORT state each)."""

    def __init__(self, reid_engine_path, n_streams: int, device=None, max_dets=config.YOLO_TOPK,
                 max_tracks: int = 256, max_crops: Optional[int] = None,
                 max_cosine_distance=config.DEEPSORT_MAX_DIST, nn_budget=config.DEEPSORT_NN_BUDGET,
                 max_iou_distance=config.DEEPSORT_MAX_IOU_DISTANCE, max_age=config.DEEPSORT_MAX_AGE,
                 n_init=config.DEEPSORT_N_INIT, min_detection_confidence=config.DEEPSORT_MIN_CONFIDENCE):
        self.max_crops = int(max_crops if max_crops is not None else min(n_streams * max_dets, 4096))
        self.reid = TRTEngine(reid_engine_path, device, max_batch=self.max_crops)
        if self.reid.kind != _lib.KIND_REID:
            raise RuntimeError("BatchTracker needs a reid weight blob")
        if nn_budget is None:
            raise RuntimeError("nn_budget=None (unbounded galleries) is not supported on the device")
        self.lib = _lib.load()
        self.device = dev = self.reid.device
        self.S, self.K, self.T = n_streams, max_dets, max_tracks
        self.min_conf = float(min_detection_confidence)
        self.F = self.reid.feature_dim
        cfg = _lib.TrackerConfig(n_streams, max_tracks, max_dets, self.F, float(max_cosine_distance),
                                 float(max_iou_distance), int(max_age), int(n_init), int(nn_budget), dev.index)
        self._h = C.c_void_p()
        _lib.check(self.lib.aicam_tracker_create(C.byref(cfg), C.byref(self._h)))
        S, K = n_streams, max_dets
        i32 = dict(dtype=torch.int32, device=dev)
        self.det_index = torch.zeros((S, K), **i32)
        self.det_count = torch.zeros(S, **i32)
        self.crop_slot = torch.zeros((S, K), **i32)
        self.crop_rect = torch.zeros((self.max_crops, 5), **i32)
        self.crop_count = torch.zeros(2, **i32)  # [crops written, high-water mark of crops wanted]
        # crops go straight into the layout the engine's first kernel reads (NHWC8 for the fused stem)
        self._nhwc8 = bool(self.lib.aicam_engine_accepts_nhwc8(self.reid.handle))
        self.crops = torch.zeros((self.max_crops, config.REID_INPUT_SHAPE[0], config.REID_INPUT_SHAPE[1],
                                  8 if self._nhwc8 else 4),
                                 dtype=torch.bfloat16, device=dev)
        self.feats = torch.zeros((self.max_crops, self.F), dtype=torch.float32, device=dev)
        self.out_tracks = torch.zeros((S, max_tracks, 6), **i32)
        self.out_conf = torch.zeros((S, max_tracks), dtype=torch.float32, device=dev)
        self.out_count = torch.zeros(S, **i32)
        self._mask = config.tracked_class_mask()
        # optional running totals (crops, reported tracks) kept on the device, for benchmarks: no host sync
        self.count_stats = False
        self.crop_total = torch.zeros(1, dtype=torch.int64, device=dev)
        self.track_total = torch.zeros(1, dtype=torch.int64, device=dev)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self.lib.aicam_tracker_destroy(h)
            self._h = None

    def reset(self):
        _lib.check(self.lib.aicam_tracker_reset(self._h, _lib.stream_ptr(self.device)))
        self.crop_count.zero_()

    def update(self, frames, num_dets, boxes, scores, labels):
        """One frame per stream.  frames uint8 cuda [S,H,W,3] BGR or [S,H*3/2,W] NV12; detections as BatchDetector
        returns them ([S], [S,K,4], [S,K], [S,K]).  Returns device (out_tracks [S,T,6] int32 =
        x1,y1,x2,y2,id,class, out_conf [S,T], out_count [S]); asynchronous."""
        S, h, w, nv12 = frame_geometry(frames)
        crops_fn = self.lib.aicam_reid_crops_nv12 if nv12 else self.lib.aicam_reid_crops
        if S != self.S or boxes.shape[1] != self.K:
            raise _lib.AicamError(-4, "update: expected %d streams x %d detections" % (self.S, self.K))
        st = _lib.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(crops_fn(
                _lib.ptr(frames), S, h, w, _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(labels), _lib.ptr(num_dets),
                self.K, self.min_conf, self._mask[0], self._mask[1], 2 if self._nhwc8 else 1, self.max_crops,
                _lib.ptr(self.det_index),
                _lib.ptr(self.det_count), _lib.ptr(self.crop_slot), _lib.ptr(self.crop_rect), _lib.ptr(self.crops),
                _lib.ptr(self.crop_count), st))
            fwd = self.lib.aicam_reid_forward_nhwc8 if self._nhwc8 else self.lib.aicam_reid_forward
            _lib.check(fwd(self.reid.handle, _lib.ptr(self.crops), self.max_crops, _lib.ptr(self.crop_count),
                           _lib.ptr(self.feats), st))
            _lib.check(self.lib.aicam_tracker_step(
                self._h, _lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(labels), self.K, _lib.ptr(self.det_index),
                _lib.ptr(self.det_count), _lib.ptr(self.crop_slot), _lib.ptr(self.feats), _lib.ptr(self.out_tracks),
                _lib.ptr(self.out_conf), _lib.ptr(self.out_count), st))
            if self.count_stats:
                self.crop_total += self.crop_count[0:1]
                self.track_total += self.out_count.sum()
        return self.out_tracks, self.out_conf, self.out_count

    def overflow(self):
        import numpy as np
        f = np.zeros(self.S, np.int32)
        _lib.check(self.lib.aicam_tracker_overflow(self._h, _lib.ptr(f)))
        # bit 2 (every stream): some step wanted more ReID crops than max_crops, so detections lost their feature
        if int(self.crop_count[1].item()) > self.max_crops:
            f |= 4
        return f

    def snapshot(self, stream_index=0):
        import numpy as np
        ints = np.zeros((self.T, 7), np.int32)
        floats = np.zeros((self.T, 25), np.float32)
        n = self.lib.aicam_tracker_snapshot(self._h, stream_index, _lib.ptr(ints), _lib.ptr(floats), self.T)
        if n < 0:
            _lib.check(n)
        return ints[:n], floats[:n]

    def launches_per_step(self):
        return 2 + self.reid.launches_per_forward() - (1 if self._nhwc8 else 0) + 3  # no NHWC8 repack kernel


class TrackingPipeline:
    """detect + track for ``n_streams`` streams of one frame size; one call per time step."""

    def __init__(self, yolo_engine_path, reid_engine_path, n_streams: int, device=None, max_tracks: int = 256,
                 max_crops: Optional[int] = None, conf_threshold=config.YOLO_CONF_THRESHOLD, topk=config.YOLO_TOPK,
                 max_candidates=config.YOLO_MAX_CANDIDATES, **tracker_kw):
        self.detector = BatchDetector(yolo_engine_path, n_streams, device, conf_threshold=conf_threshold, topk=topk,
                                      max_candidates=max_candidates)
        # detect() drops detections below conf_threshold before update() sees them (yolo_detector.py:131); on the
        # device that filter and DeepSORT's own (deepsort_tracker.py:88-95) are one comparison
        tracker_kw.setdefault("min_detection_confidence", max(config.DEEPSORT_MIN_CONFIDENCE, conf_threshold))
        self.tracker = BatchTracker(reid_engine_path, n_streams, self.detector.device, max_dets=self.detector.topk,
                                    max_tracks=max_tracks, max_crops=max_crops, **tracker_kw)
        self.device = self.detector.device
        self.n_streams = n_streams

    def step(self, frames: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        num, boxes, scores, labels = self.detector.detect(frames)
        return self.tracker.update(frames, num, boxes, scores, labels)

    def launches_per_step(self):
        return self.detector.launches_per_step() + self.tracker.launches_per_step()
