"""``DeepSORT`` with the reference's interface (``/root/reference/src/tracker/deepsort_tracker.py``):
``DeepSORT(reid_model_path, reid_input_shape, max_cosine_distance, nn_budget, max_iou_distance,
max_age, n_init, min_detection_confidence)`` (:21-30) and
``update(bboxes_xyxy, confidences, class_ids, frame_bgr) -> [(x1, y1, x2, y2, id, class_name, conf)]``
(:63-141).

Everything between the arguments and the returned list runs on the device: class/confidence
filter, crops, ReID network, Kalman predict, gated appearance cascade, IoU matching, LSAP,
updates, initiation, pruning and output formatting.  The tracker state lives in HBM; each
instance has its own id counter starting at 1 (the reference uses a process-global counter
reset by every TrackerCore, tracker_core.py:42).

Like ``YOLODetector.detect``, the device side of ``update`` is captured in a CUDA graph after two eager calls and
replayed per frame (keyed by the frame buffer it reads); ``AICAM_NO_FACADE_GRAPH=1`` keeps the calls eager."""
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import config
from .pipeline import BatchTracker


class DeepSORT:
    def __init__(self,
                 reid_model_path: str = str(config.REID_ENGINE_PATH),
                 reid_input_shape: Tuple[int, int] = config.REID_INPUT_SHAPE,
                 max_cosine_distance: float = config.DEEPSORT_MAX_DIST,
                 nn_budget: Optional[int] = config.DEEPSORT_NN_BUDGET,
                 max_iou_distance: float = config.DEEPSORT_MAX_IOU_DISTANCE,
                 max_age: int = config.DEEPSORT_MAX_AGE,
                 n_init: int = config.DEEPSORT_N_INIT,
                 min_detection_confidence: float = config.DEEPSORT_MIN_CONFIDENCE,
                 device: Optional[torch.device] = None, max_dets: int = 128, max_tracks: int = 256):
        if tuple(reid_input_shape) != tuple(config.REID_INPUT_SHAPE):
            raise RuntimeError("this build supports the 128x64 ReID input only")
        self._trk = BatchTracker(reid_model_path, 1, device, max_dets=max_dets, max_tracks=max_tracks,
                                 max_crops=max_dets, max_cosine_distance=max_cosine_distance, nn_budget=nn_budget,
                                 max_iou_distance=max_iou_distance, max_age=max_age, n_init=n_init,
                                 min_detection_confidence=min_detection_confidence)
        self.device = self._trk.device
        self.reid_model = self._trk.reid
        self.tracker_core = self._trk
        self.min_detection_confidence = min_detection_confidence
        self.frame_count = 0
        self.K, self.T = max_dets, max_tracks
        K, T = self.K, self.T
        # [num, 3 pad words | boxes 4K | scores K | labels K]: the box table must stay 16-byte aligned
        self._in_host = torch.zeros(4 + 6 * K, dtype=torch.int32).pin_memory()
        self._in_dev = torch.zeros(4 + 6 * K, dtype=torch.int32, device=self.device)
        self._out_host = torch.zeros(1 + 7 * T, dtype=torch.int32).pin_memory()
        self._out_dev = torch.zeros(1 + 7 * T, dtype=torch.int32, device=self.device)
        self._frame_host = None
        self._frame_dev = None
        self._graphs = {}    # (frame buffer address, shape) -> CUDAGraph of _run(); False: capture failed
        self._calls = {}
        self._no_graph = bool(os.environ.get("AICAM_NO_FACADE_GRAPH"))
        print("DeepSORT Tracker initialized.")

    def _stage(self, frame_bgr):
        """(device frame [1, H, W, 3], from_host): a CUDA tensor is used where it lies; a numpy frame goes into the pinned
        buffer and is uploaded by _run()."""
        if isinstance(frame_bgr, torch.Tensor):
            return (frame_bgr if frame_bgr.dim() == 4 else frame_bgr.unsqueeze(0)), False
        shape = tuple(frame_bgr.shape)
        if self._frame_host is None or tuple(self._frame_host.shape[1:]) != shape:
            self._frame_host = torch.empty((1,) + shape, dtype=torch.uint8).pin_memory()
            self._frame_dev = torch.empty((1,) + shape, dtype=torch.uint8, device=self.device)
        self._frame_host[0].numpy()[...] = frame_bgr
        return self._frame_dev, True

    def _run(self, frames, from_host):
        """Device side of update(): detections (and, for a numpy frame, the frame) from pinned memory, K5-K12, packed result
        -> pinned host.  Asynchronous; fixed addresses: capturable."""
        K, T = self.K, self.T
        if from_host:
            self._frame_dev.copy_(self._frame_host, non_blocking=True)
        d = self._in_dev
        d.copy_(self._in_host, non_blocking=True)
        num = d[0:1]
        boxes = d[4:4 + 4 * K].view(torch.float32).view(1, K, 4)
        scores = d[4 + 4 * K:4 + 5 * K].view(torch.float32).view(1, K)
        labels = d[4 + 5 * K:4 + 6 * K].view(1, K)
        out_tracks, out_conf, out_count = self._trk.update(frames, num, boxes, scores, labels)
        o = self._out_dev
        o[0:1].copy_(out_count)
        o[1:1 + 6 * T].copy_(out_tracks.reshape(-1))
        o[1 + 6 * T:].view(torch.float32).copy_(out_conf.reshape(-1))
        self._out_host.copy_(o, non_blocking=True)

    def _run_graphed(self, frames, from_host):
        key = (frames.data_ptr(), tuple(frames.shape), from_host)
        g = False if self._no_graph else self._graphs.get(key)
        if len(self._calls) > 64:  # (callers that hand in a fresh tensor per frame: nothing to replay, keep the table small)
            self._calls.clear()
        n = self._calls[key] = self._calls.get(key, 0) + 1
        if g is None and n > 2 and len(self._graphs) < 8:
            try:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run(frames, from_host)
            except Exception as e:  # (stay eager for this buffer)
                print(f"DeepSORT: CUDA graph capture failed ({e}); running eager")
                g = False
            self._graphs[key] = g
        if g:
            g.replay()
        else:
            self._run(frames, from_host)

    def update(self, yolo_bboxes_xyxy: np.ndarray, yolo_confidences: np.ndarray, yolo_class_ids: np.ndarray,
               original_frame_bgr) -> List[Tuple[int, int, int, int, int, str, float]]:
        self.frame_count += 1
        K, T = self.K, self.T
        n = len(yolo_bboxes_xyxy)
        if n > K:
            raise RuntimeError("update: %d detections exceed max_dets=%d" % (n, K))
        h = self._in_host.numpy()
        h[0] = n
        if n:
            h[4:4 + 4 * n].view(np.float32)[:] = np.asarray(yolo_bboxes_xyxy, dtype=np.float32).reshape(-1)
            h[4 + 4 * K:4 + 4 * K + n].view(np.float32)[:] = np.asarray(yolo_confidences, dtype=np.float32)
            h[4 + 5 * K:4 + 5 * K + n] = np.asarray(yolo_class_ids).astype(np.int32)
        frames, from_host = self._stage(original_frame_bgr)
        with torch.cuda.device(self.device):
            self._run_graphed(frames, from_host)
        torch.cuda.current_stream(self.device).synchronize()
        r = self._out_host.numpy()
        m = int(r[0])
        tr = r[1:1 + 6 * T].reshape(T, 6)
        cf = r[1 + 6 * T:].view(np.float32)
        out = []
        for k in range(m):
            x1, y1, x2, y2, tid, cid = (int(v) for v in tr[k])
            name = config.CLASSES[cid] if 0 <= cid < len(config.CLASSES) else "Unknown"
            out.append((x1, y1, x2, y2, tid, name, float(cf[k])))
        return out
