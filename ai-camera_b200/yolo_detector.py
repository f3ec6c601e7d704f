"""``YOLODetector`` with the reference's interface (``/root/reference/src/detector/yolo_detector.py``):
``YOLODetector(engine_path, input_shape, conf_threshold, nms_threshold, device)`` (:15-21) and
``detect(frame_bgr) -> (bboxes_xyxy, scores, class_ids, filtered_indices)`` (:68-149).

What changes underneath: the frame is uploaded as uint8 (6.2 MB for 1080p) instead of being
letterboxed on the CPU and uploaded as fp32 (:86-91); letterbox, network, decode, NMS,
confidence filter and un-letterboxing all run on the device; one D2H copy returns the
result.  Return types and the empty-result convention (:126) are the reference's.

After two eager calls the whole device side of ``detect`` - frame upload from the pinned buffer, K1, the network,
decode, NMS, the result copy - is captured in a CUDA graph and replayed per frame (one launch instead of ~80 through
ctypes); ``AICAM_NO_FACADE_GRAPH=1`` keeps the calls eager."""
import os
from typing import Tuple

import numpy as np
import torch

from . import config
from .pipeline import BatchDetector


class YOLODetector:
    def __init__(self,
                 engine_path: str = str(config.YOLO_ENGINE_PATH),
                 input_shape: Tuple[int, int] = config.YOLO_INPUT_SHAPE,
                 conf_threshold: float = config.YOLO_CONF_THRESHOLD,
                 nms_threshold: float = config.YOLO_NMS_THRESHOLD,
                 device: torch.device = torch.device('cuda:0' if torch.cuda.is_available() else 'cpu')):
        if tuple(input_shape) != tuple(config.YOLO_INPUT_SHAPE):
            raise RuntimeError("this build supports the 640x640 network input only")
        self.engine_path = engine_path
        self.input_shape = input_shape
        self.conf_threshold = conf_threshold
        self.nms_threshold = nms_threshold
        self.device = torch.device(device)
        self._det = BatchDetector(engine_path, 1, self.device, conf_threshold, nms_threshold)
        self.trt_engine = self._det.engine
        self.device = self._det.device
        self.input_name = self.trt_engine.get_input_details()[0].name
        self.output_names = {'num_dets': 'num_dets', 'bboxes': 'bboxes', 'scores': 'scores', 'labels': 'labels'}
        k = self._det.topk
        # one pinned staging buffer for the result: [num | boxes 4k | scores k | labels k] as 32-bit words
        self._host = torch.empty(1 + 6 * k, dtype=torch.int32).pin_memory()
        self._pack = torch.empty(1 + 6 * k, dtype=torch.int32, device=self.device)
        self._frame_dev = None
        self._frame_host = None
        # the frame of the last detect() call as it lies in HBM (uint8 [1, H, W, 3]): DeepSORT.update accepts it in
        # place of the numpy frame, so a detect -> update pair uploads the frame once
        self.device_frame = None
        self._graph = None   # CUDAGraph of _run() for the current frame shape; False: capture failed / disabled
        self._calls = 0
        if os.environ.get("AICAM_NO_FACADE_GRAPH"):
            self._graph = False
        print(f"YOLODetector initialized with engine: {engine_path}")
        print(f"  Input name: {self.input_name}, Input shape: {self.input_shape}")

    def _stage(self, frame_bgr: np.ndarray):
        """The frame into the pinned host buffer (buffers are re-made, and the graph dropped, when the frame size changes)."""
        shape = tuple(frame_bgr.shape)
        if self._frame_host is None or tuple(self._frame_host.shape[1:]) != shape:
            self._frame_host = torch.empty((1,) + shape, dtype=torch.uint8).pin_memory()
            self._frame_dev = torch.empty((1,) + shape, dtype=torch.uint8, device=self.device)
            if self._graph is not False:
                self._graph, self._calls = None, 0
        self._frame_host[0].numpy()[...] = frame_bgr
        self.device_frame = self._frame_dev

    def _run(self):
        """Device side of detect(): pinned frame -> HBM, K1-K4, packed result -> pinned host.  Asynchronous; every address
        is fixed, so the sequence can be captured once and replayed."""
        k = self._det.topk
        self._frame_dev.copy_(self._frame_host, non_blocking=True)
        num, boxes, scores, labels = self._det.detect(self._frame_dev)
        p = self._pack
        p[0:1].copy_(num)
        p[1:1 + 4 * k].view(torch.float32).copy_(boxes.reshape(-1))
        p[1 + 4 * k:1 + 5 * k].view(torch.float32).copy_(scores.reshape(-1))
        p[1 + 5 * k:].copy_(labels.reshape(-1))
        self._host.copy_(p, non_blocking=True)

    def _run_graphed(self):
        self._calls += 1
        if self._graph is None and self._calls > 2:
            try:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run()
                self._graph = g
            except Exception as e:  # (stay eager)
                print(f"YOLODetector: CUDA graph capture failed ({e}); running eager")
                self._graph = False
        if self._graph:
            self._graph.replay()
        else:
            self._run()

    def detect(self, frame_bgr: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        empty = (np.empty((0, 4)), np.empty(0), np.empty(0), np.empty(0, dtype=int))
        if frame_bgr.ndim != 3 or frame_bgr.shape[2] != 3 or frame_bgr.dtype != np.uint8:
            raise ValueError("detect expects an HxWx3 uint8 BGR frame")
        k = self._det.topk
        try:
            self._stage(frame_bgr)
            with torch.cuda.device(self.device):
                self._run_graphed()
            torch.cuda.current_stream(self.device).synchronize()
        except Exception as e:  # yolo_detector.py:117-122: report and return empty
            print(f"Error processing engine outputs: {e}")
            return empty
        h = self._host.numpy()
        n = int(h[0])
        if n == 0:
            return empty
        bboxes = h[1:1 + 4 * k].view(np.float32).reshape(k, 4)[:n].copy()
        sc = h[1 + 4 * k:1 + 5 * k].view(np.float32)[:n].copy()
        cls = h[1 + 5 * k:][:n].astype(np.int32)
        confident = sc >= self.conf_threshold  # yolo_detector.py:131
        bboxes, sc, cls = bboxes[confident], sc[confident], cls[confident]
        if bboxes.shape[0] == 0:
            return empty
        return bboxes, sc, cls, np.where(confident)[0]
