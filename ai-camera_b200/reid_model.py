"""``ReIDModel`` with the reference's interface (``/root/reference/src/tracker/reid_model.py``):
``ReIDModel(engine_path, input_shape, device)`` (:18-22) and
``extract_features_batched(list of BGR crops) -> (N, feature_dim) float32`` (:67-126).

The crops are packed into one uint8 atlas, uploaded once, and resized/normalised on the device
by the same K5 kernel the tracker uses (bit-exact with cv2's fixed-point bilinear), instead of
one cv2.resize per crop on the CPU (:84-94).  There is no CPU mock mode (:51-56, :104-107):
without a CUDA device or without the weight blob the constructor raises."""
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, config
from .trt_engine import TRTEngine


class ReIDModel:
    def __init__(self, engine_path: str = str(config.REID_ENGINE_PATH),
                 input_shape: Tuple[int, int] = config.REID_INPUT_SHAPE,
                 device: Optional[torch.device] = None, max_batch: int = 256):
        if tuple(input_shape) != tuple(config.REID_INPUT_SHAPE):
            raise RuntimeError("this build supports the 128x64 ReID input only")
        self.engine_path = engine_path
        self.input_shape = input_shape
        self.device = torch.device(device) if device is not None else torch.device(
            'cuda:0' if torch.cuda.is_available() else 'cpu')
        if not os.path.exists(self.engine_path):
            raise FileNotFoundError(f"ReID weight blob not found at {self.engine_path}")
        self.trt_engine = TRTEngine(engine_path, device=self.device, max_batch=max_batch)
        self.device = self.trt_engine.device
        self.input_name = self.trt_engine.get_input_details()[0].name
        self.output_name = self.trt_engine.get_output_details()[0].name
        self.feature_dim = self.trt_engine.feature_dim
        self.max_batch = max_batch
        self._lib = _lib.load()
        self._mask = config.tracked_class_mask()
        print(f"ReIDModel initialized with engine: {engine_path}")

    def extract_features_batched(self, image_crops_bgr: List[np.ndarray]) -> np.ndarray:
        if not image_crops_bgr:
            return np.empty((0, self.feature_dim), dtype=np.float32)
        valid = []
        for i, crop in enumerate(image_crops_bgr):
            if not isinstance(crop, np.ndarray) or crop.ndim != 3 or crop.shape[0] == 0 or crop.shape[1] == 0 \
                    or crop.shape[2] != 3:
                print(f"Warning: Invalid image crop at index {i} received in ReIDModel. Skipping.")
                continue
            valid.append(crop)
        if not valid:
            return np.empty((0, self.feature_dim), dtype=np.float32)
        try:
            out = []
            for s in range(0, len(valid), self.max_batch):
                out.append(self._run(valid[s:s + self.max_batch]))
            return np.concatenate(out, axis=0)
        except Exception as e:  # reid_model.py:121-123
            print(f"Error during ReID feature extraction: {e}")
            return np.empty((0, self.feature_dim), dtype=np.float32)

    def _run(self, crops: List[np.ndarray]) -> np.ndarray:
        n = len(crops)
        w = max(c.shape[1] for c in crops)
        h = sum(c.shape[0] for c in crops)
        atlas = torch.zeros((1, h, w, 3), dtype=torch.uint8).pin_memory()
        a = atlas[0].numpy()
        boxes = np.zeros((1, n, 4), np.float32)
        y = 0
        for i, c in enumerate(crops):
            a[y:y + c.shape[0], :c.shape[1]] = c
            boxes[0, i] = (0, y, c.shape[1], y + c.shape[0])
            y += c.shape[0]
        dev = self.device
        i32 = dict(dtype=torch.int32, device=dev)
        ad = atlas.to(dev, non_blocking=True)
        bd = torch.from_numpy(boxes).to(dev)
        sd = torch.ones((1, n), dtype=torch.float32, device=dev)
        ld = torch.zeros((1, n), **i32)
        nd = torch.full((1,), n, **i32)
        det_index, det_count, crop_slot = torch.zeros((1, n), **i32), torch.zeros(1, **i32), torch.zeros((1, n), **i32)
        crop_rect, crop_count = torch.zeros((n, 5), **i32), torch.zeros(2, **i32)
        x = torch.empty((n, self.input_shape[0], self.input_shape[1], 4), dtype=torch.bfloat16, device=dev)
        feats = torch.empty((n, self.feature_dim), dtype=torch.float32, device=dev)
        st = _lib.stream_ptr(dev)
        with torch.cuda.device(dev):
            _lib.check(self._lib.aicam_reid_crops(
                _lib.ptr(ad), 1, h, w, _lib.ptr(bd), _lib.ptr(sd), _lib.ptr(ld), _lib.ptr(nd), n, 0.0, 1, 0, 1, n,
                _lib.ptr(det_index), _lib.ptr(det_count), _lib.ptr(crop_slot), _lib.ptr(crop_rect), _lib.ptr(x),
                _lib.ptr(crop_count), st))
            _lib.check(self._lib.aicam_reid_forward(self.trt_engine.handle, _lib.ptr(x), n, None, _lib.ptr(feats), st))
        return feats.detach().cpu().numpy()
