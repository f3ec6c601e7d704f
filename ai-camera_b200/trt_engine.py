"""``TRTEngine``-shaped loader and runner on top of the aicam C library.

Mirrors ``/root/reference/src/trt_utils/trt_engine.py``: ``TRTEngine(engine_path, device)``
(:28), ``infer(dict) -> dict`` (:151-203; asynchronous on torch's current stream, outputs
allocated with torch.empty, the caller's dict is updated with the contiguous inputs, :169),
``__call__`` (:205-210), ``get_input_details`` / ``get_output_details`` (:212-216) returning
``TensorInfo(name, dtype, shape, is_dynamic)`` (:11).  ``engine_path`` names a flat ".aicw"
weight blob instead of a serialized TensorRT engine.  Constructor errors keep the reference's
types: FileNotFoundError for a missing file (:46-47), RuntimeError otherwise (:55-60).
"""
import ctypes as C
from pathlib import Path
from typing import Dict, List, NamedTuple, Optional, Tuple, Union

import torch

from . import _lib, config

TensorInfo = NamedTuple('TensorInfo', [('name', str), ('dtype', torch.dtype), ('shape', Tuple[int, ...]),
                                       ('is_dynamic', bool)])


class TRTEngine(torch.nn.Module):
    def __init__(self, engine_path: Union[str, Path], device: Optional[torch.device] = None,
                 max_batch: Optional[int] = None, topk: int = config.YOLO_TOPK,
                 score_threshold: float = config.YOLO_CONF_THRESHOLD,
                 nms_threshold: float = config.YOLO_NMS_THRESHOLD,
                 max_candidates: int = config.YOLO_MAX_CANDIDATES):
        super().__init__()
        self.engine_path = Path(engine_path)
        self.device = device if device is not None else torch.device('cuda:0' if torch.cuda.is_available() else 'cpu')
        self.device = torch.device(self.device)
        if self.device.type != 'cuda':
            raise RuntimeError("TRTEngine needs a CUDA device: this build has no CPU path.")
        if not self.engine_path.exists():
            raise FileNotFoundError(f"Weight blob not found: {self.engine_path}")
        self._lib = _lib.load()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', dev_index)
        from .weights import KIND_REID, read_kind
        kind = read_kind(self.engine_path)
        if max_batch is None:
            max_batch = 256 if kind == KIND_REID else 1
        handle = C.c_void_p()
        rc = self._lib.aicam_engine_create(str(self.engine_path).encode(), dev_index, int(max_batch), C.byref(handle))
        if rc != 0:
            raise RuntimeError("Failed to load weight blob %s: %s" % (
                self.engine_path, self._lib.aicam_last_error().decode(errors="replace")))
        self._h = handle
        self.kind = self._lib.aicam_engine_kind(self._h)
        self.max_batch = int(max_batch)
        self.topk, self.max_candidates = int(topk), int(max_candidates)
        self.score_threshold, self.nms_threshold = float(score_threshold), float(nms_threshold)
        if self.kind == _lib.KIND_YOLOV8:
            self.nc = self._lib.aicam_engine_num_classes(self._h)
            self.anchors = self._lib.aicam_engine_num_anchors(self._h)
            h, w = config.YOLO_INPUT_SHAPE
            self.input_info_list: List[TensorInfo] = self._bindings(0)
            self.output_info_list: List[TensorInfo] = self._bindings(1)
            self._nms = _lib.NmsParams(self.score_threshold, self.nms_threshold, self.topk, self.max_candidates, 0, 0)
            ws = self._lib.aicam_decode_nms_workspace(self.max_batch, self.anchors, C.byref(self._nms))
            self._ws = torch.empty(ws, dtype=torch.uint8, device=self.device)
            # the fp32 head tensor exists only when the engine cannot decode inside its Detect-head epilogues
            self.fused_decode = bool(self._lib.aicam_engine_fused_decode(self._h))
            self._head = None if self.fused_decode else torch.empty((self.max_batch, self.anchors, 64 + self.nc),
                                                                    dtype=torch.float32, device=self.device)
            self._nhwc = torch.empty((self.max_batch, h, w, 4), dtype=torch.bfloat16, device=self.device)
        else:
            self.feature_dim = self._lib.aicam_engine_num_classes(self._h)
            h, w = config.REID_INPUT_SHAPE
            self.input_info_list = self._bindings(0)
            self.output_info_list = self._bindings(1)
            self._nhwc = torch.empty((self.max_batch, h, w, 4), dtype=torch.bfloat16, device=self.device)

    def _bindings(self, is_output: int) -> List[TensorInfo]:
        """trt_engine.py:62-91 walks the engine's bindings; here the library reports them."""
        out = []
        for i in range(self._lib.aicam_engine_io_count(self._h, is_output)):
            ti = _lib.TensorInfo()
            _lib.check(self._lib.aicam_engine_io_info(self._h, is_output, i, self.topk, C.byref(ti)))
            dt = torch.float32 if ti.dtype == 0 else torch.int32
            out.append(TensorInfo(ti.name.decode(), dt, tuple(ti.shape[:ti.ndim]), bool(ti.is_dynamic)))
        return out

    def __del__(self):
        try:
            h = self.__dict__.get("_h")
            if h:
                self.__dict__["_h"] = None
                self.__dict__["_lib"].aicam_engine_destroy(h)
        except Exception:  # interpreter shutdown
            pass

    # -- raw entry points used by the batched pipeline ------------------------------------------
    @property
    def handle(self):
        return self._h

    def flops_per_item(self) -> float:
        return self._lib.aicam_engine_flops_per_item(self._h)

    def launches_per_forward(self) -> int:
        return self._lib.aicam_engine_num_launches(self._h)

    def get_bias(self, name: str) -> torch.Tensor:
        import numpy as np
        from .weights import read_blob
        n = read_blob(self.engine_path)[2][name + ".bias"].shape[0]
        out = np.empty(n, np.float32)
        _lib.check(self._lib.aicam_engine_get_bias(self._h, name.encode(), _lib.ptr(out), n))
        return torch.from_numpy(out)

    def set_bias(self, name: str, values) -> None:
        import numpy as np
        v = np.ascontiguousarray(np.asarray(values, dtype=np.float32))
        _lib.check(self._lib.aicam_engine_set_bias(self._h, name.encode(), _lib.ptr(v), v.shape[0]))

    @torch.no_grad()
    def infer(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        for info in self.input_info_list:
            t = inputs.get(info.name)
            if t is None:
                raise ValueError(f"Missing input: '{info.name}'")
            if t.device != self.device:
                t = t.to(self.device)
            if t.dtype != info.dtype:
                print(f"Warning: Input tensor '{info.name}' dtype mismatch. Expected {info.dtype}, got {t.dtype}. Casting...")
                t = t.to(info.dtype)
            inputs[info.name] = t.contiguous()
        x = inputs[self.input_info_list[0].name]
        n = x.shape[0]
        if x.dim() != 4 or x.shape[1] != 3 or tuple(x.shape[2:]) != tuple(self.input_info_list[0].shape[2:]):
            raise RuntimeError(f"execute failed for engine {self.engine_path.name}: bad input shape {tuple(x.shape)}")
        st = _lib.stream_ptr(self.device)
        outs: Dict[str, torch.Tensor] = {}
        with torch.cuda.device(self.device):
            if self.kind == _lib.KIND_YOLOV8:
                if n > self.max_batch:
                    raise RuntimeError(f"execute failed for engine {self.engine_path.name}: batch {n} > {self.max_batch}")
                _lib.check(self._lib.aicam_nchw_to_nhwc4(_lib.ptr(x), n, x.shape[2], x.shape[3], _lib.ptr(self._nhwc), st))
                outs['num_dets'] = torch.empty((n, 1), dtype=torch.int32, device=self.device)
                outs['bboxes'] = torch.empty((n, self.topk, 4), dtype=torch.float32, device=self.device)
                outs['scores'] = torch.empty((n, self.topk), dtype=torch.float32, device=self.device)
                outs['labels'] = torch.empty((n, self.topk), dtype=torch.int32, device=self.device)
                _lib.check(self._lib.aicam_yolo_detect(
                    self._h, _lib.ptr(self._nhwc), 0, n, C.byref(self._nms), _lib.ptr(self._head) if self._head is not None else None,
                    _lib.ptr(outs['num_dets']), _lib.ptr(outs['bboxes']), None, _lib.ptr(outs['scores']), _lib.ptr(outs['labels']),
                    _lib.ptr(self._ws), self._ws.numel(), st))
            else:
                outs['output'] = torch.empty((n, self.feature_dim), dtype=torch.float32, device=self.device)
                for s in range(0, n, self.max_batch):
                    nb = min(self.max_batch, n - s)
                    _lib.check(self._lib.aicam_nchw_to_nhwc4(_lib.ptr(x[s:s + nb]), nb, x.shape[2], x.shape[3],
                                                             _lib.ptr(self._nhwc), st))
                    _lib.check(self._lib.aicam_reid_forward(self._h, _lib.ptr(self._nhwc), nb, None,
                                                            _lib.ptr(outs['output'][s:s + nb]), st))
        stream = torch.cuda.current_stream(self.device)
        for t in outs.values():
            t.record_stream(stream)
        return outs

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if not isinstance(inputs, dict):
            raise TypeError(f"Input to {self.engine_path.name} TRTEngine must be a dictionary mapping "
                            "input names to torch.Tensors.")
        return self.infer(inputs)

    def get_input_details(self) -> List[TensorInfo]:
        return self.input_info_list

    def get_output_details(self) -> List[TensorInfo]:
        return self.output_info_list
