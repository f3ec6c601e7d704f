"""Overlay of tracks / detections / status text on frames that stay in device memory.

Same functions and argument meaning as the reference's ``src/utils/visualization.py`` (``draw_detections`` :9-69,
``draw_tracks`` :72-124, ``draw_fps`` :127-167, ``draw_info_panel`` :170-227), which runs ``cv2.rectangle`` /
``cv2.putText`` on a host copy of every frame.  Here a frame is a CUDA tensor (uint8 ``[H, W, 3]`` BGR, or a batch
``[S, H, W, 3]``) and everything is drawn by ONE kernel launch per call (``aicam_overlay_draw``, csrc/overlay.cu):

* box outlines and filled rectangles are rasterised on the device, pixel for pixel what ``cv2.rectangle`` fills;
* anti-aliased text is a *decal*: the text (with the filled background behind it) is drawn once by ``cv2`` itself on a
  black and on a white canvas, which gives per pixel the affine map ``out = B + in * A / 255`` that reproduces both; the
  decal of a label such as ``"ID:17 person"`` is cached in a device atlas, so steady state uploads nothing but the item
  list.  Inside the label background the decal is exactly cv2's output; on the few anti-aliased pixels that spill outside
  it the affine map is within a grey level or two of cv2's integer blending (asserted in tests/test_gpu_overlay.py).

numpy frames are accepted as well (uploaded, drawn, downloaded): the module is then a drop-in for the reference's, with
the drawing done on the GPU.  There is no CPU drawing path.
"""
import ctypes as C
from collections import OrderedDict
from typing import List, Sequence

import numpy as np
import torch

from . import _lib, config

LINE_AA = 16  # cv2.LINE_AA
_MARGIN = 4   # canvas pixels kept around a label for anti-aliased fringes


def _bgr_int(color):
    b, g, r = (int(c) & 255 for c in color[:3])
    return b | (g << 8) | (r << 16)


def _affine_decal(draw, h, w):
    """Run ``draw(canvas)`` on a black and on a white canvas: uint8 [h, w, 8] = B (3), 0, A (3), 0 per pixel and the
    (width, height) of the part that differs from pass-through (B = 0, A = 255)."""
    import cv2  # noqa: F401  (the drawing callbacks use it)
    black = np.zeros((h, w, 3), np.uint8)
    white = np.full((h, w, 3), 255, np.uint8)
    draw(black)
    draw(white)
    out = np.zeros((h, w, 8), np.uint8)
    out[:, :, 0:3] = black
    # a later opaque draw can only make white <= 255 and >= black: A = white - black is in 0..255
    out[:, :, 4:7] = white.astype(np.int16) - black.astype(np.int16)
    touched = (out[:, :, 0:3] != 0).any(-1) | (out[:, :, 4:7] != 255).any(-1)
    if not touched.any():
        return out, (0, 0)
    ys, xs = np.nonzero(touched)
    return out, (int(xs.max()) + 1, int(ys.max()) + 1)


class Overlay:
    """Device-side overlay state of one GPU: the decal atlas (label cache) and the item upload buffers."""

    def __init__(self, device="cuda:0", slots=256, slot_w=448, slot_h=48, panel_w=1024, panel_h=192):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.slots, self.slot_w, self.slot_h = int(slots), int(slot_w), int(slot_h)
        self.atlas = torch.zeros((self.slots, self.slot_h, self.slot_w, 8), dtype=torch.uint8, device=self.device)
        # the status panel / fps box change every frame: their decal has an atlas of its own, one slot per frame of a batch
        self.panel_w, self.panel_h = int(panel_w), int(panel_h)
        self.panels = None
        self._cache = OrderedDict()  # key -> (slot, w, h)
        self._free = list(range(self.slots - 1, -1, -1))

    # ---- decals ------------------------------------------------------------------------------------------------------
    def _label(self, text, color, scale, pad_above, text_dy, x1, y1, W, H):
        """Decal of a label box as the reference draws it above a box corner (x1, y1): filled background
        (x1, y1 - th - baseline - pad_above) .. (x1 + tw, y1) in `color`, white text at (x1, y1 - baseline // 2 - text_dy)
        (visualization.py:45-67 with pad_above = 0, text_dy = 0; :101-123 with 2 and 1).  Returns the decal item, or None
        when the label lies outside the W x H frame.

        OpenCV clips anti-aliased shapes at the image border before rasterising them, so a label cut by a frame edge is
        not a shifted copy of the unclipped one: such labels are rendered on a canvas whose edge coincides with that frame
        edge (and cached under that placement)."""
        import cv2
        (tw, th), base = cv2.getTextSize(text, config.FONT, scale, config.FONT_THICKNESS)
        top = th + base + pad_above + _MARGIN  # extent above the anchor row, fringe included
        ax, ay = _MARGIN, self.slot_h - 2 * _MARGIN
        if x1 - _MARGIN < 0:
            ax = x1                              # canvas column 0 = frame column 0
        elif x1 + tw + _MARGIN > W:
            ax = self.slot_w - (W - x1)          # canvas right edge = frame right edge
        if y1 - top < 0:
            ay = y1                              # canvas row 0 = frame row 0
        elif y1 + _MARGIN > H:
            ay = self.slot_h - (H - y1)          # canvas bottom edge = frame bottom edge
        # nothing of the label inside the frame (or the placement does not fit the canvas: drawn unclipped-style)
        if x1 - ax >= W or y1 - ay >= H or x1 - ax + self.slot_w <= 0 or y1 - ay + self.slot_h <= 0:
            return None
        key = (text, tuple(int(c) for c in color[:3]), scale, pad_above, text_dy, ax, ay)
        hit = self._cache.get(key)
        if hit is not None:
            self._cache.move_to_end(key)
            return (2, x1 - ax, y1 - ay, hit[1], hit[2], 0, hit[0])
        col = tuple(int(c) for c in color[:3])

        def draw(img):
            cv2.rectangle(img, (ax, ay - th - base - pad_above), (ax + tw, ay), col, -1)
            cv2.putText(img, text, (ax, ay - base // 2 - text_dy), config.FONT, scale, (255, 255, 255), config.FONT_THICKNESS,
                        cv2.LINE_AA)

        decal, (w, h) = _affine_decal(draw, self.slot_h, self.slot_w)
        if not self._free:  # evict the least recently used label
            _, (old_slot, _, _) = self._cache.popitem(last=False)
            self._free.append(old_slot)
        slot = self._free.pop()
        self.atlas[slot].copy_(torch.from_numpy(decal), non_blocking=False)
        self._cache[key] = (slot, w, h)
        return (2, x1 - ax, y1 - ay, w, h, 0, slot)

    # ---- item lists --------------------------------------------------------------------------------------------------
    def track_items(self, tracked_objects: Sequence, frame_hw) -> List[tuple]:
        """Items of ``draw_tracks`` for one frame of frame_hw = (H, W): tuples (x1, y1, x2, y2, track_id, class_name[, score])."""
        items = []
        H, W = int(frame_hw[0]), int(frame_hw[1])
        for obj in tracked_objects:
            x1, y1, x2, y2 = (int(v) for v in obj[:4])
            color = config.get_track_color(obj[5])
            label = "ID:%s %s" % (obj[4], obj[5])
            if len(obj) > 6:
                label += " %.2f" % obj[6]
            items.append((0, x1, y1, x2, y2, _bgr_int(color), 0))
            decal = self._label(label, color, config.FONT_SCALE_ID, 2, 1, x1, y1, W, H)
            if decal is not None:
                items.append(decal)
        return items

    def detection_items(self, bboxes_xyxy, scores, class_ids, class_names, frame_hw) -> List[tuple]:
        """Items of ``draw_detections`` for one frame of frame_hw = (H, W)."""
        items = []
        H, W = int(frame_hw[0]), int(frame_hw[1])
        for i in range(len(bboxes_xyxy)):
            x1, y1, x2, y2 = (int(v) for v in bboxes_xyxy[i])
            cid = int(class_ids[i])
            if cid < 0 or cid >= len(class_names):
                name, color = "Unknown", (128, 128, 128)
            else:
                name = class_names[cid]
                color = config.get_class_color(name)
            items.append((0, x1, y1, x2, y2, _bgr_int(color), 0))
            decal = self._label("%s: %.2f" % (name, float(scores[i])), color, config.FONT_SCALE_ID, 0, 0, x1, y1, W, H)
            if decal is not None:
                items.append(decal)
        return items

    def _panel_decal(self, draw):
        decal, (w, h) = _affine_decal(draw, self.panel_h, self.panel_w)
        return decal, w, h

    @staticmethod
    def _draw_info_panel(img, info_lines):
        """The reference's panel geometry (visualization.py:180-227) on `img` (the same arithmetic, cv2 does the pixels)."""
        import cv2
        x0, y0 = 10, 30
        sizes = [cv2.getTextSize(t, config.FONT, config.FONT_SCALE_INFO, config.FONT_THICKNESS) for t in info_lines]
        if not info_lines:
            return
        step = sizes[0][0][1] + sizes[0][1] + 10
        widest = max(s[0][0] for s in sizes)
        cv2.rectangle(img, (x0 - 5, y0 - step + 15), (x0 + widest + 5, y0 + len(info_lines) * step - step + 15), (50, 50, 50), -1)
        y = y0
        for text, ((_, th), base) in zip(info_lines, sizes):
            cv2.putText(img, text, (x0, y + base + th // 2), config.FONT, config.FONT_SCALE_INFO, (255, 255, 255), config.FONT_THICKNESS,
                        cv2.LINE_AA)
            y += step

    @staticmethod
    def _draw_fps(img, fps):
        import cv2
        text = "FPS: %.2f" % fps
        (tw, th), base = cv2.getTextSize(text, config.FONT, config.FONT_SCALE_INFO, config.FONT_THICKNESS)
        tx, ty = 10, th + 10 + base // 2
        cv2.rectangle(img, (tx - 5, ty - th - base - 5), (tx + tw + 5, ty + 5), (50, 50, 50), -1)
        cv2.putText(img, text, (tx, ty - base // 2), config.FONT, config.FONT_SCALE_INFO, (255, 255, 255), config.FONT_THICKNESS, cv2.LINE_AA)

    # ---- drawing -----------------------------------------------------------------------------------------------------
    def draw(self, frames: torch.Tensor, items_per_frame: Sequence[Sequence[tuple]], panels: Sequence = None, stream=None):
        """Draw every frame's items in order, in place.  frames: uint8 CUDA [S, H, W, 3] (or [H, W, 3]).  panels: per frame an
        optional callable ``draw(img)`` (status panel / fps box, rendered on the host into a per-frame decal and drawn LAST)."""
        if frames.dim() == 3:
            frames = frames.unsqueeze(0)
        if frames.dtype != torch.uint8 or not frames.is_cuda or not frames.is_contiguous() or frames.shape[-1] != 3:
            raise _lib.AicamError(-1, "overlay: frames must be a contiguous uint8 CUDA tensor [S, H, W, 3]")
        S, H, W = int(frames.shape[0]), int(frames.shape[1]), int(frames.shape[2])
        if len(items_per_frame) != S:
            raise _lib.AicamError(-1, "overlay: one item list per frame expected")
        st = stream if stream is not None else _lib.stream_ptr(frames.device)
        flat, start = [], [0]
        for it in items_per_frame:
            flat.extend(it)
            start.append(len(flat))
        if flat:
            self._launch(frames, S, H, W, flat, start, self.atlas, self.slot_w, self.slot_h, st)
        if panels is not None and any(p is not None for p in panels):
            if self.panels is None or self.panels.shape[0] < S:
                self.panels = torch.zeros((S, self.panel_h, self.panel_w, 8), dtype=torch.uint8, device=self.device)
            flat, start = [], [0]
            for n, p in enumerate(panels):
                if p is not None:
                    decal, w, h = self._panel_decal(p)
                    self.panels[n].copy_(torch.from_numpy(decal))
                    flat.append((2, 0, 0, w, h, 0, n))
                start.append(len(flat))
            self._launch(frames, S, H, W, flat, start, self.panels, self.panel_w, self.panel_h, st)
        return frames

    def _launch(self, frames, S, H, W, flat, start, atlas, slot_w, slot_h, st):
        arr = np.zeros((len(flat), 8), np.int32)
        arr[:, :7] = np.asarray(flat, np.int64).astype(np.int32)
        items = torch.from_numpy(arr).to(self.device)
        starts = torch.tensor(start, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.aicam_overlay_draw(_lib.ptr(frames), S, H, W, _lib.ptr(items), _lib.ptr(starts), _lib.ptr(atlas),
                                                   slot_w, slot_h, st))
        # (items / starts are released after the launch; same-stream ordering keeps them alive for the kernel)


_default = {}


def _overlay_for(device):
    key = str(device)
    if key not in _default:
        _default[key] = Overlay(device)
    return _default[key]


def _on_device(frame):
    """(device batch of one, was_numpy)."""
    if isinstance(frame, np.ndarray):
        if not torch.cuda.is_available():
            raise RuntimeError("visualization: drawing runs on the GPU (libaicam.so); there is no CPU path")
        return torch.from_numpy(np.ascontiguousarray(frame)).to("cuda:0").unsqueeze(0), True
    return (frame if frame.dim() == 4 else frame.unsqueeze(0)), False


def _back(frame, dev, was_numpy):
    if was_numpy:
        frame[...] = dev[0].cpu().numpy()  # the reference draws in place and returns the frame
        return frame
    return frame


def draw_tracks(frame, tracked_objects: list):
    """visualization.py:72-124: boxes with "ID:<id> <class>[ <score>]" labels.  tracked_objects: tuples
    (x1, y1, x2, y2, track_id, class_name[, score])."""
    dev, was_numpy = _on_device(frame)
    ov = _overlay_for(dev.device)
    ov.draw(dev, [ov.track_items(tracked_objects, dev.shape[1:3])])
    return _back(frame, dev, was_numpy)


def draw_detections(frame, bboxes_xyxy, scores, class_ids, class_names: tuple):
    """visualization.py:9-69: raw detections with "<class>: <score>" labels."""
    dev, was_numpy = _on_device(frame)
    ov = _overlay_for(dev.device)
    ov.draw(dev, [ov.detection_items(bboxes_xyxy, scores, class_ids, class_names, dev.shape[1:3])])
    return _back(frame, dev, was_numpy)


def draw_fps(frame, fps: float):
    """visualization.py:127-167."""
    dev, was_numpy = _on_device(frame)
    ov = _overlay_for(dev.device)
    ov.draw(dev, [[]], panels=[lambda img: Overlay._draw_fps(img, fps)])
    return _back(frame, dev, was_numpy)


def draw_info_panel(frame, info_lines: List[str]):
    """visualization.py:170-227."""
    dev, was_numpy = _on_device(frame)
    ov = _overlay_for(dev.device)
    lines = list(info_lines)
    ov.draw(dev, [[]], panels=[(lambda img: Overlay._draw_info_panel(img, lines)) if lines else None])
    return _back(frame, dev, was_numpy)
