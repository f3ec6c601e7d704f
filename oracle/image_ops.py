"""Image operations of the hot path, restated in numpy (oracle; test infrastructure).

Follows ``/root/reference/src/utils/image_processing.py``:
  letterbox               :7-70
  preprocess_yolo_input   :73-102
  preprocess_reid_input   :105-138
  scale_bboxes            :141-183
and ``_extract_image_crops`` of ``src/tracker/deepsort_tracker.py:143-159``.

The reference resizes with ``cv2.resize(..., INTER_LINEAR)`` on uint8 images
(image_processing.py:64,123).  OpenCV is a third-party dependency that is not
vendored under /root/reference (requirements.txt pins opencv-python 4.11.0.86;
the build container holds 4.13.0).  ``resize_linear_u8`` restates OpenCV's
published fixed-point algorithm for that case (modules/imgproc/src/resize.cpp:
11-bit coefficients, horizontal pass in int32, vertical pass
``((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2``, and the exact-2x shortcut
to the 2x2 box average).  It is pinned bit for bit against cv2 itself through
``tests/golden/imageops.npz`` (recorded from the reference's functions) and,
where cv2 is importable, directly in ``tests/test_oracle_imageops.py``.
No function here calls cv2.
"""
import numpy as np

from .constants import IMAGENET_MEAN, IMAGENET_STD, LETTERBOX_PAD_VALUE

F32 = np.float32
COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def _axis_coeffs(src, dst):
    """Source index pair and 11-bit weights for one axis (OpenCV resize.cpp, linear branch).

    Returns (i0, i1, w0, w1): int32 arrays of length dst.  For the horizontal axis OpenCV
    forces frac = 0 at the borders; for the vertical axis it keeps the weights and clamps
    the row indices.  Both cases give the same result as clamping indices here, because a
    clamped pair has i0 == i1 and w0 + w1 == 2048 ... except that the two passes round
    separately, so the caller applies the horizontal rule (frac = 0) only for x."""
    scale = 1.0 / (float(dst) / float(src))  # double, as cv::resize computes it
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(F32)
    s = np.floor(f).astype(np.int32)
    frac = (f - s.astype(F32)).astype(F32)
    return s, frac


def _round_coef(x):
    # saturate_cast<short>(float) == cvRound == round half to even
    return np.rint(x.astype(F32) * F32(COEF_ONE)).astype(np.int32)


def resize_linear_u8(img, dst_w, dst_h):
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for HxWxC uint8."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    sh, sw = img.shape[:2]
    if (sh, sw) == (dst_h, dst_w):
        return img.copy()
    # exact 2x decimation on both axes is routed to the INTER_AREA fast path
    if sw == 2 * dst_w and sh == 2 * dst_h:
        a = img.astype(np.int32)
        s = a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]
        return ((s + 2) >> 2).astype(np.uint8)
    sx, fx = _axis_coeffs(sw, dst_w)
    sy, fy = _axis_coeffs(sh, dst_h)
    # horizontal: border handling zeroes the fraction
    lo = sx < 0
    hi = sx >= sw - 1
    fx = np.where(lo | hi, F32(0), fx)
    sx = np.where(lo, 0, np.where(hi, sw - 1, sx))
    ax1 = _round_coef(fx)
    ax0 = _round_coef(F32(1.0) - fx)
    sx1 = np.minimum(sx + 1, sw - 1)
    # vertical: weights kept, rows clamped
    by1 = _round_coef(fy)
    by0 = _round_coef(F32(1.0) - fy)
    sy0 = np.clip(sy, 0, sh - 1)
    sy1 = np.clip(sy + 1, 0, sh - 1)
    src = img.astype(np.int32)
    # horizontal pass on the rows that are needed
    rows = np.unique(np.concatenate([sy0, sy1]))
    hbuf = {}
    a0 = ax0[None, :, None]
    a1 = ax1[None, :, None]
    hr = src[rows][:, sx, :] * a0 + src[rows][:, sx1, :] * a1  # (nrows, dst_w, C) int32
    pos = {int(r): k for k, r in enumerate(rows)}
    i0 = np.asarray([pos[int(r)] for r in sy0])
    i1 = np.asarray([pos[int(r)] for r in sy1])
    r0 = hr[i0]
    r1 = hr[i1]
    b0 = by0[:, None, None]
    b1 = by1[:, None, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_params(h, w, new_shape=(640, 640)):
    """Geometry of ``letterbox(auto=False, scaleup=False)`` (image_processing.py:32-67).

    Returns dict(r, new_h, new_w, dw, dh, top, bottom, left, right); r/dw/dh are python floats."""
    r = min(min(new_shape[0] / h, 1.0), min(new_shape[1] / w, 1.0))
    new_h, new_w = int(round(h * r)), int(round(w * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return dict(r=r, new_h=new_h, new_w=new_w, dw=dw, dh=dh, top=top, bottom=bottom,
                left=left, right=right)


def letterbox_u8(img_bgr, new_shape=(640, 640)):
    """image_processing.py:7-70 with auto=False, scaleup=False -> (uint8 HxWx3 BGR, (r,r), (dw,dh)).

    The resize always executes (the guard at :63 compares (W,H) with (H,W)); when the
    sizes already agree cv2.resize is an identity copy."""
    h, w = img_bgr.shape[:2]
    p = letterbox_params(h, w, new_shape)
    im = resize_linear_u8(img_bgr, p["new_w"], p["new_h"])
    out = np.full((p["top"] + p["new_h"] + p["bottom"], p["left"] + p["new_w"] + p["right"], 3),
                  LETTERBOX_PAD_VALUE, np.uint8)
    out[p["top"]:p["top"] + p["new_h"], p["left"]:p["left"] + p["new_w"]] = im
    return out, (p["r"], p["r"]), (p["dw"], p["dh"])


def preprocess_yolo_input(img_bgr, target_shape=(640, 640)):
    """image_processing.py:73-102 -> ((1,3,H,W) float32 RGB in [0,1], ratios, (pad_w, pad_h))."""
    lb, ratios, pad = letterbox_u8(img_bgr, target_shape)
    rgb = lb[:, :, ::-1]
    chw = np.transpose(rgb, (2, 0, 1))
    t = np.expand_dims(chw, 0).astype(F32) / 255.0
    return np.ascontiguousarray(t), ratios, pad


def preprocess_reid_input(crop_bgr, target_shape=(128, 64)):
    """image_processing.py:105-138 -> (1,3,128,64) float32."""
    r = resize_linear_u8(crop_bgr, target_shape[1], target_shape[0])
    rgb = r[:, :, ::-1]
    mean = np.array(IMAGENET_MEAN, dtype=F32)
    std = np.array(IMAGENET_STD, dtype=F32)
    n = (rgb.astype(F32) / 255.0 - mean) / std
    return np.ascontiguousarray(np.expand_dims(np.transpose(n, (2, 0, 1)), 0), dtype=F32)


def scale_bboxes(b, original_shape, ratio, padding):
    """image_processing.py:141-183 (float32 array, python-float pad/ratio are weak scalars)."""
    b = np.asarray(b)
    if b.size == 0:
        return np.empty((0, 4), dtype=F32)
    s = b.copy()
    pad_w, pad_h = padding
    ratio_h, ratio_w = ratio
    s[:, 0] -= pad_w
    s[:, 1] -= pad_h
    s[:, 2] -= pad_w
    s[:, 3] -= pad_h
    s[:, 0] /= ratio_w
    s[:, 1] /= ratio_h
    s[:, 2] /= ratio_w
    s[:, 3] /= ratio_h
    oh, ow = original_shape
    s[:, [0, 2]] = np.clip(s[:, [0, 2]], 0, ow)
    s[:, [1, 3]] = np.clip(s[:, [1, 3]], 0, oh)
    return s


def reid_batch(frame_bgr, rects, target_shape=(128, 64)):
    """Crops (x1,y1,x2,y2 clamped ints) -> (N,3,128,64) float32, as reid_model.py:83-100 builds it."""
    if not rects:
        return np.empty((0, 3, target_shape[0], target_shape[1]), F32)
    return np.concatenate([preprocess_reid_input(frame_bgr[y1:y2, x1:x2], target_shape)
                           for (x1, y1, x2, y2) in rects], axis=0)


def nv12_to_bgr(nv12, h, w):
    """cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12) restated (OpenCV imgproc color_yuv: ITU-R BT.601 limited
    range, 20-bit fixed point: ITUR_BT_601_CY 1220542, CUB 2116026, CUG -409993, CVG -852492, CVR 1673527,
    rounding 1 << 19).  nv12: uint8 (h*3/2, w): the Y plane, then rows of interleaved U, V at half
    resolution.  This is the conversion the video decoder under ``cv2.VideoCapture.read()``
    (/root/reference/src/aicamera_tracker.py:170) applies before the reference ever sees a frame; the
    device path can take the NV12 surface itself (aicam_preprocess_nv12 / aicam_reid_crops_nv12).
    Verified equal to cv2 4.13 in tests/test_oracle_imageops.py."""
    nv12 = np.asarray(nv12, np.uint8).reshape(h * 3 // 2, w)
    Y = nv12[:h].astype(np.int32)
    uv = nv12[h:].reshape(h // 2, w // 2, 2).astype(np.int32)
    U = np.repeat(np.repeat(uv[..., 0], 2, 0), 2, 1) - 128
    V = np.repeat(np.repeat(uv[..., 1], 2, 0), 2, 1) - 128
    yy = np.maximum(0, Y - 16) * 1220542 + (1 << 19)
    b = (yy + 2116026 * U) >> 20
    g = (yy - 852492 * V - 409993 * U) >> 20
    r = (yy + 1673527 * V) >> 20
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)
