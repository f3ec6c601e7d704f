"""The oracle's image ops are pinned to what the reference's src/utils/image_processing.py
(cv2 4.13) produced on seeded images (tests/golden/imageops.npz) - bit for bit."""
import numpy as np
import pytest

from golden_util import load, sha
from scenarios import IMAGEOPS_SEED, LETTERBOX_SIZES, REID_CROP_SIZES, synth_image

from oracle import image_ops


def _images():
    rng = np.random.default_rng(IMAGEOPS_SEED)
    lb = [synth_image(rng, h, w) for (h, w) in LETTERBOX_SIZES]
    crops = [np.ascontiguousarray(synth_image(rng, max(h, 16), max(w, 16))[:h, :w])
             for (h, w) in REID_CROP_SIZES]
    boxes = rng.uniform(-20, 660, (64, 4)).astype(np.float32)
    return lb, crops, boxes


def test_preprocess_yolo_input_matches_reference():
    g = load("imageops.npz")
    lb, _, _ = _images()
    for k, img in enumerate(lb):
        assert sha(img) == bytes(g["yolo%d_seed_digest" % k]).decode()
        t, ratios, pad = image_ops.preprocess_yolo_input(img, (640, 640))
        assert t.dtype == np.float32 and t.shape == (1, 3, 640, 640)
        assert np.array_equal(np.asarray([ratios[0], ratios[1], pad[0], pad[1]]), g["yolo%d_meta" % k])
        assert sha(t) == bytes(g["yolo%d_digest" % k]).decode(), LETTERBOX_SIZES[k]


def test_preprocess_reid_input_matches_reference():
    g = load("imageops.npz")
    _, crops, _ = _images()
    for k, img in enumerate(crops):
        assert sha(img) == bytes(g["reid%d_seed_digest" % k]).decode()
        t = image_ops.preprocess_reid_input(img, (128, 64))
        assert np.array_equal(t[0, :, ::13, ::7], g["reid%d_sample" % k])
        assert sha(t) == bytes(g["reid%d_digest" % k]).decode(), REID_CROP_SIZES[k]


def test_scale_bboxes_matches_reference():
    g = load("imageops.npz")
    _, _, boxes = _images()
    assert np.array_equal(boxes, g["scale_in"])
    for k, (h, w) in enumerate(LETTERBOX_SIZES):
        p = image_ops.letterbox_params(h, w)
        got = image_ops.scale_bboxes(boxes, (h, w), (p["r"], p["r"]), (p["dw"], p["dh"]))
        assert got.dtype == np.float32
        assert np.array_equal(got.view(np.uint32), g["scale%d" % k].view(np.uint32))


def test_1080p_is_an_exact_gather():
    """for the exact 3:1 case the bilinear resize picks source pixel 3i+1."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    assert np.array_equal(image_ops.resize_linear_u8(img, 640, 360), img[1::3, 1::3])


def test_resize_against_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(25):
        sh, sw = (int(v) for v in rng.integers(1, 300, 2))
        dw, dh = (int(v) for v in rng.integers(1, 200, 2))
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(image_ops.resize_linear_u8(img, dw, dh), want), (sh, sw, dw, dh)


def test_nv12_to_bgr_equals_cv2():
    """The oracle's NV12 -> BGR restatement against OpenCV itself (the conversion under VideoCapture.read())."""
    import cv2
    from oracle import image_ops
    rng = np.random.default_rng(11)
    for (h, w) in [(36, 64), (540, 960), (1080, 1920), (2, 2)]:
        nv12 = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
        assert np.array_equal(image_ops.nv12_to_bgr(nv12, h, w), cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12)), (h, w)
    # every (Y, U, V) triple the 8-bit domain has along the extremes
    ys, us, vs = np.meshgrid(np.arange(256), [0, 1, 16, 127, 128, 129, 240, 255], [0, 1, 16, 127, 128, 129, 240, 255], indexing="ij")
    n = ys.size
    img = np.zeros((3, 2 * n), np.uint8)  # 2 rows of luma (same Y twice), 1 row of UV pairs
    img[0, 0::2] = img[0, 1::2] = img[1, 0::2] = img[1, 1::2] = ys.ravel()
    img[2, 0::2], img[2, 1::2] = us.ravel(), vs.ravel()
    assert np.array_equal(image_ops.nv12_to_bgr(img, 2, 2 * n), cv2.cvtColor(img, cv2.COLOR_YUV2BGR_NV12))
