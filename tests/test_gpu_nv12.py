"""NV12 frame input (the surface format of hardware video decoders; SURVEY.md 8f N1, first half): every NV12
entry point equals the BGR entry point applied to the frame cv2 would have converted it to, bit for bit (-m gpu)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nv12(rng, n, h, w, smooth=True):
    """n NV12 frames [n, h*3/2, w]; smooth: low-frequency content so that bilinear taps differ meaningfully."""
    if not smooth:
        return rng.integers(0, 256, (n, h * 3 // 2, w), dtype=np.uint8)
    out = np.zeros((n, h * 3 // 2, w), np.uint8)
    for i in range(n):
        base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
        y = np.kron(base, np.ones((8, 8), np.float32))[:h, :w] * 0.7 + rng.normal(0, 20, (h, w)) + 30
        out[i, :h] = np.clip(y, 0, 255).astype(np.uint8)
        out[i, h:] = rng.integers(60, 200, (h // 2, w), dtype=np.uint8)
    return out


def _to_bgr(nv12_dev, h, w):
    import gpu_util as G
    n = nv12_dev.shape[0]
    bgr = torch.empty((n, h, w, 3), dtype=torch.uint8, device=G.DEV)
    G.check(G.lib().aicam_nv12_to_bgr(G.ptr(nv12_dev), n, h, w, G.ptr(bgr), None))
    G.sync()
    return bgr


@pytest.mark.parametrize("hw", [(1080, 1920), (540, 960), (720, 1280), (480, 640), (300, 500), (1000, 700), (2160, 3840)], ids=str)
def test_nv12_conversion_and_preprocess_bit_exact(hw):
    import gpu_util as G
    from oracle import image_ops
    h, w = hw
    rng = np.random.default_rng(h * 13 + w)
    nv = np.concatenate([_nv12(rng, 1, h, w), _nv12(rng, 1, h, w, smooth=False)])
    nd = torch.from_numpy(nv).to(G.DEV)
    bgr = _to_bgr(nd, h, w)
    for i in range(2):  # device conversion == oracle restatement (== cv2, tests/test_oracle_imageops.py)
        assert np.array_equal(bgr[i].cpu().numpy(), image_ops.nv12_to_bgr(nv[i], h, w)), "nv12_to_bgr differs"
    for fmt in (0, 1, 2, 3):
        want = G.preprocess(bgr, fmt)
        got = torch.empty_like(want)
        G.check(G.lib().aicam_preprocess_nv12(G.ptr(nd), 2, h, w, fmt, G.ptr(got), None))
        G.sync()
        assert torch.equal(got.view(torch.int16 if fmt else torch.int32), want.view(torch.int16 if fmt else torch.int32)), (hw, fmt)
    # and against the oracle end to end for the reference's tensor
    got0 = torch.empty((2, 3, 640, 640), dtype=torch.float32, device=G.DEV)
    G.check(G.lib().aicam_preprocess_nv12(G.ptr(nd), 2, h, w, 0, G.ptr(got0), None))
    G.sync()
    want0, _, _ = image_ops.preprocess_yolo_input(image_ops.nv12_to_bgr(nv[0], h, w))
    assert np.array_equal(got0[0].cpu().numpy().view(np.uint32), want0[0].view(np.uint32))


def test_nv12_rejects_odd_sizes():
    import gpu_util as G
    x = torch.zeros((1, 30, 11), dtype=torch.uint8, device=G.DEV)
    out = torch.zeros((1, 3, 640, 640), dtype=torch.float32, device=G.DEV)
    assert G.lib().aicam_preprocess_nv12(G.ptr(x), 1, 20, 11, 0, G.ptr(out), None) == -1


def test_nv12_crops_bit_exact():
    import gpu_util as G
    from ai_camera_b200.config import tracked_class_mask
    rng = np.random.default_rng(5)
    B, H, W, K = 2, 540, 960, 20
    nd = torch.from_numpy(_nv12(rng, B, H, W)).to(G.DEV)
    bgr = _to_bgr(nd, H, W)
    boxes = np.zeros((B, K, 4), np.float32)
    special = [(10.2, 20.7, 74.9, 148.99), (100, 100, 228, 356), (-30.5, -10.2, 40.3, 90.8), (900.4, 500.1, 990.0, 560.0),
               (300.9, 200.2, 301.1, 201.9), (5, 5, 700, 530), (400, 300, 464, 428), (401, 301, 465, 429), (33, 77, 161, 333)]
    for b in range(B):
        for k in range(K):
            if k < len(special):
                boxes[b, k] = special[k]
            else:
                x, y = rng.uniform(-20, W), rng.uniform(-20, H)
                boxes[b, k] = (x, y, x + rng.uniform(2, 300), y + rng.uniform(2, 400))
    bd = torch.from_numpy(boxes).to(G.DEV)
    sd = torch.full((B, K), 0.9, dtype=torch.float32, device=G.DEV)
    ld = torch.zeros((B, K), dtype=torch.int32, device=G.DEV)
    nm = torch.full((B,), K, dtype=torch.int32, device=G.DEV)
    lo, hi = tracked_class_mask()
    cap = B * K
    res = []
    for fn, fr in ((G.lib().aicam_reid_crops, bgr), (G.lib().aicam_reid_crops_nv12, nd)):
        outs = []
        for fmt, shape, dt in ((0, (cap, 3, 128, 64), torch.float32), (2, (cap, 128, 64, 8), torch.bfloat16)):
            i32 = dict(dtype=torch.int32, device=G.DEV)
            det_index, det_count, crop_slot = torch.zeros((B, K), **i32), torch.zeros(B, **i32), torch.zeros((B, K), **i32)
            crop_rect, crop_count = torch.zeros((cap, 5), **i32), torch.zeros(2, **i32)
            crops = torch.zeros(shape, dtype=dt, device=G.DEV)
            G.check(fn(G.ptr(fr), B, H, W, G.ptr(bd), G.ptr(sd), G.ptr(ld), G.ptr(nm), K, 0.3, lo, hi, fmt, cap,
                       G.ptr(det_index), G.ptr(det_count), G.ptr(crop_slot), G.ptr(crop_rect), G.ptr(crops), G.ptr(crop_count), None))
            G.sync()
            outs.append((crops.view(torch.int32 if fmt == 0 else torch.int16).cpu(), crop_slot.cpu(), crop_rect.cpu(), int(crop_count[0])))
        res.append(outs)
    for a, b in zip(res[0], res[1]):
        assert a[3] == b[3] and a[3] > 10
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[0], b[0])


def test_pipeline_nv12_equals_pipeline_bgr(tmp_path):
    """Whole path: NV12 frames in -> exactly the track tables of the BGR path on the converted frames."""
    import gpu_util as G
    from ai_camera_b200 import synth
    from ai_camera_b200.pipeline import TrackingPipeline
    yolo, reid = synth.make_blobs(str(tmp_path))
    video = synth.SynthVideo(3, (540, 960), n_frames=6, seed=99)
    nv = synth.bgr_to_nv12(video.ring.view(-1, 540, 960, 3)).view(6, 3, 810, 960)
    pipes = [TrackingPipeline(yolo, reid, 3, max_tracks=64, n_init=2) for _ in range(2)]
    bias = synth.shifted_class_bias(yolo, -0.6)
    for p in pipes:
        synth.apply_class_bias(p.detector.engine, bias)
    total = 0
    for t in range(6):
        bgr = _to_bgr(nv[t].contiguous(), 540, 960)
        a = [x.clone() for x in pipes[0].step(bgr)]
        b = [x.clone() for x in pipes[1].step(nv[t].contiguous())]
        G.sync()
        n = a[2].cpu().numpy()
        assert torch.equal(a[2], b[2])
        for s in range(3):
            assert torch.equal(a[0][s, :n[s]], b[0][s, :n[s]]) and torch.equal(a[1][s, :n[s]], b[1][s, :n[s]])
        total += int(n.sum())
    assert total > 0
