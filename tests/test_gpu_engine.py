"""YOLOv8 / ReID engines (bf16 tcgen05 convolutions) vs the PyTorch-CPU fp32 oracle nets (-m gpu).

Tolerances are the ones BASELINE.json's north_star states: boxes within 1e-2 relative after
NMS, embeddings cosine >= 0.999."""
import ctypes as C

import numpy as np
import pytest
import torch

from scenarios import synth_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def blobs(tmp_path_factory):
    from ai_camera_b200 import weights as W
    d = tmp_path_factory.mktemp("blobs")
    paths = {}
    for name, gen in (("yolov8n", lambda: W.synth_yolov8_weights("n", seed=0)),
                      ("yolov8s", lambda: W.synth_yolov8_weights("s", seed=0)),
                      ("yolov8m", lambda: W.synth_yolov8_weights("m", seed=0)),
                      ("reid", lambda: W.synth_reid_weights(seed=1))):
        kind, params, tensors = gen()
        paths[name] = str(d / (name + ".aicw"))
        W.write_blob(paths[name], kind, params, tensors)
    return paths


def _engine(path, max_batch):
    import gpu_util as G
    h = C.c_void_p()
    G.check(G.lib().aicam_engine_create(path.encode(), 0, max_batch, C.byref(h)))
    return h


def test_yolov8n_head_matches_oracle(blobs):
    import gpu_util as G
    from oracle import image_ops, nets
    rng = np.random.default_rng(0)
    frames = np.stack([synth_image(rng, 540, 960), rng.integers(0, 256, (540, 960, 3), dtype=np.uint8)])
    x = G.preprocess(torch.from_numpy(frames).to(G.DEV), 1)
    e = _engine(blobs["yolov8n"], 2)
    try:
        assert abs(G.lib().aicam_engine_flops_per_item(e) / 1e9 - 8.74) < 0.1
        head = torch.empty((2, 8400, 144), dtype=torch.float32, device=G.DEV)
        G.check(G.lib().aicam_yolo_forward(e, G.ptr(x), 2, G.ptr(head), None))
        G.sync()
    finally:
        G.lib().aicam_engine_destroy(e)
    net = nets.load_net(blobs["yolov8n"])
    xin = np.concatenate([image_ops.preprocess_yolo_input(f)[0] for f in frames])
    want = net.head_flat(torch.from_numpy(xin)).numpy()
    got = head.cpu().numpy()
    err = np.abs(got - want)
    scale = np.abs(want).max()
    print("head: max abs err %.4f (logit range %.2f), mean abs err %.5f" % (err.max(), scale, err.mean()))
    assert err.mean() < 0.02 and err.max() < 0.35
    # boxes after NMS within 1e-2 relative (of the box size) for detections both sides keep
    from oracle import detect_post
    for b in range(2):
        gb, gs, gl = detect_post.decode(got[b])
        wb, ws, wl = detect_post.decode(want[b])
        top = np.argsort(-ws)[:200]
        size = np.maximum(wb[top, 2] - wb[top, 0], wb[top, 3] - wb[top, 1])[:, None]
        assert (np.abs(gb[top] - wb[top]) / size).max() < 1e-2
        assert np.abs(gs[top] - ws[top]).max() < 2e-2


@pytest.mark.parametrize("scale,gflop", [("s", 28.6), ("m", 78.9)])
def test_yolov8_s_m_head_matches_oracle(blobs, scale, gflop):
    """BASELINE.json configs[3]: the larger YOLOv8 scales run on the same kernels (detection only)."""
    import gpu_util as G
    from oracle import detect_post, image_ops, nets
    rng = np.random.default_rng(3)
    frames = synth_image(rng, 540, 960)[None]
    x = G.preprocess(torch.from_numpy(frames).to(G.DEV), 1)
    path = blobs["yolov8" + scale]
    e = _engine(path, 1)
    try:
        assert abs(G.lib().aicam_engine_flops_per_item(e) / 1e9 - gflop) < 0.02 * gflop
        head = torch.empty((1, 8400, 144), dtype=torch.float32, device=G.DEV)
        G.check(G.lib().aicam_yolo_forward(e, G.ptr(x), 1, G.ptr(head), None))
        G.sync()
    finally:
        G.lib().aicam_engine_destroy(e)
    net = nets.load_net(path)
    xin = image_ops.preprocess_yolo_input(frames[0])[0]
    want = net.head_flat(torch.from_numpy(xin)).numpy()
    got = head.cpu().numpy()
    err = np.abs(got - want)
    print("yolov8%s head: max abs err %.4f (logit range %.2f), mean abs err %.5f" % (scale, err.max(), np.abs(want).max(), err.mean()))
    assert err.mean() < 0.03 and err.max() < 0.5
    gb, gs, gl = detect_post.decode(got[0])
    wb, ws, wl = detect_post.decode(want[0])
    top = np.argsort(-ws)[:200]
    size = np.maximum(wb[top, 2] - wb[top, 0], wb[top, 3] - wb[top, 1])[:, None]
    assert (np.abs(gb[top] - wb[top]) / size).max() < 1e-2
    assert np.abs(gs[top] - ws[top]).max() < 3e-2


@pytest.mark.parametrize("name,batch", [("yolov8n", 3), ("yolov8s", 1)])
def test_fused_decode_equals_decode_kernel(blobs, name, batch):
    """aicam_yolo_detect decodes inside the Detect-head epilogues (no fp32 head tensor); the dense per-anchor arrays it
    hands to the NMS kernel and the final detections must be bit-identical with forward -> decode_kernel -> NMS."""
    import gpu_util as G
    rng = np.random.default_rng(7)
    frames = np.stack([synth_image(rng, 540, 960) for _ in range(batch)])
    frames[-1] = rng.integers(0, 256, (540, 960, 3), dtype=np.uint8)
    x = G.preprocess(torch.from_numpy(frames).to(G.DEV), 1)
    lib = G.lib()
    e = _engine(blobs[name], batch)
    try:
        assert lib.aicam_engine_fused_decode(e) == 1
        p = G.NmsParams(0.05, 0.5, 100, 1024, 540, 960)
        nws = lib.aicam_decode_nms_workspace(batch, 8400, C.byref(p))

        def outs():
            return (torch.zeros(batch, dtype=torch.int32, device=G.DEV), torch.zeros((batch, 100, 4), device=G.DEV),
                    torch.zeros((batch, 100, 4), device=G.DEV), torch.zeros((batch, 100), device=G.DEV),
                    torch.zeros((batch, 100), dtype=torch.int32, device=G.DEV))
        head = torch.empty((batch, 8400, 144), dtype=torch.float32, device=G.DEV)
        G.check(lib.aicam_yolo_forward(e, G.ptr(x), batch, G.ptr(head), None))
        db, ds, dl = G.decode(head)
        ws1 = torch.zeros(nws, dtype=torch.uint8, device=G.DEV)
        o1 = outs()
        G.check(lib.aicam_decode_nms(G.ptr(head), batch, 8400, 80, C.byref(p), *[G.ptr(t) for t in o1], G.ptr(ws1), nws, None))
        ws2 = torch.zeros(nws, dtype=torch.uint8, device=G.DEV)
        o2 = outs()
        G.check(lib.aicam_yolo_detect(e, G.ptr(x), 0, batch, C.byref(p), None, *[G.ptr(t) for t in o2], G.ptr(ws2), nws, None))
        G.sync()
    finally:
        lib.aicam_engine_destroy(e)
    n = batch * 8400
    fb = ws2[:n * 16].view(torch.float32).reshape(batch, 8400, 4)
    fs = ws2[n * 16:n * 20].view(torch.float32).reshape(batch, 8400)
    fl = ws2[n * 20:n * 24].view(torch.int32).reshape(batch, 8400)
    assert torch.equal(fl, dl), "labels differ"
    assert torch.equal(fs, ds), "scores differ: max %g" % (fs - ds).abs().max().item()
    assert torch.equal(fb, db), "boxes differ: max %g" % (fb - db).abs().max().item()
    assert int(o1[0].sum()) > 0
    for a, b in zip(o1, o2):
        assert torch.equal(a, b)


def test_4x4_block_stem_matches_2x2_block_stem(blobs):
    """yolov8n runs its stem over 4x4 pixel blocks (aicam_preprocess format 3, in_is_s2d = 2): the same products summed in
    another order.  The stem's bf16 output may differ in the last bit; the detections must agree with the format-2 path
    far inside the 1e-2 box bar, and with the oracle exactly as the format-2 path does (test_yolov8n_head_matches_oracle)."""
    import gpu_util as G
    rng = np.random.default_rng(11)
    frames = np.stack([synth_image(rng, 540, 960) for _ in range(2)])
    fd = torch.from_numpy(frames).to(G.DEV)
    lib = G.lib()
    e = _engine(blobs["yolov8n"], 2)
    try:
        assert lib.aicam_engine_accepts_s2d(e) == 2
        p = G.NmsParams(0.05, 0.5, 100, 1024, 540, 960)
        nws = lib.aicam_decode_nms_workspace(2, 8400, C.byref(p))
        res = []
        for s2d in (1, 2):
            x = G.preprocess(fd, 1 + s2d)
            ws = torch.zeros(nws, dtype=torch.uint8, device=G.DEV)
            o = (torch.zeros(2, dtype=torch.int32, device=G.DEV), torch.zeros((2, 100, 4), device=G.DEV), torch.zeros((2, 100, 4), device=G.DEV),
                 torch.zeros((2, 100), device=G.DEV), torch.zeros((2, 100), dtype=torch.int32, device=G.DEV))
            G.check(lib.aicam_yolo_detect(e, G.ptr(x), s2d, 2, C.byref(p), None, *[G.ptr(t) for t in o], G.ptr(ws), nws, None))
            G.sync()
            n = 2 * 8400
            res.append((ws[:n * 16].view(torch.float32).reshape(2, 8400, 4).cpu().numpy().copy(),
                        ws[n * 16:n * 20].view(torch.float32).reshape(2, 8400).cpu().numpy().copy(), [t.cpu().numpy() for t in o]))
    finally:
        lib.aicam_engine_destroy(e)
    (b2, s2, o2), (b3, s3, o3) = res
    assert o2[0].sum() > 0
    size = np.maximum(np.maximum(b2[..., 2] - b2[..., 0], b2[..., 3] - b2[..., 1]), 1.0)[..., None]
    strong = s2 > 0.05
    print("4x4 vs 2x2 stem: max box diff / size %.2e, max score diff %.2e" % ((np.abs(b3 - b2) / size)[strong].max(), np.abs(s3 - s2).max()))
    assert (np.abs(b3 - b2) / size)[strong].max() < 2e-3
    assert np.abs(s3 - s2).max() < 5e-3
    assert np.array_equal(o2[0], o3[0])  # the same number of detections per frame


def test_reid_embeddings_match_oracle(blobs):
    import gpu_util as G
    from oracle import image_ops, nets
    rng = np.random.default_rng(1)
    frame = synth_image(rng, 540, 960)
    rects = []
    for _ in range(21):
        x, y = int(rng.integers(0, 800)), int(rng.integers(0, 300))
        rects.append((x, y, x + int(rng.integers(20, 150)), y + int(rng.integers(40, 230))))
    xin = image_ops.reid_batch(frame, rects)
    xd = torch.from_numpy(xin).to(G.DEV)
    nhwc = torch.empty((21, 128, 64, 4), dtype=torch.bfloat16, device=G.DEV)
    G.check(G.lib().aicam_nchw_to_nhwc4(G.ptr(xd), 21, 128, 64, G.ptr(nhwc), None))
    e = _engine(blobs["reid"], 8)  # 21 crops through a workspace of 8: sliced
    try:
        assert abs(G.lib().aicam_engine_flops_per_item(e) / 1e9 - 2.242) < 0.05
        feats = torch.empty((21, 512), dtype=torch.float32, device=G.DEV)
        G.check(G.lib().aicam_reid_forward(e, G.ptr(nhwc), 21, None, G.ptr(feats), None))
        # device-side count: capacity 8, only 5 present
        feats2 = torch.zeros((8, 512), dtype=torch.float32, device=G.DEV)
        n_dev = torch.tensor([5], dtype=torch.int32, device=G.DEV)
        G.check(G.lib().aicam_reid_forward(e, G.ptr(nhwc), 8, G.ptr(n_dev), G.ptr(feats2), None))
        G.sync()
    finally:
        G.lib().aicam_engine_destroy(e)
    want = nets.load_net(blobs["reid"]).forward(torch.from_numpy(xin)).numpy()
    got = feats.cpu().numpy()
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    print("reid cosine min %.6f" % cos.min())
    assert cos.min() >= 0.999
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    f2 = feats2.cpu().numpy()
    assert np.array_equal(f2[:5], got[:5]) and not f2[5:].any()


def test_engine_errors(blobs, tmp_path):
    import gpu_util as G
    h = C.c_void_p()
    assert G.lib().aicam_engine_create(str(tmp_path / "missing.aicw").encode(), 0, 1, C.byref(h)) == -3
    bad = tmp_path / "bad.aicw"
    bad.write_bytes(b"not a blob at all, definitely" * 4)
    assert G.lib().aicam_engine_create(str(bad).encode(), 0, 1, C.byref(h)) == -3
    assert b"AICW0001" in G.lib().aicam_last_error()


def test_engine_rejects_corrupt_tensor_entries(blobs, tmp_path):
    """A blob whose tensor table lies (short byte count for its dims, offset past the file, offset + size wrapping
    around 2^64, unaligned offset) is refused with AICAM_ERR_IO instead of being read out of bounds."""
    import struct
    import gpu_util as G
    raw = bytearray(open(blobs["reid"], "rb").read())
    entry = struct.Struct("<64sI4IQQ")
    name, nd, d0, d1, d2, d3, off, nbytes = entry.unpack_from(raw, 48)
    h = C.c_void_p()

    def attempt(new_off, new_nbytes, tag):
        bad = bytearray(raw)
        entry.pack_into(bad, 48, name, nd, d0, d1, d2, d3, new_off, new_nbytes)
        path = tmp_path / ("corrupt_%s.aicw" % tag)
        path.write_bytes(bytes(bad))
        rc = G.lib().aicam_engine_create(str(path).encode(), 0, 1, C.byref(h))
        assert rc == -3, (tag, rc, G.lib().aicam_last_error())

    attempt(off, nbytes // 2, "short")                     # dims promise twice the floats the entry holds
    attempt(len(raw) + 64, nbytes, "past_end")
    attempt(2 ** 64 - 64, 128, "wraps")                    # off + nbytes overflows to a small number
    attempt(off + 2, nbytes, "unaligned")
    attempt(off, nbytes + 4, "long")
