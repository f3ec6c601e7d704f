"""K5 detection filter + ROI crop/resize/normalise vs the oracle, bit for bit (-m gpu)."""
import numpy as np
import pytest
import torch

from scenarios import synth_image

pytestmark = pytest.mark.gpu


def test_crops_bit_exact_and_filter():
    import gpu_util as G
    from ai_camera_b200.config import tracked_class_mask
    from oracle import image_ops
    from oracle.tracker import DeepSORT
    from oracle.tracker import crop_rect as oracle_crop_rect
    rng = np.random.default_rng(42)
    B, H, W, K = 3, 540, 960, 24
    frames = np.stack([synth_image(rng, H, W) for _ in range(B)])
    boxes = np.zeros((B, K, 4), np.float32)
    scores = np.zeros((B, K), np.float32)
    labels = np.zeros((B, K), np.int32)
    num = np.asarray([K, 17, 0], np.int32)
    special = [(10.2, 20.7, 74.9, 148.99), (100, 100, 228, 356), (-30.5, -10.2, 40.3, 90.8), (900.4, 500.1, 990.0, 560.0),
               (50, 50, 50.5, 90), (300.9, 200.2, 301.1, 201.9), (-50, -50, -10, -5), (5, 5, 700, 530), (400, 300, 464, 428)]
    for b in range(B):
        for k in range(K):
            if k < len(special):
                boxes[b, k] = special[k]
            else:
                x, y = rng.uniform(-20, W), rng.uniform(-20, H)
                boxes[b, k] = (x, y, x + rng.uniform(2, 300), y + rng.uniform(2, 400))
            scores[b, k] = rng.choice([0.2, 0.3, 0.31, 0.9])
            labels[b, k] = rng.choice([0, 2, 3, 5, 7, 1, 9, 79])
    fd = torch.from_numpy(frames).to(G.DEV)
    bd, sd, ld, nd = (torch.from_numpy(a).to(G.DEV) for a in (boxes, scores, labels, num))
    cap = B * K
    det_index = torch.full((B, K), -7, dtype=torch.int32, device=G.DEV)
    det_count = torch.zeros(B, dtype=torch.int32, device=G.DEV)
    crop_slot = torch.full((B, K), -7, dtype=torch.int32, device=G.DEV)
    crop_rect = torch.zeros((cap, 5), dtype=torch.int32, device=G.DEV)
    crop_count = torch.zeros(2, dtype=torch.int32, device=G.DEV)
    crops0 = torch.zeros((cap, 3, 128, 64), dtype=torch.float32, device=G.DEV)
    crops1 = torch.zeros((cap, 128, 64, 4), dtype=torch.bfloat16, device=G.DEV)
    crops2 = torch.full((cap, 128, 64, 8), 7.0, dtype=torch.bfloat16, device=G.DEV)  # NHWC8 for the fused stem
    lo, hi = tracked_class_mask()
    for fmt, crops in ((0, crops0), (1, crops1), (2, crops2)):
        G.check(G.lib().aicam_reid_crops(G.ptr(fd), B, H, W, G.ptr(bd), G.ptr(sd), G.ptr(ld), G.ptr(nd), K, 0.3, lo, hi,
                                         fmt, cap, G.ptr(det_index), G.ptr(det_count), G.ptr(crop_slot), G.ptr(crop_rect),
                                         G.ptr(crops), G.ptr(crop_count), None))
    G.sync()
    di, dc, cs = det_index.cpu().numpy(), det_count.cpu().numpy(), crop_slot.cpu().numpy()
    cr, cc = crop_rect.cpu().numpy(), int(crop_count[0].item())
    ds = DeepSORT()
    row = 0
    for b in range(B):
        keep = ds.filter_indices(scores[b, :num[b]], labels[b, :num[b]])
        assert dc[b] == len(keep) and np.array_equal(di[b, :dc[b]], keep)
        for k, i in enumerate(keep):
            r = oracle_crop_rect(boxes[b, i], H, W)
            if r is None:
                assert cs[b, k] == -1
                continue
            assert cs[b, k] == row and tuple(cr[row]) == (b,) + r
            want = image_ops.reid_batch(frames[b], [r])[0]
            assert np.array_equal(crops0[row].cpu().numpy().view(np.uint32), want.view(np.uint32)), (b, i, r)
            wb = torch.from_numpy(want).to(torch.bfloat16).float().numpy().transpose(1, 2, 0)
            assert np.array_equal(crops1[row].float().cpu().numpy()[..., :3], wb)
            c2 = crops2[row].float().cpu().numpy()
            assert np.array_equal(c2[..., :3], wb) and not c2[..., 3:].any()
            row += 1
    assert cc == row
