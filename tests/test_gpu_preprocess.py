"""K1 (fused letterbox / BGR->RGB / normalise / layout) vs the oracle, bit for bit (-m gpu)."""
import numpy as np
import pytest
import torch

from scenarios import IMAGEOPS_SEED, LETTERBOX_SIZES, synth_image

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hw", LETTERBOX_SIZES + [(2160, 3840), (360, 640), (1280, 1280)], ids=str)
def test_preprocess_bit_exact(hw):
    import gpu_util as G
    from oracle import image_ops
    rng = np.random.default_rng(IMAGEOPS_SEED + hw[0] * 7 + hw[1])
    frames = np.stack([synth_image(rng, *hw), rng.integers(0, 256, (*hw, 3), dtype=np.uint8)])
    fd = torch.from_numpy(frames).to(G.DEV)
    got0 = G.preprocess(fd, 0).cpu().numpy()
    got1 = G.preprocess(fd, 1).float().cpu().numpy()
    for i in range(2):
        want, ratios, pad = image_ops.preprocess_yolo_input(frames[i])
        assert np.array_equal(got0[i].view(np.uint32), want[0].view(np.uint32)), \
            "fp32 NCHW differs: %d px" % (got0[i] != want[0]).sum()
        wb = torch.from_numpy(want[0]).to(torch.bfloat16).float().numpy().transpose(1, 2, 0)
        assert np.array_equal(got1[i][..., :3], wb)
        assert not got1[i][..., 3].any()
    m = __import__("ai_camera_b200._lib", fromlist=["Letterbox"]).Letterbox()
    import ctypes as C
    G.check(G.lib().aicam_letterbox_params(hw[0], hw[1], C.byref(m)))
    assert (m.ratio, m.pad_w, m.pad_h) == (ratios[0], pad[0], pad[1])


def test_preprocess_full_batch_checksum():
    """BASELINE config 2 size: 64 x 1080p frames; the exact 3:1 case is a pure gather."""
    import gpu_util as G
    g = torch.Generator(device="cpu").manual_seed(1234)
    frames = torch.randint(0, 256, (64, 1080, 1920, 3), dtype=torch.uint8, generator=g)
    out = G.preprocess(frames.to(G.DEV), 0).cpu()
    want = torch.full((64, 3, 640, 640), 114.0 / 255.0)
    want[:, :, 140:500, :] = (frames[:, 1::3, 1::3, :].flip(-1).permute(0, 3, 1, 2).float() / 255.0)
    assert torch.equal(out, want)


def test_space_to_depth_format_is_a_pure_permutation():
    """Format 2 (what the 2x2-window stem consumes) holds exactly the NHWC4 values of format 1:
    block (Y, X) channel (sy*2 + sx)*4 + c  ==  pixel (2Y + sy, 2X + sx) channel c."""
    import gpu_util as G
    rng = np.random.default_rng(7)
    frames = torch.from_numpy(rng.integers(0, 256, (2, 540, 960, 3), dtype=np.uint8)).to(G.DEV)
    a = G.preprocess(frames, 1).view(torch.int16).cpu().numpy()           # [B,640,640,4]
    b = G.preprocess(frames, 2).view(torch.int16).cpu().numpy()           # [B,320,320,16]
    want = a.reshape(2, 320, 2, 320, 2, 4).transpose(0, 1, 3, 2, 4, 5).reshape(2, 320, 320, 16)
    assert np.array_equal(b, want)


@pytest.mark.parametrize("hw", [(1080, 1920), (540, 960), (333, 517)])
def test_4x4_block_format_is_a_pure_permutation(hw):
    """Format 3 (what the 4x4-block stem of yolov8n consumes) = the format-2 blocks grouped 2x2 once more:
    block (Y, X) channel ((by*2 + bx)*4 + (py*2 + px))*4 + c  ==  pixel (4Y + 2by + py, 4X + 2bx + px) channel c.
    1080p runs the row-staged decimation kernel, the other sizes the generic one."""
    import gpu_util as G
    rng = np.random.default_rng(8)
    frames = torch.from_numpy(rng.integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)).to(G.DEV)
    a = G.preprocess(frames, 1).view(torch.int16).cpu().numpy()           # [B,640,640,4]
    b = G.preprocess(frames, 3).view(torch.int16).cpu().numpy()           # [B,160,160,64]
    want = a.reshape(2, 160, 2, 2, 160, 2, 2, 4).transpose(0, 1, 4, 2, 5, 3, 6, 7).reshape(2, 160, 160, 64)
    assert np.array_equal(b, want)
