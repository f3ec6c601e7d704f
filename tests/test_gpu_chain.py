"""Fused convolution chains (csrc/conv_chain.cu) vs the same layers composed in PyTorch fp32 (-m gpu).

The chains are the ones the detector engine fuses: Bottleneck pairs (3x3 -> 3x3 + shortcut), C2f tails (pair + the 1x1
over the concatenation) and Detect-head branches (3x3 -> 3x3 -> 1x1, fp32 out).  Operands are bf16 on the device; the
reference rounds every intermediate activation to bf16 exactly where the kernel does, so the tolerance is the one of a
single layer (bf16 output rounding plus accumulation-order noise propagated through the chain)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def pair(c, res=True, act=1):
    """Bottleneck: 3x3 -> 3x3 (+ the pair's input)."""
    return dict(in_c=c, stages=[
        dict(cin=c, cout=c, k=3, act=act, src=[(0, 0, c)]),
        dict(cin=c, cout=c, k=3, act=act, src=[(1, 0, c)], res=(0, 0, 1) if res else None)])


def c2f_tail(c, cout, res=True):
    """C2f after cv1: the input holds cv1's 2c channels; Bottleneck on the second half; 1x1 over [cv1 out | m out]."""
    return dict(in_c=2 * c, stages=[
        dict(cin=c, cout=c, k=3, act=1, src=[(0, c, c)]),
        dict(cin=c, cout=c, k=3, act=1, src=[(1, 0, c)], res=(0, c, 1) if res else None),
        dict(cin=3 * c, cout=cout, k=1, act=1, src=[(0, 0, 2 * c), (2, 0, c)])])


def c2f_full(cin, c, cout):
    """Whole C2f (n = 1): cv1 1x1 -> Bottleneck -> cv2 1x1."""
    return dict(in_c=cin, stages=[
        dict(cin=cin, cout=2 * c, k=1, act=1, src=[(0, 0, cin)]),
        dict(cin=c, cout=c, k=3, act=1, src=[(1, c, c)]),
        dict(cin=c, cout=c, k=3, act=1, src=[(2, 0, c)], res=(1, c, 1)),
        dict(cin=3 * c, cout=cout, k=1, act=1, src=[(1, 0, 2 * c), (3, 0, c)])])


def head(cin, cm, cout):
    """Detect branch: 3x3 -> 3x3 -> 1x1 (no activation, fp32 out)."""
    return dict(in_c=cin, out_f32=True, stages=[
        dict(cin=cin, cout=cm, k=3, act=1, src=[(0, 0, cin)]),
        dict(cin=cm, cout=cm, k=3, act=1, src=[(1, 0, cm)]),
        dict(cin=cm, cout=cout, k=1, act=0, src=[(2, 0, cm)])])


CASES = [
    # name, B, H, W, chain
    ("pair16_res", 2, 40, 56, pair(16)),
    ("pair16_160", 1, 160, 160, pair(16)),
    ("pair32_res", 2, 80, 80, pair(32)),
    ("pair32_nores_odd", 3, 37, 45, pair(32, res=False)),
    ("pair64_res_stream", 2, 40, 40, pair(64)),
    ("pair64_nores", 3, 40, 40, pair(64, res=False)),
    ("pair128_res_stream", 3, 20, 20, pair(128)),
    ("pair64_relu_res2", 2, 24, 16, dict(in_c=64, stages=[
        dict(cin=64, cout=64, k=3, act=2, src=[(0, 0, 64)]),
        dict(cin=64, cout=64, k=3, act=2, src=[(1, 0, 64)], res=(0, 0, 2))])),
    ("tail16", 2, 160, 160, c2f_tail(16, 32)),
    ("tail32", 2, 80, 80, c2f_tail(32, 64)),
    ("tail32_nores", 1, 80, 80, c2f_tail(32, 64, res=False)),
    ("tail64_nores", 2, 40, 40, c2f_tail(64, 128, res=False)),
    ("full16", 2, 160, 160, c2f_full(32, 16, 32)),
    ("head64_64_64", 2, 80, 80, head(64, 64, 64)),
    ("head64_80_80", 2, 80, 80, head(64, 80, 80)),
    ("head128_64_64", 2, 40, 40, head(128, 64, 64)),
    ("head128_80_80", 3, 40, 40, head(128, 80, 80)),
    ("head256_64_64", 3, 20, 20, head(256, 64, 64)),
    ("head256_64_80", 5, 20, 20, head(256, 64, 80)),
    ("tiny_7x9", 1, 7, 9, pair(32)),
    ("one_row", 2, 1, 33, pair(16)),
]


def bf16r(t):
    return t.to(torch.bfloat16).float()


def run_reference(x, chain, ws, bs):
    """x: fp32 NHWC torch (bf16-rounded values).  Intermediates are rounded to bf16 like the kernel's buffers."""
    bufs = [x.permute(0, 3, 1, 2)]
    n = len(chain["stages"])
    for s, st in enumerate(chain["stages"]):
        xin = torch.cat([bufs[b][:, co:co + cc] for (b, co, cc) in st["src"]], 1)
        y = F.conv2d(xin, torch.from_numpy(ws[s]), torch.from_numpy(bs[s]), padding=st["k"] // 2)
        r = None
        if st.get("res"):
            rb, rco, rmode = st["res"]
            r = bufs[rb][:, rco:rco + st["cout"]]
            if rmode == 2:
                y = y + r
        y = F.silu(y) if st["act"] == 1 else (F.relu(y) if st["act"] == 2 else y)
        if r is not None and st["res"][2] == 1:
            y = y + r
        if s + 1 < n:
            y = bf16r(y)
        bufs.append(y)
    return bufs[-1].permute(0, 2, 3, 1).contiguous()


def run_chain(xd, chain, ws, bs, B, H, W):
    import gpu_util as G
    from ai_camera_b200._lib import ChainDesc, check, ptr
    d = ChainDesc()
    d.batch, d.h, d.w, d.in_c = B, H, W, chain["in_c"]
    d.nstages = len(chain["stages"])
    d.out_f32 = 1 if chain.get("out_f32") else 0
    for s, st in enumerate(chain["stages"]):
        e = d.st[s]
        e.cin, e.cout, e.ksize, e.act, e.nsrc = st["cin"], st["cout"], st["k"], st["act"], len(st["src"])
        for j, (b, co, cc) in enumerate(st["src"]):
            e.src_buf[j], e.src_coff[j], e.src_c[j] = b, co, cc
        if st.get("res"):
            e.res_buf, e.res_coff, e.res_mode = st["res"]
        else:
            e.res_buf, e.res_coff, e.res_mode = -1, 0, 0
    n = d.nstages
    wp = (C.c_void_p * n)(*[w.ctypes.data for w in ws])
    bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
    cl = chain["stages"][-1]["cout"]
    out = torch.full((B, H, W, cl), 7.0, dtype=torch.float32 if d.out_f32 else torch.bfloat16, device=G.DEV)
    check(G.lib().aicam_conv_chain(C.byref(d), ptr(xd), wp, bp, ptr(out), None))
    return out


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_chain_matches_torch(case):
    import gpu_util as G
    name, B, H, W, chain = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    x = G.bf16_round_np(rng.normal(0, 1, (B, H, W, chain["in_c"])))
    ws, bs = [], []
    for st in chain["stages"]:
        k = st["k"]
        ws.append(np.ascontiguousarray(G.bf16_round_np(rng.normal(0, 1.0 / np.sqrt(st["cin"] * k * k), (st["cout"], st["cin"], k, k)))))
        bs.append(rng.normal(0, 0.3, st["cout"]).astype(np.float32))
    xd = torch.from_numpy(x).to(G.DEV).to(torch.bfloat16)
    got = run_chain(xd, chain, ws, bs, B, H, W).float().cpu().numpy()
    want = run_reference(torch.from_numpy(x), chain, ws, bs).numpy()
    assert got.shape == want.shape
    err = np.abs(got - want)
    # an intermediate that rounds the other way in bf16 (1 ulp = 2^-8 relative) moves the next layer's sums slightly
    tol = 3e-2 + 1.5e-2 * np.abs(want)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d elements off, max abs err %.4g at %s (got %.5g want %.5g)" % (
        name, bad.sum(), bad.size, err.max(), np.unravel_index(err.argmax(), err.shape),
        got.flat[err.argmax()], want.flat[err.argmax()])


def test_engine_with_fused_chains_matches_oracle():
    """The detector engine built with its chains fused (AICAM_CHAIN=1: Bottleneck pairs, C2f tails, Detect-head
    branches as single launches) must pass the same oracle comparison as the default single-layer engine; the switch is
    read when the engine is built, hence the subprocess."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AICAM_CHAIN="1", AICAM_CHAIN_DEBUG="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_engine.py"), "-q", "-x", "-m", "gpu",
                        "-k", "yolov8n_head or yolov8_s_m", "-p", "no:cacheprovider", "-s"], env=env, capture_output=True, text=True, cwd=root)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert "conv_chain:" in r.stderr + r.stdout, "the engine did not launch any fused chain\n" + tail
