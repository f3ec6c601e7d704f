"""Flat weight blobs (".aicw") and seeded synthetic weights for the two CNNs.

The reference loads serialized TensorRT engines
(``/root/reference/src/trt_utils/trt_engine.py:45-60``) built from ONNX files it
downloads at install time (``scripts/download_models.sh:7-8``); neither ships with
it.  This build replaces the engine file by a flat blob of BN-folded convolution
weights that the C library (``aicam_ctx_create``) uploads once.  Tensor names
follow the public checkpoints of the named architectures (Ultralytics YOLOv8:
``model.<i>...conv.weight``; deep_sort_pytorch ReID ``Net``: ``layer1.0.conv1.weight``),
so that a real checkpoint can be imported by folding BN and writing the same names.

Blob layout (little endian):
    0   char[8]  magic "AICW0001"
    8   u32      model kind   (1 = yolov8 detect, 2 = deepsort reid)
    12  u32      params[8]    yolov8: c1,c2,c3,c4,c5,n_small,n_large,nc ; reid: feature_dim,0...
    44  u32      n_tensors
    48  entries  n_tensors x { char name[64]; u32 ndim; u32 dims[4]; u64 offset; u64 nbytes }
    ..  data     float32 tensors, each 64-byte aligned, offsets relative to file start
"""
import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"AICW0001"
KIND_YOLOV8 = 1
KIND_REID = 2
_ENTRY = struct.Struct("<64sI4IQQ")

YOLOV8_SCALES = {  # depth multiple, width multiple, max channels (Ultralytics yolov8.yaml)
    "n": (0.33, 0.25, 1024),
    "s": (0.33, 0.50, 1024),
    "m": (0.67, 0.75, 768),
}


def yolov8_dims(scale="n", nc=80):
    """Channel widths c1..c5, C2f repeats (n_small for base 3, n_large for base 6) and head widths."""
    d, w, mx = YOLOV8_SCALES[scale]

    def ch(c):
        c = min(c, mx) * w
        return int(-(-c // 8) * 8)  # make_divisible(c, 8): ceil to a multiple of 8

    c = [ch(64), ch(128), ch(256), ch(512), ch(1024)]
    n_small = max(round(3 * d), 1)
    n_large = max(round(6 * d), 1)
    cb = max(16, c[2] // 4, 64)
    cc = max(c[2], min(nc, 100))
    return dict(c=c, n_small=n_small, n_large=n_large, nc=nc, cb=cb, cc=cc)


def yolov8_conv_specs(scale="n", nc=80):
    """Ordered list of every convolution: (name, cin, cout, k, stride, act).

    act: "silu" for Conv+BN+SiLU blocks (BN folded into weight/bias), "none" for the
    plain biased 1x1 convolutions that end the Detect branches."""
    D = yolov8_dims(scale, nc)
    c1, c2, c3, c4, c5 = D["c"]
    ns, nl = D["n_small"], D["n_large"]
    specs = []

    def conv(name, cin, cout, k, s):
        specs.append((name + ".conv", cin, cout, k, s, "silu"))

    def c2f(name, cin, cout, n):
        c = cout // 2
        conv(name + ".cv1", cin, 2 * c, 1, 1)
        for j in range(n):
            conv("%s.m.%d.cv1" % (name, j), c, c, 3, 1)
            conv("%s.m.%d.cv2" % (name, j), c, c, 3, 1)
        conv(name + ".cv2", (2 + n) * c, cout, 1, 1)

    conv("model.0", 3, c1, 3, 2)
    conv("model.1", c1, c2, 3, 2)
    c2f("model.2", c2, c2, ns)
    conv("model.3", c2, c3, 3, 2)
    c2f("model.4", c3, c3, nl)
    conv("model.5", c3, c4, 3, 2)
    c2f("model.6", c4, c4, nl)
    conv("model.7", c4, c5, 3, 2)
    c2f("model.8", c5, c5, ns)
    conv("model.9.cv1", c5, c5 // 2, 1, 1)
    conv("model.9.cv2", c5 // 2 * 4, c5, 1, 1)
    c2f("model.12", c5 + c4, c4, ns)
    c2f("model.15", c4 + c3, c3, ns)
    conv("model.16", c3, c3, 3, 2)
    c2f("model.18", c3 + c4, c4, ns)
    conv("model.19", c4, c4, 3, 2)
    c2f("model.21", c4 + c5, c5, ns)
    cb, cc = D["cb"], D["cc"]
    for l, cin in enumerate((c3, c4, c5)):
        conv("model.22.cv2.%d.0" % l, cin, cb, 3, 1)
        conv("model.22.cv2.%d.1" % l, cb, cb, 3, 1)
        specs.append(("model.22.cv2.%d.2" % l, cb, 64, 1, 1, "none"))
        conv("model.22.cv3.%d.0" % l, cin, cc, 3, 1)
        conv("model.22.cv3.%d.1" % l, cc, cc, 3, 1)
        specs.append(("model.22.cv3.%d.2" % l, cc, nc, 1, 1, "none"))
    return specs


def reid_conv_specs():
    """deep_sort_pytorch ``Net``: (name, cin, cout, k, stride, act); BN folded."""
    specs = [("conv.0", 3, 64, 3, 1, "relu")]
    cin = 64
    for li, (cout, down) in enumerate(((64, False), (128, True), (256, True), (512, True)), start=1):
        for b in range(2):
            s = 2 if (down and b == 0) else 1
            bc = cin if b == 0 else cout
            specs.append(("layer%d.%d.conv1" % (li, b), bc, cout, 3, s, "relu"))
            specs.append(("layer%d.%d.conv2" % (li, b), cout, cout, 3, 1, "none"))
            if b == 0 and (down or bc != cout):
                specs.append(("layer%d.%d.downsample.0" % (li, b), bc, cout, 1, s, "none"))
        cin = cout
    return specs


def _bf16_round(x):
    """Round float32 to the nearest bfloat16-representable float32 (ties to even)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def _synth_conv(rng, cin, cout, k, gain, with_bn, eps, zero_mean=False, bias_std=0.1):
    fan_in = cin * k * k
    w = rng.normal(0.0, gain / np.sqrt(fan_in), (cout, cin, k, k)).astype(np.float32)
    if zero_mean:
        # zero-sum filters: the positive mean of post-activation features then adds no constant
        # per-channel offset, which would otherwise dominate a random network's output and make
        # all embeddings / score maps look alike
        w = w - w.mean(axis=(1, 2, 3), keepdims=True)
    if with_bn:
        gamma = rng.uniform(0.8, 1.2, cout).astype(np.float32)
        beta = rng.normal(0, bias_std, cout).astype(np.float32)
        mean = rng.normal(0, bias_std, cout).astype(np.float32)
        var = rng.uniform(0.8, 1.2, cout).astype(np.float32)
        s = gamma / np.sqrt(var + np.float32(eps))
        w = w * s[:, None, None, None]
        b = beta - mean * s
    else:
        b = rng.normal(0, bias_std, cout).astype(np.float32)
    # weights are bf16-representable by construction: the device path (bf16 operands) and
    # the fp32 CPU restatement then multiply by identical weight values
    return _bf16_round(w), b.astype(np.float32)


# class prior of the synthetic detector: a street scene dominated by the classes the
# tracker keeps (src/config.py:53), so that most synthetic detections reach the tracker
SYNTH_CLASS_PRIOR = {0: 2.5, 2: 1.5, 3: 1.0, 5: 1.0, 7: 1.0}


def synth_yolov8_weights(scale="n", nc=80, seed=0, cls_bias=-6.0, cls_gain=15.0, dfl_gain=8.0):
    """Seeded synthetic YOLOv8 weights (BN already folded), variance-preserving so that
    activations stay O(0.1-1) through the ~60 layers.  The last convolution of each branch
    is scaled so that class logits and DFL logits vary by about one unit across anchors;
    ``cls_bias`` (plus SYNTH_CLASS_PRIOR) sets how many anchors pass the score threshold."""
    rng = np.random.default_rng(seed)
    tensors = OrderedDict()
    for name, cin, cout, k, s, act in yolov8_conv_specs(scale, nc):
        if act == "silu":
            w, b = _synth_conv(rng, cin, cout, k, 1.55, True, 1e-3, zero_mean=cin > 3)
        elif ".cv3." in name:
            w, b = _synth_conv(rng, cin, cout, k, cls_gain, False, 0.0)
            b = (b + np.float32(cls_bias)).astype(np.float32)
            for cid, boost in SYNTH_CLASS_PRIOR.items():
                if cid < cout:
                    b[cid] += np.float32(boost)
        else:
            w, b = _synth_conv(rng, cin, cout, k, dfl_gain, False, 0.0)
        tensors[name + ".weight"] = w
        tensors[name + ".bias"] = b
    D = yolov8_dims(scale, nc)
    params = D["c"] + [D["n_small"], D["n_large"], nc]
    return KIND_YOLOV8, params, tensors


def synth_reid_weights(seed=1):
    rng = np.random.default_rng(seed)
    tensors = OrderedDict()
    for name, cin, cout, k, s, act in reid_conv_specs():
        gain = 1.35 if act == "relu" else 0.7
        w, b = _synth_conv(rng, cin, cout, k, gain, True, 1e-5, zero_mean=cin > 3, bias_std=0.01)
        tensors[name + ".weight"] = w
        tensors[name + ".bias"] = b
    return KIND_REID, [512, 0, 0, 0, 0, 0, 0, 0], tensors


def write_blob(path, kind, params, tensors):
    params = list(params) + [0] * (8 - len(params))
    n = len(tensors)
    head = 48 + n * _ENTRY.size
    off = (head + 63) // 64 * 64
    entries, chunks = [], []
    for name, t in tensors.items():
        t = np.ascontiguousarray(t, dtype=np.float32)
        dims = list(t.shape) + [1] * (4 - t.ndim)
        entries.append(_ENTRY.pack(name.encode(), t.ndim, *dims, off, t.nbytes))
        pad = (-t.nbytes) % 64
        chunks.append((t.tobytes(), pad))
        off += t.nbytes + pad
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I8II", kind, *params, n))
        for e in entries:
            f.write(e)
        f.write(b"\0" * ((-head) % 64))
        for data, pad in chunks:
            f.write(data)
            f.write(b"\0" * pad)
    return path


def read_blob(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != MAGIC:
        raise RuntimeError("%s is not an AICW0001 weight blob" % path)
    vals = struct.unpack_from("<I8II", raw, 8)
    kind, params, n = vals[0], list(vals[1:9]), vals[9]
    tensors = OrderedDict()
    for i in range(n):
        name, ndim, d0, d1, d2, d3, off, nbytes = _ENTRY.unpack_from(raw, 48 + i * _ENTRY.size)
        shape = (d0, d1, d2, d3)[:ndim]
        tensors[name.rstrip(b"\0").decode()] = np.frombuffer(
            raw, np.float32, nbytes // 4, off).reshape(shape).copy()
    return kind, params, tensors


def read_kind(path):
    """Model kind of a blob without reading its tensors."""
    with open(path, "rb") as f:
        head = f.read(12)
    if head[:8] != MAGIC:
        raise RuntimeError("%s is not an AICW0001 weight blob" % path)
    return struct.unpack_from("<I", head, 8)[0]
