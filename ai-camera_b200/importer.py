"""Real-checkpoint importer: ONNX initialisers or a PyTorch state dict -> flat ".aicw" weight blob.

Replaces the reference's model provisioning (``/root/reference/scripts/download_models.sh:7-8`` fetches
``yolov8n.onnx`` / ``deepsort.onnx``; ``scripts/export_trt_engines.sh:25-37`` turns them into TensorRT
engines): the same files become the blob ``aicam_engine_create`` loads.

    python -m ai_camera_b200.importer yolov8n.onnx  models/detection/yolov8n.aicw
    python -m ai_camera_b200.importer ckpt.t7       models/reid/deepsort_reid.aicw

Accepted sources
  * ``.onnx``: read with a minimal protobuf wire-format reader (the ``onnx`` package is not needed): every
    ``graph.initializer`` TensorProto (float32 / float16, raw_data or float_data) by name;
  * a ``torch.save``d state dict (or a dict holding one under ``net_dict`` / ``state_dict`` / ``model``), e.g.
    deep_sort_pytorch's ``ckpt.t7``.
Tensor names are the public ones (Ultralytics ``model.<i>....conv.weight`` + ``.bn.*``; deep_sort_pytorch
``conv.0.weight`` + ``conv.1.*``, ``layer<k>.<b>.conv1.weight`` + ``bn1.*``, ``downsample.0`` + ``downsample.1.*``).
BatchNorm is folded here (w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps)) unless the
export already fused it (a ``.bias`` next to the convolution weight).  ONNX exports rename initialisers that went
through Conv+BN fusion to ``onnx::Conv_<n>``: those are matched to layers by ORDER and shape against the
architecture's layer list (``weights.yolov8_conv_specs`` / ``weights.reid_conv_specs``).

Weights are stored as float32; the device path rounds them to bf16 at load, the CPU oracle uses them as they are,
so parity tolerances against a real checkpoint include that rounding (pass ``round_bf16=True`` to store
bf16-representable values, as the synthetic blobs do).
"""
import sys
from collections import OrderedDict

import numpy as np

from . import weights as W

# ---- minimal protobuf reader (wire format: varint / 64-bit / length-delimited / 32-bit) ---------------------------


def _varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) of one message; value is an int or a memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fno, wt, v


_ONNX_FLOAT, _ONNX_FLOAT16, _ONNX_DOUBLE, _ONNX_INT64 = 1, 10, 11, 7


def _tensor_proto(buf):
    """TensorProto: dims = 1, data_type = 2, float_data = 4, int64_data = 7, name = 8, raw_data = 9, double_data = 10."""
    dims, dtype, name, raw, floats = [], 0, "", None, []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(v)
            else:  # packed
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    dims.append(d)
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 4:
            floats.append(np.frombuffer(bytes(v), "<f4") if wt == 2 else np.frombuffer(v, "<f4"))
    if dtype == _ONNX_FLOAT:
        arr = np.frombuffer(raw, "<f4") if raw is not None else (np.concatenate(floats) if floats else np.zeros(0, np.float32))
    elif dtype == _ONNX_FLOAT16 and raw is not None:
        arr = np.frombuffer(raw, "<f2").astype(np.float32)
    elif dtype == _ONNX_DOUBLE and raw is not None:
        arr = np.frombuffer(raw, "<f8").astype(np.float32)
    else:
        return name, None
    arr = np.array(arr, np.float32)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def read_onnx_initializers(path):
    """OrderedDict name -> float32 array of every float initialiser, in graph order.  ModelProto.graph = 7,
    GraphProto.initializer = 5."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out = OrderedDict()
    for fno, wt, v in _fields(buf):
        if fno == 7 and wt == 2:
            for gno, gwt, gv in _fields(v):
                if gno == 5 and gwt == 2:
                    name, arr = _tensor_proto(gv)
                    if arr is not None:
                        out[name] = arr
    if not out:
        raise RuntimeError("%s holds no float initialisers (is it an ONNX model?)" % path)
    return out


def read_state_dict(path):
    import torch
    obj = torch.load(path, map_location="cpu", weights_only=True)
    for key in ("net_dict", "state_dict", "model"):
        if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
            obj = obj[key]
    if not isinstance(obj, dict):
        raise RuntimeError("%s does not hold a state dict" % path)
    return OrderedDict((k, v.detach().float().numpy()) for k, v in obj.items() if hasattr(v, "detach") and v.dtype.is_floating_point)


# ---- BatchNorm folding and layer matching ----------------------------------------------------------------------------
def fold_bn(w, gamma, beta, mean, var, eps):
    s = (gamma / np.sqrt(var + np.float32(eps))).astype(np.float32)
    return (w * s[:, None, None, None]).astype(np.float32), (beta - mean * s).astype(np.float32)


def _bn_prefix(conv_name, kind):
    """Name of the BatchNorm that follows a convolution in the public checkpoints."""
    if kind == W.KIND_YOLOV8:  # model.N...cvK.conv -> model.N...cvK.bn
        return conv_name[:-len(".conv")] + ".bn" if conv_name.endswith(".conv") else None
    if conv_name == "conv.0":
        return "conv.1"
    if conv_name.endswith(".conv1"):
        return conv_name[:-len("conv1")] + "bn1"
    if conv_name.endswith(".conv2"):
        return conv_name[:-len("conv2")] + "bn2"
    if conv_name.endswith(".downsample.0"):
        return conv_name[:-1] + "1"
    return None


def _infer_yolov8(tensors):
    """(scale, nc) from the shapes of a YOLOv8 checkpoint's named tensors."""
    w0 = tensors.get("model.0.conv.weight")
    for scale in W.YOLOV8_SCALES:
        if w0 is not None and w0.shape[0] == W.yolov8_dims(scale)["c"][0]:
            # cv3.*.2 carries the class count
            head = tensors.get("model.22.cv3.0.2.weight")
            # (n and s differ in c1: 16 vs 32; m: 48)
            return scale, int(head.shape[0]) if head is not None else 80
    raise RuntimeError("cannot infer the YOLOv8 scale from model.0.conv.weight")


def _assemble(specs, kind, tensors, eps):
    named = OrderedDict()
    missing = []
    for name, cin, cout, k, s, act in specs:
        w = tensors.get(name + ".weight")
        if w is None:
            missing.append(name)
            continue
        if tuple(w.shape) != (cout, cin, k, k):
            raise RuntimeError("%s.weight has shape %s, expected %s" % (name, tuple(w.shape), (cout, cin, k, k)))
        b = tensors.get(name + ".bias")
        bn = _bn_prefix(name, kind)
        if bn is not None and (bn + ".running_mean") in tensors:
            base_b = b if b is not None else np.zeros(cout, np.float32)
            w, bb = fold_bn(w, tensors[bn + ".weight"], tensors[bn + ".bias"], tensors[bn + ".running_mean"],
                            tensors[bn + ".running_var"], eps)
            s_ = tensors[bn + ".weight"] / np.sqrt(tensors[bn + ".running_var"] + np.float32(eps))
            b = (bb + base_b * s_).astype(np.float32)  # a conv bias in front of a BN is scaled by it as well
        elif b is None:
            raise RuntimeError("%s has neither a bias nor a BatchNorm to fold" % name)
        named[name + ".weight"] = np.ascontiguousarray(w, np.float32)
        named[name + ".bias"] = np.ascontiguousarray(b, np.float32)
    return named, missing


def _match_by_order(specs, tensors):
    """ONNX exports with fused Conv+BN: anonymous (weight [cout,cin,k,k], bias [cout]) pairs in layer order."""
    convs = [(n, t) for n, t in tensors.items() if t.ndim == 4]
    biases = {n: t for n, t in tensors.items() if t.ndim == 1}
    out = OrderedDict()
    ci = 0
    names = list(tensors)
    for name, cin, cout, k, s, act in specs:
        while ci < len(convs) and tuple(convs[ci][1].shape) != (cout, cin, k, k):
            ci += 1
        if ci == len(convs):
            raise RuntimeError("no initialiser of shape %s left for layer %s" % ((cout, cin, k, k), name))
        wname, w = convs[ci]
        ci += 1
        # its bias: the next 1-D initialiser of length cout after the weight in graph order
        b = None
        for n in names[names.index(wname) + 1:]:
            if n in biases and biases[n].shape[0] == cout:
                b = biases[n]
                break
            if tensors[n].ndim == 4:
                break
        if b is None:
            raise RuntimeError("no bias initialiser follows %s (layer %s)" % (wname, name))
        out[name + ".weight"], out[name + ".bias"] = np.ascontiguousarray(w, np.float32), np.ascontiguousarray(b, np.float32)
    return out


def import_tensors(tensors, kind=None, round_bf16=False):
    """dict name -> array of a checkpoint -> (kind, params, blob tensors)."""
    is_yolo = any(n.startswith("model.") for n in tensors) if kind is None else kind == W.KIND_YOLOV8
    has_names = any(n.endswith(".weight") and (n.startswith("model.") or n.startswith("layer") or n.startswith("conv.")) for n in tensors)
    if kind is None and not has_names:
        # anonymous ONNX initialisers: a first convolution over 3 channels with 64 outputs at stride 1 is the ReID stem
        first = next(t for t in tensors.values() if t.ndim == 4)
        is_yolo = not (first.shape[0] == 64 and first.shape[1] == 3)
    if is_yolo:
        if has_names:
            scale, nc = _infer_yolov8(tensors)
        else:
            first = next(t for t in tensors.values() if t.ndim == 4)
            scale = next(s for s in W.YOLOV8_SCALES if W.yolov8_dims(s)["c"][0] == first.shape[0])
            nc = 80
        specs = W.yolov8_conv_specs(scale, nc)
        D = W.yolov8_dims(scale, nc)
        params, kind, eps = D["c"] + [D["n_small"], D["n_large"], nc], W.KIND_YOLOV8, 1e-3
    else:
        specs, params, kind, eps = W.reid_conv_specs(), [512, 0, 0, 0, 0, 0, 0, 0], W.KIND_REID, 1e-5
    if has_names:
        out, missing = _assemble(specs, kind, tensors, eps)
        if missing:
            raise RuntimeError("checkpoint lacks layers: %s" % ", ".join(missing[:6]))
    else:
        out = _match_by_order(specs, tensors)
    if round_bf16:
        for n in out:
            if n.endswith(".weight"):
                out[n] = W._bf16_round(out[n])
    return kind, params, out


def import_file(src, dst, round_bf16=False):
    tensors = read_onnx_initializers(src) if str(src).lower().endswith(".onnx") else read_state_dict(src)
    kind, params, out = import_tensors(tensors, round_bf16=round_bf16)
    W.write_blob(dst, kind, params, out)
    return kind, params, out


# ---- minimal ONNX writer (initialisers only): lets the round-trip test build a file without the onnx package ------
def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_field(fno, wt, payload):
    key = _enc_varint((fno << 3) | wt)
    if wt == 2:
        return key + _enc_varint(len(payload)) + payload
    return key + payload


def write_onnx_initializers(path, tensors, use_float_data=()):
    """A syntactically valid ModelProto holding only graph.initializer entries (float32)."""
    graph = bytearray()
    for name, t in tensors.items():
        t = np.asarray(t, np.float32)  # (ascontiguousarray would turn a 0-d scalar into shape (1,))
        tp = bytearray()
        for d in t.shape:
            tp += _enc_field(1, 0, _enc_varint(int(d)))
        tp += _enc_field(2, 0, _enc_varint(_ONNX_FLOAT))
        tp += _enc_field(8, 2, name.encode())
        if name in use_float_data:
            tp += _enc_field(4, 2, t.astype("<f4").tobytes())
        else:
            tp += _enc_field(9, 2, t.astype("<f4").tobytes())
        graph += _enc_field(5, 2, bytes(tp))
    model = _enc_field(1, 0, _enc_varint(8)) + _enc_field(7, 2, bytes(graph))  # ir_version = 8
    with open(path, "wb") as f:
        f.write(model)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 2:
        print(__doc__)
        return 2
    kind, params, out = import_file(argv[0], argv[1], round_bf16="--bf16" in argv)
    print("wrote %s: kind %d, params %s, %d tensors" % (argv[1], kind, params, len(out)))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
