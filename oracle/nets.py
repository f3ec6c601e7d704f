"""PyTorch-CPU float32 restatements of the two CNNs the hot path runs (oracle; test
infrastructure).

PARITY UNPINNED: the reference runs both networks inside TensorRT engines built from
ONNX files fetched from a third-party URL (``/root/reference/scripts/download_models.sh:7-8``,
call sites ``src/detector/yolo_detector.py:97`` and ``src/tracker/reid_model.py:115``);
no weights, no ONNX and no CPU runtime ship with it.  What is restated here are the
*named* architectures: Ultralytics YOLOv8 detect (yolov8.yaml,
scales n/s/m) and the deep_sort_pytorch ReID ``Net``.  The reference pins only the I/O
contract: input ``images`` 1x3x640x640 fp32 in [0,1]
(``scripts/export_trt_engines.sh:25-28``), ReID input ``input`` Nx3x128x64
(``export_trt_engines.sh:31-34``), 512-d output (``reid_model.py:56``).

Weights come from the same ".aicw" blob the device path loads (BN folded, values
bf16-representable), so both sides multiply by identical weights; the device path
additionally rounds activations to bf16 between layers, which is what the tolerances
in tests/ (boxes 1e-2 relative, embedding cosine >= 0.999) cover.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(tensors, name):
    return torch.from_numpy(np.ascontiguousarray(tensors[name]))


class YoloV8:
    """forward(images (B,3,640,640) fp32) -> list of 3 head maps (B,64+nc,H,W): the raw
    DFL logits and class logits per level (strides 8, 16, 32)."""

    def __init__(self, params, tensors):
        self.c = list(params[:5])
        self.n_small, self.n_large, self.nc = params[5], params[6], params[7]
        self.w = {k: _t(tensors, k) for k in tensors}

    def conv(self, x, name, k, s, act=True):
        y = F.conv2d(x, self.w[name + ".weight"], self.w[name + ".bias"], stride=s, padding=k // 2)
        return F.silu(y) if act else y

    def cbs(self, x, name, k, s):
        return self.conv(x, name + ".conv", k, s)

    def c2f(self, x, name, n, shortcut):
        y = self.cbs(x, name + ".cv1", 1, 1)
        c = y.shape[1] // 2
        ys = [y[:, :c], y[:, c:]]
        for j in range(n):
            z = self.cbs(self.cbs(ys[-1], "%s.m.%d.cv1" % (name, j), 3, 1), "%s.m.%d.cv2" % (name, j), 3, 1)
            ys.append(ys[-1] + z if shortcut else z)
        return self.cbs(torch.cat(ys, 1), name + ".cv2", 1, 1)

    def sppf(self, x, name):
        x = self.cbs(x, name + ".cv1", 1, 1)
        p1 = F.max_pool2d(x, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        p3 = F.max_pool2d(p2, 5, 1, 2)
        return self.cbs(torch.cat([x, p1, p2, p3], 1), name + ".cv2", 1, 1)

    @torch.no_grad()
    def forward(self, images):
        ns, nl = self.n_small, self.n_large
        x = self.cbs(images, "model.0", 3, 2)
        x = self.cbs(x, "model.1", 3, 2)
        x = self.c2f(x, "model.2", ns, True)
        x = self.cbs(x, "model.3", 3, 2)
        p3 = self.c2f(x, "model.4", nl, True)
        x = self.cbs(p3, "model.5", 3, 2)
        p4 = self.c2f(x, "model.6", nl, True)
        x = self.cbs(p4, "model.7", 3, 2)
        x = self.c2f(x, "model.8", ns, True)
        p5 = self.sppf(x, "model.9")
        x = torch.cat([F.interpolate(p5, scale_factor=2, mode="nearest"), p4], 1)
        n12 = self.c2f(x, "model.12", ns, False)
        x = torch.cat([F.interpolate(n12, scale_factor=2, mode="nearest"), p3], 1)
        o3 = self.c2f(x, "model.15", ns, False)
        x = torch.cat([self.cbs(o3, "model.16", 3, 2), n12], 1)
        o4 = self.c2f(x, "model.18", ns, False)
        x = torch.cat([self.cbs(o4, "model.19", 3, 2), p5], 1)
        o5 = self.c2f(x, "model.21", ns, False)
        outs = []
        for l, f in enumerate((o3, o4, o5)):
            b = self.cbs(self.cbs(f, "model.22.cv2.%d.0" % l, 3, 1), "model.22.cv2.%d.1" % l, 3, 1)
            b = self.conv(b, "model.22.cv2.%d.2" % l, 1, 1, act=False)
            c = self.cbs(self.cbs(f, "model.22.cv3.%d.0" % l, 3, 1), "model.22.cv3.%d.1" % l, 3, 1)
            c = self.conv(c, "model.22.cv3.%d.2" % l, 1, 1, act=False)
            outs.append(torch.cat([b, c], 1))
        return outs

    def head_flat(self, images):
        """(B, 8400, 64+nc) float32: anchors in level-major, row-major order - the layout the
        device decode kernel reads."""
        outs = self.forward(images)
        return torch.cat([o.flatten(2).transpose(1, 2) for o in outs], 1).contiguous()


class ReIDNet:
    """forward(crops (N,3,128,64) fp32, ImageNet-normalised) -> (N,512) L2-normalised."""

    def __init__(self, params, tensors):
        self.w = {k: _t(tensors, k) for k in tensors}

    def conv(self, x, name, k, s):
        return F.conv2d(x, self.w[name + ".weight"], self.w[name + ".bias"], stride=s, padding=k // 2)

    def block(self, x, name, s):
        y = F.relu(self.conv(x, name + ".conv1", 3, s))
        y = self.conv(y, name + ".conv2", 3, 1)
        if (name + ".downsample.0.weight") in self.w:
            x = self.conv(x, name + ".downsample.0", 1, s)
        return F.relu(x + y)

    @torch.no_grad()
    def forward(self, crops):
        x = F.relu(self.conv(crops, "conv.0", 3, 1))
        x = F.max_pool2d(x, 3, 2, 1)
        for li in range(1, 5):
            x = self.block(x, "layer%d.0" % li, 2 if li > 1 else 1)
            x = self.block(x, "layer%d.1" % li, 1)
        x = F.avg_pool2d(x, (8, 4), 1).flatten(1)
        return x / x.norm(p=2, dim=1, keepdim=True)


def load_net(path):
    from ai_camera_b200.weights import KIND_REID, KIND_YOLOV8, read_blob
    kind, params, tensors = read_blob(path)
    if kind == KIND_YOLOV8:
        return YoloV8(params, tensors)
    if kind == KIND_REID:
        return ReIDNet(params, tensors)
    raise RuntimeError("unknown blob kind %d" % kind)
