"""Seeded synthetic multi-object sequences for tracker parity tests.

A scenario is a list of frames; each frame is a dict with
    boxes   (N,4) float32 xyxy in original-frame pixels
    scores  (N,)  float32
    classes (N,)  int32 COCO ids
    feats   (N,F) float32 appearance features ("planted": they replace the ReID net)
The generator exercises what the reference's self-tests exercise
(src/tracker/core/tracker_core.py:201-331, deepsort_tracker.py:203-345): births, deaths,
misses, confirmation through the IoU fallback, deletion after max_age, tentative tracks
deleted on the first miss - plus low-confidence and untracked-class detections, boxes
that leave the frame (invalid crops -> feature None), near-duplicate appearance and
crossing objects.
"""
import numpy as np

TRACKED = (0, 2, 3, 5, 7)
UNTRACKED = (1, 9, 16, 24, 56)


def make_scenario(seed, n_frames=40, n_objects=8, frame_hw=(1080, 1920), feat_dim=512,
                  p_miss=0.1, p_fp=0.1, p_lowconf=0.05, p_untracked=0.05, feat_noise=0.05,
                  shuffle=True, size_range=(40.0, 260.0)):
    rng = np.random.default_rng(seed)
    H, W = frame_hw
    objs = []

    def new_obj(t0):
        h = rng.uniform(*size_range)
        w = h * rng.uniform(0.3, 0.7)
        base = rng.normal(size=feat_dim).astype(np.float32)
        base /= np.linalg.norm(base)
        return dict(
            cx=rng.uniform(-0.05 * W, 1.05 * W), cy=rng.uniform(-0.05 * H, 1.05 * H),
            vx=rng.normal(0, 6.0), vy=rng.normal(0, 3.0), w=w, h=h,
            cls=int(rng.choice(TRACKED)), base=base, t0=t0,
            t1=t0 + int(rng.integers(5, max(6, n_frames))),
            occl=(lambda a: (a, a + int(rng.integers(1, 6))))(int(rng.integers(t0, t0 + n_frames))))

    for _ in range(n_objects):
        objs.append(new_obj(0 if rng.random() < 0.7 else int(rng.integers(0, n_frames // 2 + 1))))
    frames = []
    for t in range(n_frames):
        if rng.random() < 0.15:
            objs.append(new_obj(t))
        boxes, scores, classes, feats = [], [], [], []
        for o in objs:
            if not (o["t0"] <= t < o["t1"]):
                continue
            if o["occl"][0] <= t < o["occl"][1] or rng.random() < p_miss:
                continue
            dt = t - o["t0"]
            cx = o["cx"] + o["vx"] * dt + rng.normal(0, 1.0)
            cy = o["cy"] + o["vy"] * dt + rng.normal(0, 1.0)
            w = o["w"] * (1 + rng.normal(0, 0.01))
            h = o["h"] * (1 + rng.normal(0, 0.01))
            boxes.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
            s = rng.uniform(0.1, 0.29) if rng.random() < p_lowconf else rng.uniform(0.35, 0.95)
            scores.append(s)
            classes.append(int(rng.choice(UNTRACKED)) if rng.random() < p_untracked else o["cls"])
            f = o["base"] * rng.uniform(0.5, 2.0) + feat_noise * rng.normal(size=feat_dim)
            feats.append(f.astype(np.float32))
        n_fp = rng.poisson(p_fp * max(1, n_objects))
        for _ in range(n_fp):
            h = rng.uniform(*size_range)
            w = h * rng.uniform(0.3, 0.7)
            cx, cy = rng.uniform(0, W), rng.uniform(0, H)
            boxes.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
            scores.append(rng.uniform(0.3, 0.9))
            classes.append(int(rng.choice(TRACKED)))
            feats.append(rng.normal(size=feat_dim).astype(np.float32))
        n = len(boxes)
        order = rng.permutation(n) if shuffle else np.arange(n)
        frames.append(dict(
            boxes=np.asarray(boxes, np.float32).reshape(n, 4)[order],
            scores=np.asarray(scores, np.float32).reshape(n)[order],
            classes=np.asarray(classes, np.int32).reshape(n)[order],
            feats=np.asarray(feats, np.float32).reshape(n, feat_dim)[order]))
    return frames


# name -> scenario kwargs; shared by the golden generator and the parity tests
GOLDEN_SCENARIOS = {
    "small": dict(seed=11, n_frames=30, n_objects=4),
    "medium": dict(seed=12, n_frames=60, n_objects=12, p_miss=0.15),
    "crowded": dict(seed=13, n_frames=40, n_objects=40, p_miss=0.1, p_fp=0.05,
                    size_range=(30.0, 120.0)),
    "lookalike": dict(seed=14, n_frames=50, n_objects=10, feat_noise=0.6, p_miss=0.2),
    "longgap": dict(seed=15, n_frames=120, n_objects=6, p_miss=0.45, p_fp=0.02),
    "tiny_dim": dict(seed=16, n_frames=25, n_objects=5, feat_dim=16),
    "fifo": dict(seed=17, n_frames=45, n_objects=5, p_miss=0.03, p_fp=0.02, feat_dim=64),
}

# name -> non-default tracker parameters (DeepSORT constructor keywords,
# src/tracker/deepsort_tracker.py:21-30); "fifo" drives the gallery budget and max_age paths
GOLDEN_TRACKER_KW = {
    "fifo": dict(nn_budget=4, n_init=2, max_age=4),
}


def synth_image(rng, h, w):
    """Seeded HxWx3 uint8 test image: blocky low-frequency colour + a ramp + noise, so that
    bilinear interpolation sees edges, gradients and texture.  numpy only."""
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[:h, :w]
    ramp = (np.arange(w, dtype=np.float32)[None, :, None] * 0.11 +
            np.arange(h, dtype=np.float32)[:, None, None] * 0.07)
    img = 0.6 * img + 0.4 * (ramp % 256.0) + rng.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


LETTERBOX_SIZES = [(1080, 1920), (540, 960), (720, 1280), (480, 640), (300, 500), (1000, 700),
                   (641, 1283), (640, 640), (100, 37)]
REID_CROP_SIZES = [(300, 120), (128, 64), (64, 32), (17, 9), (500, 333), (1, 1), (2, 200), (200, 3),
                   (90, 41), (256, 128)]
IMAGEOPS_SEED = 77
