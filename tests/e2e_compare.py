"""Compare two end-to-end runs (device path vs CPU oracle) of detect -> track on the same frames.

The two detectors are different arithmetic (bf16 tensor cores vs fp32), so a detection whose score sits within
the bf16 noise of the 0.3 threshold, or a pair within it of the NMS IoU threshold, can exist on one side only;
it starts a tentative track there and shifts every later id.  What must hold regardless:
  * tracks that both sides report coincide in class and (to a pixel or two) in position,
  * the id correspondence oracle id -> device id is CONSISTENT over time (a relabelling, not id switches),
  * before the first frame where the detection sets differ, the outputs are identical.
The harness logs the minimum margins to the decision thresholds (0.3 score, 0.2 cosine, 0.7 IoU
cost, 9.4877 gate) so that a divergence can be attributed.
"""
import numpy as np


def _iou(a, b):
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0]))
    ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    u = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - iw * ih
    return iw * ih / u if u > 0 else 0.0


def compare_runs(got, want, iou_thr=0.8, px_tol=2):
    """got / want: per frame, list of (x1, y1, x2, y2, id, class_name, conf).  Returns a dict of statistics."""
    mapping, votes = {}, {}
    matched = total = consistent = 0
    strict_prefix = None
    max_px = 0
    for t, (g, w) in enumerate(zip(got, want)):
        same = len(g) == len(w) and all(a[4] == b[4] and a[5] == b[5] and max(abs(a[i] - b[i]) for i in range(4)) <= px_tol
                                        for a, b in zip(g, w))
        if not same and strict_prefix is None:
            strict_prefix = t
        used = set()
        for b in w:
            total += 1
            best, bi = 0.0, -1
            for j, a in enumerate(g):
                if j in used or a[5] != b[5]:
                    continue
                v = _iou(a, b)
                if v > best:
                    best, bi = v, j
            if bi >= 0 and best >= iou_thr:
                used.add(bi)
                matched += 1
                a = g[bi]
                max_px = max(max_px, max(abs(a[i] - b[i]) for i in range(4)))
                votes.setdefault(b[4], {}).setdefault(a[4], 0)
                votes[b[4]][a[4]] += 1
    for oid, v in votes.items():
        did, n = max(v.items(), key=lambda kv: kv[1])
        mapping[oid] = did
        consistent += n
    # injective: two oracle ids must not map to the same device id
    injective = len(set(mapping.values())) == len(mapping)
    return dict(frames=len(want), oracle_tracks=total, matched=matched, consistent=consistent, injective=injective,
                strict_prefix=len(want) if strict_prefix is None else strict_prefix, max_px=max_px,
                distinct_ids=len(mapping))


def oracle_margins(cost_log, det_scores):
    """Minimum distances of the oracle's decision values to their thresholds."""
    m = dict(score=1e9, cosine=1e9, iou=1e9)
    for s in det_scores:
        if len(s):
            m["score"] = min(m["score"], float(np.min(np.abs(np.asarray(s) - 0.3))))
    for entry in cost_log:
        cm, thr = entry[3], entry[4]
        fin = cm[cm < 1e4]
        if fin.size:
            key = "cosine" if thr < 0.5 else "iou"
            m[key] = min(m[key], float(np.min(np.abs(fin - thr))))
    return m
