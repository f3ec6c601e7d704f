"""Helpers shared by the oracle and CUDA parity tests: load golden fixtures, run the oracle
tracker over a scenario and pack the result in the fixtures' layout."""
import hashlib
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name))


def sha(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def scenario_digest(frames):
    h = hashlib.sha1()
    for f in frames:
        h.update(sha(f["boxes"], f["scores"], f["classes"], f["feats"]).encode())
    return h.hexdigest()


def run_oracle(frames, frame_hw=(1080, 1920), **kw):
    """Oracle DeepSORT over a scenario -> dict in the golden layout."""
    from oracle.tracker import DeepSORT
    trk = DeepSORT(**kw)
    outs, out_conf, out_off = [], [], [0]
    trk_i, trk_f, trk_off = [], [], [0]
    from oracle.constants import CLASSES
    name_to_id = {n: i for i, n in enumerate(CLASSES)}
    for f in frames:
        res = trk.update(f["boxes"], f["scores"], f["classes"], frame_hw=frame_hw,
                         planted_features=f["feats"])
        for (x1, y1, x2, y2, tid, cname, conf) in res:
            outs.append([x1, y1, x2, y2, tid, name_to_id[cname]])
            out_conf.append(conf)
        out_off.append(len(outs))
        for t in trk.tracker_core.tracks:
            trk_i.append([t.track_id, t.state, t.hits, t.age, t.time_since_update, t.class_id,
                          len(t.features)])
            trk_f.append(np.concatenate([t.mean, t.cov, [np.float32(t.confidence)]]))
        trk_off.append(len(trk_i))
    return dict(out=np.asarray(outs, np.int64).reshape(-1, 6),
                out_conf=np.asarray(out_conf, np.float64),
                out_off=np.asarray(out_off, np.int64),
                trk_i=np.asarray(trk_i, np.int64).reshape(-1, 7),
                trk_f=np.asarray(trk_f, np.float32).reshape(-1, 25),
                trk_off=np.asarray(trk_off, np.int64))


def assert_same_tracking(got, want, what=""):
    for k in ("out_off", "out", "trk_off", "trk_i"):
        assert np.array_equal(got[k], want[k]), "%s: %s differs" % (what, k)
    assert np.array_equal(got["out_conf"], want["out_conf"]), "%s: out_conf differs" % what
    # float state compared bit for bit
    assert np.array_equal(got["trk_f"].view(np.uint32), want["trk_f"].view(np.uint32)), \
        "%s: track float state differs (max abs %g)" % (
            what, np.abs(got["trk_f"] - want["trk_f"]).max())
