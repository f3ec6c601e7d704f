"""The oracle tracker is pinned, bit for bit, to the golden fixtures recorded from the REAL
reference (tests/golden/make_golden.py): returned tuples, ids, lifecycle counters and the
float32 Kalman state of every live track after every frame."""
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, assert_same_tracking, load, run_oracle, scenario_digest
from scenarios import GOLDEN_SCENARIOS, GOLDEN_TRACKER_KW, make_scenario

from oracle import kalman
from oracle.tracker import Det, iou


@pytest.mark.parametrize("name", sorted(GOLDEN_SCENARIOS))
def test_tracker_matches_reference_golden(name):
    frames = make_scenario(**GOLDEN_SCENARIOS[name])
    want = load("tracker_%s.npz" % name)
    assert scenario_digest(frames) == bytes(want["digest"]).decode(), \
        "scenario generator no longer reproduces the inputs the golden file was recorded on"
    got = run_oracle(frames, **GOLDEN_TRACKER_KW.get(name, {}))
    assert_same_tracking(got, want, name)


def test_kalman_matches_reference_golden():
    g = load("kalman.npz")
    for z, want in zip(g["z0"], g["init"]):
        m, c = kalman.initiate(z)
        assert np.array_equal(np.concatenate([m, c]).view(np.uint32), want.view(np.uint32))
    m, c = kalman.predict(g["pred_in"][:, :8], g["pred_in"][:, 8:])
    assert np.array_equal(np.concatenate([m, c], 1).view(np.uint32), g["pred_out"].view(np.uint32))
    for s, z, want in zip(g["upd_in"], g["upd_z"], g["upd_out"]):
        m, c = kalman.update(s[:8], s[8:], z)
        assert np.array_equal(np.concatenate([m, c]).view(np.uint32), want.view(np.uint32))
    seen = set()
    for s, Z, n, want in zip(g["gate_in"], g["gate_z"], g["gate_n"], g["gate_out"]):
        d2 = kalman.gating_distance(s[:8], s[8:], Z[:n])
        assert np.array_equal(d2.view(np.uint32), want[:n].view(np.uint32))
        seen.add(int(n) == 1)
    assert seen == {True, False}  # both the N = 1 (divide) and N >= 2 (reciprocal) paths


def test_known_answers_from_reference_selftests():
    with open(os.path.join(GOLDEN_DIR, "known_answers.json")) as f:
        ka = json.load(f)
    for case in ka["xyah"]:
        d = Det(case["tlwh"], 0.9, 0, None)
        assert np.allclose(d.to_xyah(), case["xyah"])
    for case in ka["iou"]:
        v = iou(np.asarray(case["a"], np.float32), np.asarray([case["b"]], np.float32))
        assert np.isclose(v[0], case["iou"])
    assert kalman.CHI2_GATE == ka["chi2inv95_4"]


def test_empty_and_ragged_frames():
    """update accepts empty arrays (deepsort_tracker.py:321-323) and ages tracks out."""
    from oracle.tracker import DeepSORT
    t = DeepSORT(n_init=2, max_age=1)
    e = np.array([])
    assert t.update(e, e, e, frame_hw=(100, 100), planted_features=np.zeros((0, 4))) == []
    b = np.array([[10, 10, 30, 50]], np.float32)
    f = np.ones((1, 4), np.float32)
    for _ in range(2):
        out = t.update(b, np.array([0.9], np.float32), np.array([0]), frame_hw=(100, 100),
                       planted_features=f)
    assert [o[4] for o in out] == [1]
    t.update(e, e, e, frame_hw=(100, 100), planted_features=np.zeros((0, 4)))
    assert len(t.tracker_core.tracks) == 1  # tsu = 1, not > max_age
    t.update(e, e, e, frame_hw=(100, 100), planted_features=np.zeros((0, 4)))
    assert len(t.tracker_core.tracks) == 0
