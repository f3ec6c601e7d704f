"""Size-independent properties of the whole hot path at BASELINE.json's full configuration
(64 streams of 1080p frames on one GPU), where the CPU oracle would take minutes (-m gpu):

* determinism: the same frames through a freshly reset pipeline give bit-identical track tables;
* stream independence: permuting the streams of the batch permutes the outputs and nothing else
  (each stream owns its tracker state and id counter; no cross-stream arithmetic anywhere);
* sanity of the tables: ids unique per stream and >= 1, boxes well-formed (the reference reports the
  un-clipped Kalman posterior, deepsort_tracker.py:125-141), counts <= capacity.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

S, HW, STEPS = 64, (1080, 1920), 5


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    from ai_camera_b200 import synth
    from ai_camera_b200.pipeline import TrackingPipeline
    yolo, reid = synth.make_blobs(str(tmp_path_factory.mktemp("blobs")))
    video = synth.SynthVideo(S, HW, n_frames=STEPS, device="cuda:0", seed=4321)
    pipe = TrackingPipeline(yolo, reid, S, "cuda:0", max_tracks=128, max_crops=S * 40)
    synth.apply_class_bias(pipe.detector.engine, synth.shifted_class_bias(yolo))
    return video, pipe


def _run(pipe, video, perm=None):
    pipe.tracker.reset()
    outs = []
    for t in range(STEPS):
        fr = video.ring[t] if perm is None else video.ring[t][perm]
        tr, cf, cnt = pipe.step(fr)
        torch.cuda.synchronize()
        outs.append((tr.cpu().numpy().copy(), cf.cpu().numpy().copy(), cnt.cpu().numpy().copy()))
    assert not pipe.tracker.overflow().any()
    return outs


def test_full_size_step_is_deterministic_and_stream_independent(world):
    video, pipe = world
    a = _run(pipe, video)
    b = _run(pipe, video)
    perm = torch.from_numpy(np.random.default_rng(3).permutation(S)).to("cuda:0")
    c = _run(pipe, video, perm)
    p = perm.cpu().numpy()
    reported = 0
    for t in range(STEPS):
        (tr, cf, cnt), (tr2, cf2, cnt2), (tr3, cf3, cnt3) = a[t], b[t], c[t]
        assert np.array_equal(cnt, cnt2) and np.array_equal(cnt[p], cnt3)
        for s in range(S):
            n = int(cnt[s])
            assert 0 <= n <= tr.shape[1]
            assert np.array_equal(tr[s, :n], tr2[s, :n]) and np.array_equal(cf[s, :n].view(np.uint32), cf2[s, :n].view(np.uint32))
        for k in range(S):  # stream p[k] of the first run is stream k of the permuted run
            n = int(cnt3[k])
            assert np.array_equal(tr3[k, :n], tr[p[k], :n])
            assert np.array_equal(cf3[k, :n].view(np.uint32), cf[p[k], :n].view(np.uint32))
        for s in range(S):
            n = int(cnt[s])
            rows = tr[s, :n]
            assert len(set(rows[:, 4].tolist())) == n and (rows[:, 4] >= 1).all()
            assert (rows[:, 0] <= rows[:, 2]).all() and (rows[:, 1] <= rows[:, 3]).all()
            reported += n
    assert reported > S  # tracks are confirmed (n_init = 3) and reported within the five steps
