"""ByteTrack (detect.py:22, 553-577) -- device Kalman predict / update / initiate + IoU cost
matrix, host association and lifecycle -- against the numpy restatement of supervision's ByteTrack
(oracle/bytetrack.py; supervision itself is not installed: parity unpinned)."""
import numpy as np
import pytest
import torch

import manual_yolo_b200 as m
from manual_yolo_b200 import tracking
from oracle import bytetrack as obt

pytestmark = pytest.mark.gpu


def _sequence(n_frames=70, n_obj=14, seed=0):
    """Objects moving linearly with jitter; per frame some are missed (occlusion -> lost -> re-found), some get a low
    score (second association), new ones appear late, two overlap heavily for a while (duplicate handling)."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(100, 1500, (n_obj, 2))
    vel = rng.uniform(-6, 6, (n_obj, 2))
    size = rng.uniform(40, 140, (n_obj, 2))
    birth = np.where(np.arange(n_obj) < n_obj - 4, 0, rng.integers(10, 40, n_obj))
    pos[1] = pos[0] + 4.0                                           # near-duplicates
    vel[1], size[1] = vel[0], size[0]
    frames = []
    for f in range(n_frames):
        rows = []
        for o in range(n_obj):
            if f < birth[o]:
                continue
            c = pos[o] + vel[o] * f + rng.normal(0, 1.5, 2)
            if rng.random() < 0.12 or (o == 3 and 20 <= f < 45):    # missed detections; object 3 is occluded for 25 frames
                continue
            s = rng.uniform(0.5, 0.95)
            if rng.random() < 0.15:
                s = rng.uniform(0.12, 0.24)                          # low-score detection: second association
            wh = size[o] * rng.uniform(0.95, 1.05, 2)
            rows.append([c[0] - wh[0] / 2, c[1] - wh[1] / 2, c[0] + wh[0] / 2, c[1] + wh[1] / 2, s, float(o % 5)])
        if rng.random() < 0.3:                                       # a spurious one-frame detection (unconfirmed -> removed)
            c = rng.uniform(100, 1500, 2)
            rows.append([c[0], c[1], c[0] + 50, c[1] + 60, rng.uniform(0.4, 0.6), 1.0])
        rng.shuffle(rows)
        frames.append(np.asarray(rows, np.float32).reshape(-1, 6))
    return frames


@pytest.mark.parametrize("kw", [dict(), dict(minimum_consecutive_frames=3, lost_track_buffer=10, track_activation_threshold=0.3)])
def test_bytetrack_equals_numpy_restatement_frame_by_frame(cuda_dev, kw):
    frames = _sequence()
    max_det = 64
    trk = tracking.ByteTrack(device=cuda_dev, capacity=128, max_det=max_det, **kw)
    ref = obt.ByteTrackRef(**kw)
    n_ids = set()
    for f, rows in enumerate(frames):
        n = rows.shape[0]
        pad = torch.zeros((1, max_det, 6))
        pad[0, :n] = torch.from_numpy(rows)
        det = m.Detections(pad.to(cuda_dev), torch.zeros((1, max_det), dtype=torch.int32, device=cuda_dev),
                           torch.tensor([n], dtype=torch.int32, device=cuda_dev))
        got = trk.update(det, 0)
        exp = ref.update_with_detections(rows[:, :4], rows[:, 4])
        assert got.tolist() == exp.tolist(), f
        n_ids |= set(got[got >= 0].tolist())
        assert len(trk.tracked) == len(ref.tracked_tracks) and len(trk.lost) == len(ref.lost_tracks), f
    assert len(n_ids) >= 12                                              # the objects were actually tracked
    # Kalman states of the live tracks agree with the float64 numpy filter
    st = trk.states()
    rt = {t.internal_track_id: t for t in ref.tracked_tracks + ref.lost_tracks}
    assert set(st) == set(rt) and len(st) > 8
    for k, (state, act, ext, mean, cov) in st.items():
        assert (state, act, ext) == (rt[k].state, rt[k].is_activated, rt[k].external_track_id)
        # 70 frames of float64 recursion with different (but equally valid) summation orders: ~1e-9 relative drift
        assert np.allclose(mean, rt[k].mean, rtol=1e-6, atol=1e-7) and np.allclose(cov, rt[k].covariance, rtol=1e-5, atol=1e-9)


def test_kalman_kernels_against_numpy_filter(cuda_dev):
    """initiate -> predict x3 -> update -> predict(zero vh) on 50 random boxes, against KalmanFilterRef."""
    from manual_yolo_b200 import _lib, api
    rng = np.random.default_rng(1)
    n = 50
    xy = rng.uniform(0, 1500, (n, 2)).astype(np.float32)
    wh = rng.uniform(20, 300, (n, 2)).astype(np.float32)
    boxes = np.concatenate([xy, xy + wh], 1)
    boxes2 = boxes + rng.normal(0, 3, boxes.shape).astype(np.float32)
    dev = cuda_dev
    mean = torch.zeros((n, 8), dtype=torch.float64, device=dev)
    cov = torch.zeros((n, 8, 8), dtype=torch.float64, device=dev)
    slots = torch.arange(n, dtype=torch.int32, device=dev)
    b1, b2 = torch.from_numpy(boxes).to(dev), torch.from_numpy(boxes2).to(dev)
    tlbr = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    lib = _lib.load()
    st = api._stream()
    assert lib.b200yolo_kalman_initiate(api._ptr(mean), api._ptr(cov), api._ptr(slots), api._ptr(b1), 4, api._ptr(slots), n, st) == 0
    zero = torch.zeros((n,), dtype=torch.int32, device=dev)
    for _ in range(3):
        assert lib.b200yolo_kalman_predict(api._ptr(mean), api._ptr(cov), api._ptr(slots), api._ptr(zero), n, api._ptr(tlbr), st) == 0
    assert lib.b200yolo_kalman_update(api._ptr(mean), api._ptr(cov), api._ptr(slots), api._ptr(b2), 4, api._ptr(slots), n, st) == 0
    one = torch.ones((n,), dtype=torch.int32, device=dev)
    assert lib.b200yolo_kalman_predict(api._ptr(mean), api._ptr(cov), api._ptr(slots), api._ptr(one), n, api._ptr(tlbr), st) == 0
    kf = obt.KalmanFilterRef()
    for i in range(n):
        mu, P = kf.initiate(obt.STrackRef.tlwh_to_xyah(obt.STrackRef.tlbr_to_tlwh(boxes[i].astype(np.float64))))
        for _ in range(3):
            mu, P = kf.predict(mu, P)
        mu, P = kf.update(mu, P, obt.STrackRef.tlwh_to_xyah(obt.STrackRef.tlbr_to_tlwh(boxes2[i].astype(np.float64))))
        mu[7] = 0
        mu, P = kf.predict(mu, P)
        assert np.allclose(mean[i].cpu().numpy(), mu, rtol=1e-11, atol=1e-11), i
        assert np.allclose(cov[i].cpu().numpy(), P, rtol=1e-9, atol=1e-12), i
        w = mu[2] * mu[3]
        exp = np.array([mu[0] - w / 2, mu[1] - mu[3] / 2, mu[0] - w / 2 + w, mu[1] - mu[3] / 2 + mu[3]], np.float32)
        assert np.allclose(tlbr[i].cpu().numpy(), exp, rtol=1e-6, atol=1e-4)


def test_bytetrack_empty_frames_and_capacity(cuda_dev):
    """Frames without detections (tracks go lost, then are removed after the buffer) follow the restatement; running out
    of track slots raises instead of dropping tracks silently."""
    frames = _sequence(n_frames=50, n_obj=6, seed=3)
    for f in list(range(12, 16)) + list(range(30, 50)):
        frames[f] = np.zeros((0, 6), np.float32)
    max_det = 32
    kw = dict(lost_track_buffer=8)
    trk = tracking.ByteTrack(device=cuda_dev, capacity=64, max_det=max_det, **kw)
    ref = obt.ByteTrackRef(**kw)
    for f, rows in enumerate(frames):
        n = rows.shape[0]
        pad = torch.zeros((1, max_det, 6))
        pad[0, :n] = torch.from_numpy(rows)
        det = m.Detections(pad.to(cuda_dev), torch.zeros((1, max_det), dtype=torch.int32, device=cuda_dev),
                           torch.tensor([n], dtype=torch.int32, device=cuda_dev))
        got = trk.update(det, 0)
        exp = ref.update_with_detections(rows[:, :4], rows[:, 4])
        assert got.tolist() == exp.tolist(), f
        assert len(trk.tracked) == len(ref.tracked_tracks) and len(trk.lost) == len(ref.lost_tracks), f
    assert not trk.tracked and not trk.lost                               # everything timed out during the empty tail
    assert len(trk._free) == 64                                           # ... and every slot came back
    small = tracking.ByteTrack(device=cuda_dev, capacity=3, max_det=max_det)
    rows = _sequence(n_frames=1, n_obj=10, seed=5)[0]
    rows[:, 4] = 0.9
    pad = torch.zeros((1, max_det, 6))
    pad[0, :rows.shape[0]] = torch.from_numpy(rows)
    det = m.Detections(pad.to(cuda_dev), torch.zeros((1, max_det), dtype=torch.int32, device=cuda_dev),
                       torch.tensor([rows.shape[0]], dtype=torch.int32, device=cuda_dev))
    with pytest.raises(RuntimeError, match="capacity"):
        small.update(det, 0)
