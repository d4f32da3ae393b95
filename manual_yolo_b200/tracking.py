"""ByteTrack over the path's detections: the tracker the reference builds with
``sv.ByteTrack()`` (``/root/reference/detect.py:22``) and feeds once per frame with
``tracker.update_with_detections(detections)`` (``detect.py:553-577``), same constructor arguments and defaults.

Split of the work:

* on the device, batched over tracks (``csrc/track.cu``, ``csrc/assoc.cu``): the Kalman prediction of every pooled
  track (+ its predicted box), the track x detection IoU cost matrix straight from the padded NMS output, the Kalman
  update of every matched track and the initiation of new ones -- the track states (mean (8,) / covariance (8,8),
  float64) never leave the device;
* on the host (this file): ByteTrack's two-stage association (high-score detections against tracked + lost tracks with
  score-fused costs, then low-score detections against the still unmatched tracked ones), the unconfirmed-track
  stage, the linear assignments (scipy) and the track lifecycle (activation, re-activation, loss, removal after
  ``lost_track_buffer`` frames, duplicate removal) -- small per-frame bookkeeping on a few dozen tracks.

``supervision==0.26.1`` (``requirements.txt:83``) is not installed in the build container: this follows its published
algorithm and is checked frame by frame against the numpy restatement ``oracle/bytetrack.py`` -- parity unpinned.
"""

from __future__ import annotations

from typing import List

import numpy as np
import torch

from . import _lib, api

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3
NO_ID = -1


class _Track:
    __slots__ = ("slot", "state", "is_activated", "start_frame", "frame_id", "tracklet_len", "score", "internal_id",
                 "external_id", "tlbr")

    def __init__(self, slot, score):
        self.slot, self.state, self.is_activated = slot, NEW, False
        self.start_frame = self.frame_id = self.tracklet_len = 0
        self.score, self.internal_id, self.external_id = score, NO_ID, NO_ID
        self.tlbr = None                                   # fp32 (4,) host copy of the current box (end of last frame)


def _box_iou_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """supervision ``box_iou_batch`` in fp32 (host side: duplicate removal and the final id hand-out on <= ~100 boxes)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    tl = np.maximum(a[:, None, :2], b[:, :2])
    br = np.minimum(a[:, None, 2:], b[:, 2:])
    inter = np.prod(np.clip(br - tl, a_min=0, a_max=None), 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / (area_a[:, None] + area_b - inter)


def linear_assignment(cost: np.ndarray, thresh: float):
    """supervision ``matching.linear_assignment``: costs above ``thresh`` are clipped to ``thresh + 1e-4``, minimum-cost
    assignment, pairs costlier than ``thresh`` are unmatched.  Returns (matches (k,2), unmatched rows, unmatched cols)."""
    from scipy.optimize import linear_sum_assignment
    if cost.size == 0:
        return np.empty((0, 2), dtype=int), tuple(range(cost.shape[0])), tuple(range(cost.shape[1]))
    c = cost.copy()
    c[c > thresh] = thresh + 1e-4
    r, k = linear_sum_assignment(c)
    ok = c[r, k] <= thresh
    m = np.stack([r[ok], k[ok]], 1).astype(int)
    return m, tuple(sorted(set(range(c.shape[0])) - set(m[:, 0].tolist()))), tuple(sorted(set(range(c.shape[1])) - set(m[:, 1].tolist())))


class ByteTrack:
    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: int = 30, minimum_consecutive_frames: int = 1,
                 device="cuda", capacity: int = 1024, max_det: int = 300):
        if not torch.cuda.is_available():
            raise RuntimeError("manual_yolo_b200.tracking.ByteTrack needs a CUDA device (no CPU path)")
        self.track_activation_threshold = track_activation_threshold
        self.minimum_matching_threshold = minimum_matching_threshold
        self.det_thresh = track_activation_threshold + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self.frame_id = 0
        self.device = torch.device(device)
        self.capacity, self.max_det = int(capacity), int(max_det)
        self.mean = torch.zeros((self.capacity, 8), dtype=torch.float64, device=self.device)
        self.cov = torch.zeros((self.capacity, 8, 8), dtype=torch.float64, device=self.device)
        self._free = list(range(self.capacity - 1, -1, -1))
        self._boxes = torch.zeros((1, self.capacity, 4), dtype=torch.float32, device=self.device)   # predicted / current tlbr
        self._cost = torch.empty((1, self.capacity, self.max_det), dtype=torch.float32, device=self.device)
        self._tcount = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self.tracked: List[_Track] = []
        self.lost: List[_Track] = []
        self.removed: List[_Track] = []
        self._next_internal, self._next_external = 0, 1

    # ---- device helpers --------------------------------------------------------------------------------------
    def _i32(self, values):
        return torch.tensor(list(values), dtype=torch.int32, device=self.device)

    def _predict(self, tracks: List[_Track], modes: List[int]):
        """modes: 0 predict, 1 zero vh then predict (not Tracked), 2 no prediction (unconfirmed tracks / read-out).
        One launch; the boxes (tlbr, fp32) of the given tracks land in rows [0, n) of the track box buffer."""
        n = len(tracks)
        if n == 0:
            return
        slots, md = self._i32(t.slot for t in tracks), self._i32(modes)
        _lib.check(_lib.load().b200yolo_kalman_predict(api._ptr(self.mean), api._ptr(self.cov), api._ptr(slots), api._ptr(md), n,
                                                       api._ptr(self._boxes), api._stream()), "kalman_predict")

    def _kf(self, fn, tracks: List[_Track], det_rows: torch.Tensor, det_idx: List[int]):
        if not tracks:
            return
        slots, idx = self._i32(t.slot for t in tracks), self._i32(det_idx)
        _lib.check(fn(api._ptr(self.mean), api._ptr(self.cov), api._ptr(slots), api._ptr(det_rows), int(det_rows.stride(0)),
                      api._ptr(idx), len(tracks), api._stream()), "kalman")

    # ---- lifecycle (host) ------------------------------------------------------------------------------------
    def _activate(self, trk: _Track):
        trk.internal_id = self._next_internal
        self._next_internal += 1
        trk.tracklet_len, trk.state = 0, TRACKED
        if self.frame_id == 1:
            trk.is_activated = True
        if self.minimum_consecutive_frames == 1:
            trk.external_id = self._next_external
            self._next_external += 1
        trk.frame_id = trk.start_frame = self.frame_id

    def _update(self, trk: _Track, score: float):
        trk.frame_id = self.frame_id
        trk.tracklet_len += 1
        trk.state = TRACKED
        if trk.tracklet_len == self.minimum_consecutive_frames:
            trk.is_activated = True
            if trk.external_id == NO_ID:
                trk.external_id = self._next_external
                self._next_external += 1
        trk.score = score

    def _reactivate(self, trk: _Track, score: float):
        trk.tracklet_len, trk.state, trk.frame_id, trk.score = 0, TRACKED, self.frame_id, score

    @staticmethod
    def _joint(a, b):
        seen, out = set(), []
        for t in list(a) + list(b):
            if t.internal_id not in seen:
                seen.add(t.internal_id)
                out.append(t)
        return out

    @staticmethod
    def _sub(a, b):
        ids = {t.internal_id for t in b}
        return [t for t in a if t.internal_id not in ids]

    # ---- one frame ---------------------------------------------------------------------------------------------
    def update(self, det: api.Detections, b: int = 0) -> np.ndarray:
        """Feed frame ``b`` of a padded NMS output; returns ``tracker_id`` per detection row of that frame (-1 where
        supervision would drop the detection: no confirmed track matched it) -- ``update_with_detections``."""
        rows = det.rows[b]                                               # (max_det, 6) on the device
        if rows.shape[0] != self.max_det:
            raise ValueError(f"tracker was built for max_det={self.max_det}")
        n_det = int(det.count[b])
        scores = rows[:n_det, 4].cpu().numpy().astype(np.float32)        # one small device->host read
        self.frame_id += 1
        activated, refind, lost, removed = [], [], [], []
        remain = np.nonzero(scores > np.float32(self.track_activation_threshold))[0]
        second = np.nonzero((scores > np.float32(0.1)) & (scores < np.float32(self.track_activation_threshold)))[0]
        unconfirmed = [t for t in self.tracked if not t.is_activated]
        tracked = [t for t in self.tracked if t.is_activated]
        pool = self._joint(tracked, self.lost)
        everyone = pool + unconfirmed
        self._predict(everyone, [0 if t.state == TRACKED else 1 for t in pool] + [2] * len(unconfirmed))
        T = len(everyone)
        cost = np.zeros((T, n_det), np.float32)
        if T and n_det:
            self._tcount.fill_(T)
            out = self._cost.view(-1)[:T * self.max_det].view(1, T, self.max_det)
            api.iou_cost_matrix(self._boxes[:, :T], self._tcount, api.Detections(det.rows[b:b + 1], det.anchor[b:b + 1],
                                                                                 det.count[b:b + 1]), fuse_score=False, out=out)
            cost = out[0, :, :n_det].cpu().numpy()                       # (tracks x detections) 1 - IoU, fp32
        P = len(pool)

        def fused(c, cols):                                              # matching.fuse_score, fp32
            if c.size == 0:
                return c
            return (np.float32(1) - (np.float32(1) - c) * scores[cols][None, :]).astype(np.float32)
        upd_trk, upd_det = [], []
        # first association: pooled tracks x high-score detections, score-fused costs
        m, u_track, u_det = linear_assignment(fused(cost[:P][:, remain], remain), self.minimum_matching_threshold)
        for it, idet in m:
            trk, d = pool[it], int(remain[idet])
            (activated if trk.state == TRACKED else refind).append(trk)
            (self._update if trk.state == TRACKED else self._reactivate)(trk, float(scores[d]))
            upd_trk.append(trk); upd_det.append(d)
        # second association: still unmatched TRACKED tracks x low-score detections, plain IoU costs
        r_idx = [i for i in u_track if pool[i].state == TRACKED]
        m, u_track2, _ = linear_assignment(cost[r_idx][:, second] if r_idx else np.zeros((0, len(second)), np.float32), 0.5)
        for it, idet in m:
            trk, d = pool[r_idx[it]], int(second[idet])
            (activated if trk.state == TRACKED else refind).append(trk)
            (self._update if trk.state == TRACKED else self._reactivate)(trk, float(scores[d]))
            upd_trk.append(trk); upd_det.append(d)
        for it in u_track2:
            trk = pool[r_idx[it]]
            if trk.state != LOST:
                trk.state = LOST
                lost.append(trk)
        # unconfirmed tracks (one frame old) x the high-score detections left over
        left = remain[list(u_det)] if len(u_det) else np.zeros((0,), int)
        m, u_unc, u_det2 = linear_assignment(fused(cost[P:][:, left], left), 0.7)
        for it, idet in m:
            trk, d = unconfirmed[it], int(left[idet])
            self._update(trk, float(scores[d]))
            activated.append(trk)
            upd_trk.append(trk); upd_det.append(d)
        for it in u_unc:
            unconfirmed[it].state = REMOVED
            removed.append(unconfirmed[it])
        lib = _lib.load()
        self._kf(lib.b200yolo_kalman_update, upd_trk, rows, upd_det)
        # new tracks from the remaining high-score detections
        new_trk, new_det = [], []
        for inew in u_det2:
            d = int(left[inew])
            if scores[d] < np.float32(self.det_thresh):
                continue
            if not self._free:
                raise RuntimeError("tracker capacity exhausted")
            trk = _Track(self._free.pop(), float(scores[d]))
            self._activate(trk)
            activated.append(trk)
            new_trk.append(trk); new_det.append(d)
        self._kf(lib.b200yolo_kalman_initiate, new_trk, rows, new_det)
        for trk in self.lost:
            if self.frame_id - trk.frame_id > self.max_time_lost:
                trk.state = REMOVED
                removed.append(trk)
        self.tracked = [t for t in self.tracked if t.state == TRACKED]
        self.tracked = self._joint(self.tracked, activated)
        self.tracked = self._joint(self.tracked, refind)
        self.lost = self._sub(self.lost, self.tracked)
        self.lost.extend(lost)
        self.lost = self._sub(self.lost, self.removed)
        for trk in self.removed:                                         # slots of the tracks removed a frame ago are free again
            if trk.slot is not None and trk.state == REMOVED:
                self._free.append(trk.slot)
                trk.slot = None
        self.removed = removed
        # current boxes of every live track (one small read), duplicate removal, id hand-out
        live = self.tracked + self.lost
        if live:
            self._predict(live, [2] * len(live))
            boxes = self._boxes[0, :len(live)].cpu().numpy()
            for t, bx in zip(live, boxes):
                t.tlbr = bx
        kept_a, kept_b = self._remove_duplicates(self.tracked, self.lost)
        for t in self.tracked + self.lost:                               # duplicates vanish: their slots are free again
            if t not in kept_a and t not in kept_b and t.slot is not None:
                self._free.append(t.slot)
                t.slot = None
        self.tracked, self.lost = kept_a, kept_b
        out = [t for t in self.tracked if t.is_activated]
        tracker_id = np.full(n_det, -1, dtype=int)
        if out and n_det:
            det_boxes = rows[:n_det, :4].cpu().numpy()
            ious = _box_iou_f32(det_boxes, np.stack([t.tlbr for t in out]))
            m, _, _ = linear_assignment((np.float32(1) - ious).astype(np.float32), 0.5)
            for i_det, i_trk in m:
                tracker_id[i_det] = out[i_trk].external_id
        return tracker_id

    @staticmethod
    def _remove_duplicates(a, b):
        if not a or not b:
            return a, b
        pdist = (np.float32(1) - _box_iou_f32(np.stack([t.tlbr for t in a]), np.stack([t.tlbr for t in b]))).astype(np.float32)
        dupa, dupb = set(), set()
        for p, q in zip(*np.where(pdist < 0.15)):
            if a[p].frame_id - a[p].start_frame > b[q].frame_id - b[q].start_frame:
                dupb.add(int(q))
            else:
                dupa.add(int(p))
        return [t for i, t in enumerate(a) if i not in dupa], [t for i, t in enumerate(b) if i not in dupb]

    def states(self):
        """{internal id: (state, is_activated, external id, mean (8,), covariance (8,8))} of the live tracks (debug /
        tests; one device->host read)."""
        live = self.tracked + self.lost
        if not live:
            return {}
        idx = torch.tensor([t.slot for t in live], device=self.device)
        mean, cov = self.mean[idx].cpu().numpy(), self.cov[idx].cpu().numpy()
        return {t.internal_id: (t.state, t.is_activated, t.external_id, mean[i], cov[i]) for i, t in enumerate(live)}
