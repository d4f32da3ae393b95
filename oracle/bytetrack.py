"""ORACLE (test infrastructure only) -- supervision's ByteTrack, SURVEY.md section 8(f) row N2.

The reference builds ``tracker = sv.ByteTrack()`` (``/root/reference/detect.py:22``) and calls
``tracker.update_with_detections(detections)`` once per frame (``detect.py:557``).  ``supervision==0.26.1``
(``requirements.txt:83``) is a third-party dependency that is neither vendored in /root/reference nor installed here;
this file restates its published algorithm in numpy, keeping upstream's structure and names
(``supervision/tracker/byte_tracker/{core,kalman_filter,matching,single_object_track}.py``): PARITY UNPINNED.
All Kalman arithmetic is float64 here (upstream mixes float32 measurements into float64 matrices).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.optimize import linear_sum_assignment

from .assoc import box_iou_batch_ref

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3


class KalmanFilterRef:
    def __init__(self):
        ndim, dt = 4, 1.0
        self._motion_mat = np.eye(2 * ndim, 2 * ndim)
        for i in range(ndim):
            self._motion_mat[i, ndim + i] = dt
        self._update_mat = np.eye(ndim, 2 * ndim)
        self._std_weight_position = 1.0 / 20
        self._std_weight_velocity = 1.0 / 160

    def initiate(self, measurement):
        mean_pos = np.asarray(measurement, np.float64)
        mean = np.r_[mean_pos, np.zeros_like(mean_pos)]
        std = [2 * self._std_weight_position * measurement[3], 2 * self._std_weight_position * measurement[3], 1e-2,
               2 * self._std_weight_position * measurement[3], 10 * self._std_weight_velocity * measurement[3],
               10 * self._std_weight_velocity * measurement[3], 1e-5, 10 * self._std_weight_velocity * measurement[3]]
        return mean, np.diag(np.square(np.asarray(std, np.float64)))

    def predict(self, mean, covariance):
        std_pos = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-2,
                   self._std_weight_position * mean[3]]
        std_vel = [self._std_weight_velocity * mean[3], self._std_weight_velocity * mean[3], 1e-5,
                   self._std_weight_velocity * mean[3]]
        motion_cov = np.diag(np.square(np.r_[std_pos, std_vel]))
        mean = np.dot(mean, self._motion_mat.T)
        covariance = np.linalg.multi_dot((self._motion_mat, covariance, self._motion_mat.T)) + motion_cov
        return mean, covariance

    def project(self, mean, covariance):
        std = [self._std_weight_position * mean[3], self._std_weight_position * mean[3], 1e-1,
               self._std_weight_position * mean[3]]
        innovation_cov = np.diag(np.square(std))
        mean = np.dot(self._update_mat, mean)
        covariance = np.linalg.multi_dot((self._update_mat, covariance, self._update_mat.T))
        return mean, covariance + innovation_cov

    def update(self, mean, covariance, measurement):
        projected_mean, projected_cov = self.project(mean, covariance)
        chol_factor, lower = cho_factor(projected_cov, lower=True, check_finite=False)
        kalman_gain = cho_solve((chol_factor, lower), np.dot(covariance, self._update_mat.T).T, check_finite=False).T
        innovation = np.asarray(measurement, np.float64) - projected_mean
        new_mean = mean + np.dot(innovation, kalman_gain.T)
        new_covariance = covariance - np.linalg.multi_dot((kalman_gain, projected_cov, kalman_gain.T))
        return new_mean, new_covariance


class IdCounter:
    NO_ID = -1

    def __init__(self, start_id=0):
        self._id = start_id

    def new_id(self):
        v = self._id
        self._id += 1
        return v


class STrackRef:
    def __init__(self, tlwh, score, minimum_consecutive_frames, kf, internal_ids, external_ids):
        self.state, self.is_activated, self.start_frame, self.frame_id = NEW, False, 0, 0
        self._tlwh = np.asarray(tlwh, dtype=np.float32)
        self.kf, self.mean, self.covariance = kf, None, None
        self.score, self.tracklet_len = score, 0
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self.internal_ids, self.external_ids = internal_ids, external_ids
        self.internal_track_id, self.external_track_id = IdCounter.NO_ID, IdCounter.NO_ID

    @staticmethod
    def tlbr_to_tlwh(tlbr):
        ret = np.asarray(tlbr).copy()
        ret[2:] -= ret[:2]
        return ret

    @staticmethod
    def tlwh_to_xyah(tlwh):
        ret = np.asarray(tlwh, np.float64).copy()
        ret[:2] += ret[2:] / 2
        ret[2] /= ret[3]
        return ret

    @property
    def tlwh(self):
        if self.mean is None:
            return self._tlwh.copy()
        ret = self.mean[:4].copy()
        ret[2] *= ret[3]
        ret[:2] -= ret[2:] / 2
        return ret

    @property
    def tlbr(self):
        ret = self.tlwh.copy()
        ret[2:] += ret[:2]
        return ret

    def predict_multi(self):
        if self.state != TRACKED:
            self.mean[7] = 0
        self.mean, self.covariance = self.kf.predict(self.mean, self.covariance)

    def activate(self, frame_id):
        self.internal_track_id = self.internal_ids.new_id()
        self.mean, self.covariance = self.kf.initiate(self.tlwh_to_xyah(self._tlwh))
        self.tracklet_len, self.state = 0, TRACKED
        if frame_id == 1:
            self.is_activated = True
        if self.minimum_consecutive_frames == 1:
            self.external_track_id = self.external_ids.new_id()
        self.frame_id = self.start_frame = frame_id

    def re_activate(self, new_track, frame_id):
        self.mean, self.covariance = self.kf.update(self.mean, self.covariance, self.tlwh_to_xyah(new_track.tlwh))
        self.tracklet_len, self.state, self.frame_id, self.score = 0, TRACKED, frame_id, new_track.score

    def update(self, new_track, frame_id):
        self.frame_id = frame_id
        self.tracklet_len += 1
        self.mean, self.covariance = self.kf.update(self.mean, self.covariance, self.tlwh_to_xyah(new_track.tlwh))
        self.state = TRACKED
        if self.tracklet_len == self.minimum_consecutive_frames:
            self.is_activated = True
            if self.external_track_id == IdCounter.NO_ID:
                self.external_track_id = self.external_ids.new_id()
        self.score = new_track.score


def iou_distance(atracks, btracks):
    if not atracks or not btracks:
        return np.zeros((len(atracks), len(btracks)), dtype=np.float32)
    a = np.asarray([t.tlbr for t in atracks], np.float32)
    b = np.asarray([t.tlbr for t in btracks], np.float32)
    return (np.float32(1) - box_iou_batch_ref(a, b).astype(np.float32)).astype(np.float32)


def fuse_score(cost_matrix, detections):
    if cost_matrix.size == 0:
        return cost_matrix
    iou_sim = 1 - cost_matrix
    det_scores = np.array([d.score for d in detections], np.float32)
    det_scores = np.expand_dims(det_scores, axis=0).repeat(cost_matrix.shape[0], axis=0)
    return (1 - iou_sim * det_scores).astype(np.float32)


def linear_assignment(cost_matrix, thresh):
    if cost_matrix.size == 0:
        return np.empty((0, 2), dtype=int), tuple(range(cost_matrix.shape[0])), tuple(range(cost_matrix.shape[1]))
    cost_matrix = cost_matrix.copy()
    cost_matrix[cost_matrix > thresh] = thresh + 1e-4
    row_ind, col_ind = linear_sum_assignment(cost_matrix)
    matched = cost_matrix[row_ind, col_ind] <= thresh
    matches = np.stack([row_ind[matched], col_ind[matched]], 1).astype(int)
    ua = tuple(sorted(set(range(cost_matrix.shape[0])) - set(matches[:, 0].tolist())))
    ub = tuple(sorted(set(range(cost_matrix.shape[1])) - set(matches[:, 1].tolist())))
    return matches, ua, ub


def joint_tracks(a, b):
    seen, out = set(), []
    for t in list(a) + list(b):
        if t.internal_track_id not in seen:
            seen.add(t.internal_track_id)
            out.append(t)
    return out


def sub_tracks(a, b):
    ids = {t.internal_track_id for t in b}
    return [t for t in a if t.internal_track_id not in ids]


def remove_duplicate_tracks(a, b):
    pdist = iou_distance(a, b)
    pairs = np.where(pdist < 0.15)
    dupa, dupb = [], []
    for p, q in zip(*pairs):
        timep = a[p].frame_id - a[p].start_frame
        timeq = b[q].frame_id - b[q].start_frame
        if timep > timeq:
            dupb.append(q)
        else:
            dupa.append(p)
    return [t for i, t in enumerate(a) if i not in dupa], [t for i, t in enumerate(b) if i not in dupb]


class ByteTrackRef:
    def __init__(self, track_activation_threshold=0.25, lost_track_buffer=30, minimum_matching_threshold=0.8, frame_rate=30,
                 minimum_consecutive_frames=1):
        self.track_activation_threshold = track_activation_threshold
        self.minimum_matching_threshold = minimum_matching_threshold
        self.frame_id = 0
        self.det_thresh = track_activation_threshold + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.minimum_consecutive_frames = minimum_consecutive_frames
        self.kf = KalmanFilterRef()
        self.tracked_tracks, self.lost_tracks, self.removed_tracks = [], [], []
        self.internal_ids, self.external_ids = IdCounter(), IdCounter(start_id=1)

    def _mk(self, tlbr, score):
        return STrackRef(STrackRef.tlbr_to_tlwh(tlbr), score, self.minimum_consecutive_frames, self.kf, self.internal_ids,
                         self.external_ids)

    def update_with_detections(self, xyxy, confidence):
        """Returns tracker_id per detection (-1: no confirmed track), as ``detections.tracker_id`` before supervision drops
        the unmatched ones."""
        xyxy = np.asarray(xyxy, np.float32).reshape(-1, 4)
        confidence = np.asarray(confidence, np.float32).reshape(-1)
        tensors = np.hstack((xyxy, confidence[:, None]))
        tracks = self.update_with_tensors(tensors)
        tracker_id = np.full(len(xyxy), -1, dtype=int)
        if len(tracks) > 0 and len(xyxy) > 0:
            ious = box_iou_batch_ref(xyxy, np.asarray([t.tlbr for t in tracks], np.float32))
            matches, _, _ = linear_assignment((1 - ious).astype(np.float32), 0.5)
            for i_det, i_trk in matches:
                tracker_id[i_det] = int(tracks[i_trk].external_track_id)
        return tracker_id

    def update_with_tensors(self, tensors):
        self.frame_id += 1
        activated, refind, lost, removed = [], [], [], []
        scores, bboxes = tensors[:, 4], tensors[:, :4]
        remain = scores > self.track_activation_threshold
        second = np.logical_and(scores > 0.1, scores < self.track_activation_threshold)
        detections = [self._mk(b, s) for b, s in zip(bboxes[remain], scores[remain])]
        unconfirmed = [t for t in self.tracked_tracks if not t.is_activated]
        tracked = [t for t in self.tracked_tracks if t.is_activated]
        pool = joint_tracks(tracked, self.lost_tracks)
        for t in pool:
            t.predict_multi()
        dists = fuse_score(iou_distance(pool, detections), detections)
        matches, u_track, u_det = linear_assignment(dists, self.minimum_matching_threshold)
        for it, idet in matches:
            trk, det = pool[it], detections[idet]
            if trk.state == TRACKED:
                trk.update(det, self.frame_id)
                activated.append(trk)
            else:
                trk.re_activate(det, self.frame_id)
                refind.append(trk)
        detections_second = [self._mk(b, s) for b, s in zip(bboxes[second], scores[second])]
        r_tracked = [pool[i] for i in u_track if pool[i].state == TRACKED]
        matches, u_track2, _ = linear_assignment(iou_distance(r_tracked, detections_second), 0.5)
        for it, idet in matches:
            trk, det = r_tracked[it], detections_second[idet]
            if trk.state == TRACKED:
                trk.update(det, self.frame_id)
                activated.append(trk)
            else:
                trk.re_activate(det, self.frame_id)
                refind.append(trk)
        for it in u_track2:
            trk = r_tracked[it]
            if trk.state != LOST:
                trk.state = LOST
                lost.append(trk)
        detections = [detections[i] for i in u_det]
        dists = fuse_score(iou_distance(unconfirmed, detections), detections)
        matches, u_unconfirmed, u_det = linear_assignment(dists, 0.7)
        for it, idet in matches:
            unconfirmed[it].update(detections[idet], self.frame_id)
            activated.append(unconfirmed[it])
        for it in u_unconfirmed:
            unconfirmed[it].state = REMOVED
            removed.append(unconfirmed[it])
        for inew in u_det:
            trk = detections[inew]
            if trk.score < self.det_thresh:
                continue
            trk.activate(self.frame_id)
            activated.append(trk)
        for trk in self.lost_tracks:
            if self.frame_id - trk.frame_id > self.max_time_lost:
                trk.state = REMOVED
                removed.append(trk)
        self.tracked_tracks = [t for t in self.tracked_tracks if t.state == TRACKED]
        self.tracked_tracks = joint_tracks(self.tracked_tracks, activated)
        self.tracked_tracks = joint_tracks(self.tracked_tracks, refind)
        self.lost_tracks = sub_tracks(self.lost_tracks, self.tracked_tracks)
        self.lost_tracks.extend(lost)
        self.lost_tracks = sub_tracks(self.lost_tracks, self.removed_tracks)
        self.removed_tracks = removed
        self.tracked_tracks, self.lost_tracks = remove_duplicate_tracks(self.tracked_tracks, self.lost_tracks)
        return [t for t in self.tracked_tracks if t.is_activated]
