"""ORACLE (test infrastructure only) -- tracker association costs, SURVEY.md section 8(f) row N2.

The reference hands every frame's detections to supervision's ByteTrack (``/root/reference/detect.py:557``
``tracker.update_with_detections``).  ``supervision`` is a third-party dependency that is not vendored in
/root/reference and not installed here (``requirements.txt:83`` pins ``supervision==0.26.1``); this file restates
the published association arithmetic of ``supervision/detection/utils`` ``box_iou_batch`` and
``supervision/tracker/byte_tracker/matching.py`` (``iou_distance``, ``fuse_score``): PARITY UNPINNED.
"""
from __future__ import annotations

import numpy as np


def box_iou_batch_ref(boxes_true: np.ndarray, boxes_detection: np.ndarray) -> np.ndarray:
    def box_area(box):
        return (box[2] - box[0]) * (box[3] - box[1])
    area_true = box_area(boxes_true.T)
    area_detection = box_area(boxes_detection.T)
    top_left = np.maximum(boxes_true[:, None, :2], boxes_detection[:, :2])
    bottom_right = np.minimum(boxes_true[:, None, 2:], boxes_detection[:, 2:])
    area_inter = np.prod(np.clip(bottom_right - top_left, a_min=0, a_max=None), 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        return area_inter / (area_true[:, None] + area_detection - area_inter)


def iou_cost_ref(tracks_xyxy: np.ndarray, det_rows: np.ndarray, fuse_score=False) -> np.ndarray:
    """``1 - iou`` (``iou_distance``), optionally fused with the detection scores (``fuse_score``); fp32."""
    t = np.asarray(tracks_xyxy, np.float32)
    d = np.asarray(det_rows, np.float32)
    cost = np.float32(1) - box_iou_batch_ref(t, d[:, :4]).astype(np.float32)
    if fuse_score:
        cost = np.float32(1) - (np.float32(1) - cost) * d[None, :, 4]
    return cost.astype(np.float32)
