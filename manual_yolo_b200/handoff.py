"""Host-side hand-off formats after the path -- bookkeeping, no device work.

N2: ``sv.Detections`` arrays.  The reference turns ``Results`` into ``sv.Detections`` and cleans them
(``detect.py:253-310`` ``create_clean_detections``: xyxy float32, class_id int32 with None/NaN -> 0, confidence
float32 with None/NaN -> 0.0, tracker_id int32 with None/NaN -> -1) before ``tracker.update_with_detections``
(``detect.py:542-563``).  ``to_tracker_arrays`` produces the same arrays per frame from the padded device
output in ONE device->host read for the whole batch.  The tracker itself (supervision ByteTrack / DeepSORT,
third-party, not installed here) is outside the path.

N4: JSON emission.  The reference re-dumps the whole detection history with ``json.dump(all_detections, f,
indent=2)`` after every frame (``detect.py:679-690``) -- O(frames^2) bytes written.  ``JsonlWriter`` appends one
line per frame, once per batch, with the same per-frame object (``frame``, ``timestamp``, ``detections`` of
``frame_data`` records, ``detect.py:590-598``); ``load_jsonl_as_reference_list`` restores the reference's list.
"""

from __future__ import annotations

import json
import math
import time
from typing import Dict, List, Optional

import numpy as np


def to_tracker_arrays(det_rows, det_count, tracker_ids=None) -> List[Dict[str, np.ndarray]]:
    """Padded (B,max_det,6) rows + (B,) counts (device or host tensors / arrays) -> per-frame dicts with the
    ``sv.Detections`` field names: ``xyxy`` (n,4) float32, ``confidence`` (n,) float32, ``class_id`` (n,) int32,
    ``tracker_id`` (n,) int32 or None.  Cleaning rules of ``create_clean_detections``: NaN class -> 0, NaN
    confidence -> 0.0, NaN / None tracker id -> -1; an empty frame gives empty arrays (``sv.Detections.empty()``)."""
    rows = det_rows.detach().cpu().numpy() if hasattr(det_rows, "detach") else np.asarray(det_rows)
    counts = det_count.detach().cpu().numpy() if hasattr(det_count, "detach") else np.asarray(det_count)
    out = []
    for b, n in enumerate(counts.tolist()):
        r = rows[b, :n]
        cls = r[:, 5]
        conf = r[:, 4]
        d = {"xyxy": np.ascontiguousarray(r[:, :4], dtype=np.float32),
             "confidence": np.where(np.isnan(conf), 0.0, conf).astype(np.float32),
             "class_id": np.where(np.isnan(cls), 0, cls).astype(np.int32),
             "tracker_id": None}
        if tracker_ids is not None:
            t = tracker_ids[b]
            clean = [(-1 if (v is None or (isinstance(v, float) and math.isnan(v))) else int(v)) for v in list(t)[:n]]
            d["tracker_id"] = np.asarray(clean, dtype=np.int32)
        out.append(d)
    return out


def frame_objects(det_rows, det_count, names: Optional[dict] = None, frame_offset=0, tracker_ids=None, ocr_texts=None,
                  timestamp: Optional[float] = None) -> List[dict]:
    """One ``{"frame", "timestamp", "detections": [frame_data...]}`` object per frame, as ``detect.py:679-683``
    appends them; bbox ints are ``int()``-truncated (``detect.py:581``), conf rounded to 3 places
    (``detect.py:596``)."""
    rows = det_rows.tolist() if hasattr(det_rows, "tolist") else det_rows
    counts = det_count.tolist() if hasattr(det_count, "tolist") else det_count
    ts = time.time() if timestamp is None else timestamp
    out = []
    for b, n in enumerate(counts):
        dets = []
        for i in range(n):
            x1, y1, x2, y2, conf, cls = rows[b][i]
            cid = int(cls)
            dets.append({"frame": frame_offset + b,
                         "tracker_id": int(tracker_ids[b][i]) if tracker_ids is not None else -1,
                         "class_id": cid,
                         "class_name": names.get(cid, f"class{cid}") if names else f"class{cid}",
                         "bbox": [int(x1), int(y1), int(x2), int(y2)],
                         "conf": round(float(conf), 3),
                         "ocr_text": ocr_texts[b][i] if ocr_texts is not None else ""})
        out.append({"frame": frame_offset + b, "timestamp": ts, "detections": dets})
    return out


class JsonlWriter:
    """Append-only JSON Lines emitter: one line per frame, one write per batch (O(frames) bytes in total)."""

    def __init__(self, path: str):
        self.path = path
        self.f = open(path, "a", encoding="utf-8")
        self.frames = 0
        self.bytes = 0

    def write_batch(self, frame_objs: List[dict]):
        blob = "".join(json.dumps(o, separators=(",", ":")) + "\n" for o in frame_objs)
        self.f.write(blob)
        self.f.flush()
        self.frames += len(frame_objs)
        self.bytes += len(blob)

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def load_jsonl_as_reference_list(path: str) -> List[dict]:
    """The list the reference's ``json.dump(all_detections, ...)`` holds, rebuilt from the JSONL file."""
    with open(path, encoding="utf-8") as f:
        return [json.loads(line) for line in f if line.strip()]


def associate(cost: np.ndarray, thresh: float):
    """Linear assignment of tracks (rows) to detections (columns) on a cost matrix from ``api.iou_cost_matrix``
    (already cut to the valid (n_tracks, n_dets) block): minimum-cost matching, pairs costlier than ``thresh`` are
    left unmatched -- the role of ``matching.linear_assignment(cost, thresh)`` in supervision's ByteTrack (which uses
    ``lap.lapjv(cost_limit=thresh)``; scipy's solver is what this image has).
    Returns (matches (k,2) int, unmatched_tracks, unmatched_dets)."""
    from scipy.optimize import linear_sum_assignment
    cost = np.asarray(cost, np.float64)
    if cost.size == 0:
        return np.zeros((0, 2), int), list(range(cost.shape[0])), list(range(cost.shape[1]))
    safe = np.where(np.isfinite(cost), cost, 1e6)
    r, c = linear_sum_assignment(safe)
    ok = safe[r, c] <= thresh
    matches = np.stack([r[ok], c[ok]], 1).astype(int)
    ut = sorted(set(range(cost.shape[0])) - set(matches[:, 0].tolist()))
    ud = sorted(set(range(cost.shape[1])) - set(matches[:, 1].tolist()))
    return matches, ut, ud
