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


# ---- N1: rank-classifier hand-off (detect.py:115-139 classify_card_rank, :60-98 normalize_rank_text) --------------
VALID_CARD_RANKS = frozenset({"A", "K", "Q", "J", "10", "9", "8", "7", "6", "5", "4", "3", "2"})   # detect.py:36
_LOOKALIKE = {"O": "0", "I": "1", "S": "5", "Z": "2", "B": "8", "T": "10"}                          # detect.py:37
_NUMERIC_RANKS = frozenset(str(v) for v in range(2, 11))


def normalize_rank_text(text: str) -> str:
    """Classifier / OCR text -> one of A,K,Q,J,10,9..2, or "" (behaviour of ``detect.py:60-98``, pinned by
    ``tests/golden/rank_text_golden.json``, which was produced by executing the reference's own function)."""
    if not text:
        return ""
    t = text.strip().upper()
    if len(t) == 1:
        t = _LOOKALIKE.get(t, t)                       # single look-alike letters first
    t = t.replace(" ", "").replace("|", "1").replace("O", "0")
    if t == "T":
        t = "10"
    if t in ("A", "K", "Q", "J"):
        return t
    if t.isdigit():
        if t == "0":                                   # a lone zero is a ten that lost its one
            t = "10"
        if t in _NUMERIC_RANKS:
            return t
    if len(t) == 1 and t in _LOOKALIKE:                # last chance for a single look-alike
        m = "10" if _LOOKALIKE[t] == "0" else _LOOKALIKE[t]
        if m in _NUMERIC_RANKS:
            return m
    return ""


def rank_text_from_top1(pred_name: str, top1conf: float, det_class_name: str = "") -> str:
    """``classify_card_rank``'s decision (``detect.py:115-139``): accept the classifier's top-1 when its confidence
    reaches 0.20 for turn/river cards and 0.40 otherwise; cleaned rank if valid, else the upper-cased name."""
    low = det_class_name.lower()
    threshold = 0.20 if ("turn" in low or "river" in low) else 0.40
    if top1conf >= threshold:
        cleaned = normalize_rank_text(pred_name)
        return cleaned if cleaned in VALID_CARD_RANKS else pred_name.upper()
    return ""


def classify_rank_rois(result, forward_logits, rank_names: Optional[dict] = None,
                       det_names: Optional[dict] = None) -> List[dict]:
    """Batched form of the reference's per-crop loop (``detect.py:580-588`` -> ``:121-131``): ONE forward of the
    rank classifier over the K5 batch of a step (``result``: a ``PipelineResult``), one device->host read of
    (top1, top1conf), then the reference's thresholds and text clean-up.  ``forward_logits``: callable mapping the
    (n,3,64,64) fp32 ROI tensor to (n,13) logits (the YOLOv8n-cls network stays torch) -- e.g. a
    ``classifier.RankClassifier`` from ``classifier.load_rank_classifier("rank_classifier.pt")`` (``detect.py:21``),
    whose ``names`` are used when ``rank_names`` is not given.  Returns one dict per valid ROI: frame, det (row in that
    frame's detections), class_id, top1, top1conf, text."""
    if rank_names is None:
        rank_names = getattr(forward_logits, "names", None)
        if rank_names is None:
            raise ValueError("rank_names is required when the classifier carries no names")
    n = min(int(result.roi_count), int(result.rois.shape[0]))
    if n == 0:
        return []
    probs = forward_logits(result.rois[:n]).softmax(1)
    conf, top1 = probs.max(1)
    conf, top1 = conf.cpu().tolist(), top1.cpu().tolist()
    frames, dets, valid = result.roi_batch[:n].cpu().tolist(), result.roi_det[:n].cpu().tolist(), result.roi_valid[:n].cpu().tolist()
    cls = result.det.rows[result.roi_batch[:n].long(), result.roi_det[:n].long(), 5].cpu().tolist()
    out = []
    for i in range(n):
        if valid[i] <= 0:                               # safe_crop returned None: classify_card_rank returns ""
            continue
        cid = int(cls[i])
        dname = det_names.get(cid, f"class{cid}") if det_names else f"class{cid}"
        out.append({"frame": frames[i], "det": dets[i], "class_id": cid, "top1": top1[i], "top1conf": conf[i],
                    "text": rank_text_from_top1(rank_names.get(top1[i], ""), conf[i], dname)})
    return out


def to_pipe_records(det_rows, det_count, names: Optional[dict], image_shape) -> List[List[dict]]:
    """Per-frame detection dicts as ``pipe.py``'s ``parse_ultralytics_results`` builds them (``pipe.py:100-134``):
    ``int()``-truncated coordinates, then ``x1,y1`` clamped at 0 and ``x2,y2`` at ``w-1`` / ``h-1``; ``conf`` float,
    ``class_id`` int, ``class_name`` from ``names`` (``class<id>`` when missing).  Pinned by
    ``tests/golden/pipe_records_golden.json`` (produced by executing the reference's function)."""
    rows = det_rows.tolist() if hasattr(det_rows, "tolist") else det_rows
    counts = det_count.tolist() if hasattr(det_count, "tolist") else det_count
    h, w = int(image_shape[0]), int(image_shape[1])
    out = []
    for b, n in enumerate(counts):
        recs = []
        for i in range(n):
            x1, y1, x2, y2, conf, cls = rows[b][i]
            cid = int(cls)
            recs.append({"x1": max(0, int(x1)), "y1": max(0, int(y1)), "x2": min(w - 1, int(x2)), "y2": min(h - 1, int(y2)),
                         "conf": float(conf), "class_id": cid,
                         "class_name": names.get(cid, f"class{cid}") if names else f"class{cid}"})
        out.append(recs)
    return out
