"""ORACLE (test infrastructure only) -- SAHI-style sliced prediction, SURVEY.md section 8(f) row N3.

The reference runs ``sahi.predict.get_sliced_prediction(frame, detection_model, slice_height=640,
slice_width=640, overlap_height_ratio=0.2, overlap_width_ratio=0.2)`` (``/root/reference/pipe.py:183-194``).
``sahi`` is a third-party dependency that is neither vendored in /root/reference nor installed here
(``requirements.txt`` does not pin it: ``pipe.py:27-30`` import-guards it), so this file restates its published
algorithm: PARITY UNPINNED.  What is restated:

* ``sahi.slicing.get_slice_bboxes``: the window layout (``get_slice_bboxes_ref``);
* per slice: the window is cut out of the frame (numpy view) and goes through the detector exactly like a
  frame (the existing oracle chain: letterbox -> head decode -> NMS -> scale_boxes to the slice shape);
* ``shift_amount``: the slice origin is added to the boxes;
* merge, SAHI's defaults (``pipe.py:186-193`` passes none of them): ``perform_standard_pred=True`` (the full-frame
  prediction is appended to the slice predictions), ``postprocess_type="GREEDYNMM"``, ``postprocess_match_metric="IOS"``,
  ``postprocess_match_threshold=0.5``, class-aware -- ``greedy_nmm_ref`` / ``greedy_nmm_postprocess_ref`` restate
  ``sahi/postprocess/combine.py`` (``greedy_nmm``, ``batched_greedy_nmm``, ``GreedyNMMPostprocess.__call__``) and
  ``sahi/postprocess/utils.py`` (``has_match``, ``calculate_bbox_ios/iou``, ``merge_object_prediction_pair``);
* merge, SURVEY's N3 row ("one more NMS"): ``merge="nms"`` with torchvision's NMS semantics.
"""
from __future__ import annotations

import torch
import torchvision

from . import boxes as oboxes
from . import head as ohead
from . import letterbox as olb
from . import nms as onms


def get_slice_bboxes_ref(image_height, image_width, slice_height=640, slice_width=640,
                         overlap_height_ratio=0.2, overlap_width_ratio=0.2):
    slice_bboxes = []
    y_max = y_min = 0
    y_overlap = int(overlap_height_ratio * slice_height)
    x_overlap = int(overlap_width_ratio * slice_width)
    while y_max < image_height:
        x_min = x_max = 0
        y_max = y_min + slice_height
        while x_max < image_width:
            x_max = x_min + slice_width
            if y_max > image_height or x_max > image_width:
                xmax = min(image_width, x_max)
                ymax = min(image_height, y_max)
                xmin = max(0, xmax - slice_width)
                ymin = max(0, ymax - slice_height)
                slice_bboxes.append([xmin, ymin, xmax, ymax])
            else:
                slice_bboxes.append([x_min, y_min, x_max, y_max])
            x_min = x_max - x_overlap
        y_min = y_max - y_overlap
    return slice_bboxes


def greedy_nmm_ref(preds: torch.Tensor, match_metric="IOS", match_threshold=0.5):
    """``sahi.postprocess.combine.greedy_nmm`` on an (n, 6) fp32 tensor [x1,y1,x2,y2,score,category]: returns
    {kept index: [merged indices, best first]} in the order boxes are kept.  Ties in score: ``argsort`` is made stable
    here (upstream's is not), i.e. of equal scores the later prediction is taken first."""
    keep_to_merge = {}
    x1, y1, x2, y2, scores = preds[:, 0], preds[:, 1], preds[:, 2], preds[:, 3], preds[:, 4]
    areas = (x2 - x1) * (y2 - y1)
    order = scores.argsort(stable=True)
    while len(order) > 0:
        idx = order[-1]
        order = order[:-1]
        if len(order) == 0:
            keep_to_merge[idx.tolist()] = []
            break
        xx1 = torch.max(torch.index_select(x1, 0, order), x1[idx])
        yy1 = torch.max(torch.index_select(y1, 0, order), y1[idx])
        xx2 = torch.min(torch.index_select(x2, 0, order), x2[idx])
        yy2 = torch.min(torch.index_select(y2, 0, order), y2[idx])
        w = torch.clamp(xx2 - xx1, min=0.0)
        h = torch.clamp(yy2 - yy1, min=0.0)
        inter = w * h
        rem_areas = torch.index_select(areas, 0, order)
        if match_metric == "IOU":
            value = inter / ((rem_areas - inter) + areas[idx])
        else:
            value = inter / torch.min(rem_areas, areas[idx])
        mask = value < match_threshold
        matched = order[(mask == False).nonzero().flatten()].flip(dims=(0,))     # noqa: E712 (upstream's spelling)
        unmatched = order[mask]
        order = unmatched[scores[unmatched].argsort(stable=True)]
        keep_to_merge[idx.tolist()] = matched.tolist()
    return keep_to_merge


def _metric64(b1, b2, match_metric):
    """``calculate_bbox_ios`` / ``calculate_bbox_iou`` of sahi/postprocess/utils.py: numpy float64 on python floats."""
    import numpy as np
    b1, b2 = np.array(b1[:4], dtype=np.float64), np.array(b2[:4], dtype=np.float64)
    a1, a2 = (b1[2] - b1[0]) * (b1[3] - b1[1]), (b2[2] - b2[0]) * (b2[3] - b2[1])
    wh = (np.minimum(b1[2:], b2[2:]) - np.maximum(b1[:2], b2[:2])).clip(min=0)
    inter = wh[0] * wh[1]
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / np.minimum(a1, a2) if match_metric == "IOS" else inter / (a1 + a2 - inter)


def greedy_nmm_postprocess_ref(preds: torch.Tensor, match_metric="IOS", match_threshold=0.5, class_agnostic=False):
    """``GreedyNMMPostprocess.__call__``: batched (per category) greedy NMM, then every queued prediction is merged into
    the kept one while ``has_match`` (metric > threshold, float64) holds against the current merged box: union box,
    max score, category of the higher score.  Returns (merged (k,6) fp32 in SAHI's order: category-major when
    class-aware, descending score inside; kept index per row)."""
    n = preds.shape[0]
    if n == 0:
        return torch.zeros((0, 6)), torch.zeros((0,), dtype=torch.int64)
    if class_agnostic:
        k2m = greedy_nmm_ref(preds, match_metric, match_threshold)
    else:
        k2m = {}
        cats = preds[:, 5]
        for cid in torch.unique(cats):
            ind = torch.where(cats == cid)[0]
            sub = greedy_nmm_ref(preds[ind], match_metric, match_threshold)
            for k, ml in sub.items():
                k2m[int(ind[k])] = [int(ind[j]) for j in ml]
    out, kept = [], []
    plist = [p.tolist() for p in preds]
    for k, ml in k2m.items():
        cur = list(plist[k])
        for j in ml:
            other = plist[j]
            if _metric64(cur, other, match_metric) > match_threshold:
                box = [min(cur[0], other[0]), min(cur[1], other[1]), max(cur[2], other[2]), max(cur[3], other[3])]
                cat = cur[5] if cur[4] > other[4] else other[5]
                cur = box + [max(cur[4], other[4]), cat]
        out.append(cur)
        kept.append(k)
    return torch.tensor(out, dtype=torch.float32).reshape(-1, 6), torch.tensor(kept, dtype=torch.int64)


def sliced_prediction_ref(frames, heads, slices, new_shape=(640, 640), conf=0.25, iou=0.7, merge_iou=0.5,
                          max_det=300, max_wh=7680, strides=(8, 16, 32), merge="nms", match_metric="IOS", heads_full=None):
    """frames: numpy (F,H,W,3) uint8; heads: torch (F*S, 64+nc, A), item f*S+s = slice s of frame f.
    Returns (net_in (F*S,3,h,w), per-slice detections in slice pixels, merged per-frame detections (k,6) in frame
    pixels, merged provenance (k,) = slice*max_det + rank)."""
    F, S = frames.shape[0], len(slices)
    crops = [frames[f][y0:y1, x0:x1] for f in range(F) for (x0, y0, x1, y1) in slices]
    net_in = olb.preprocess_ref(crops, new_shape)
    in_hw = tuple(net_in.shape[2:])
    level_hw = ohead.level_shapes(*in_hw, strides)
    pred = ohead.detect_inference_ref(heads, level_hw, strides)
    out = onms.non_max_suppression_ref(pred, conf, iou, max_det=max_det)
    per_slice, merged, prov = [], [], []
    for f in range(F):
        rows, ids = [], []
        for s, (x0, y0, x1, y1) in enumerate(slices):
            o = out[f * S + s].clone()
            o[:, :4] = oboxes.scale_boxes_ref(in_hw, o[:, :4], (y1 - y0, x1 - x0))
            per_slice.append(o.clone())
            o[:, :4] += torch.tensor([x0, y0, x0, y0], dtype=torch.float32)
            rows.append(o)
            ids.append(torch.arange(o.shape[0]) + s * max_det)
        if heads_full is not None:                     # perform_standard_pred: the full-frame prediction, appended last
            fin = olb.preprocess_ref([frames[f]], new_shape)
            fhw = tuple(fin.shape[2:])
            fo = onms.non_max_suppression_ref(ohead.detect_inference_ref(heads_full[f:f + 1], ohead.level_shapes(*fhw, strides),
                                                                         strides), conf, iou, max_det=max_det)[0].clone()
            fo[:, :4] = oboxes.scale_boxes_ref(fhw, fo[:, :4], frames[f].shape[:2])
            rows.append(fo)
            ids.append(torch.arange(fo.shape[0]) + len(slices) * max_det)
        x = torch.cat(rows)
        ids = torch.cat(ids)
        if x.shape[0] and merge == "greedy_nmm":
            x, kept = greedy_nmm_postprocess_ref(x, match_metric, merge_iou)
            ids = ids[kept]
        elif x.shape[0]:
            keep = torchvision.ops.nms(x[:, :4] + x[:, 5:6] * max_wh, x[:, 4], merge_iou)[:max_det]
            x, ids = x[keep], ids[keep]
        merged.append(x)
        prov.append(ids)
    return net_in, per_slice, merged, prov
