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
* merge: one more class-aware NMS over the frame.  SAHI offers ``postprocess_type`` "NMS" and "GREEDYNMM"
  (its default, which *merges* matched boxes instead of dropping them); the build follows SURVEY's N3 row
  ("one more NMS") with torchvision's NMS semantics (``>`` against the threshold, SAHI's own loop uses ``>=``).
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


def sliced_prediction_ref(frames, heads, slices, new_shape=(640, 640), conf=0.25, iou=0.7, merge_iou=0.5,
                          max_det=300, max_wh=7680, strides=(8, 16, 32)):
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
        x = torch.cat(rows)
        ids = torch.cat(ids)
        if x.shape[0]:
            keep = torchvision.ops.nms(x[:, :4] + x[:, 5:6] * max_wh, x[:, 4], merge_iou)[:max_det]
            x, ids = x[keep], ids[keep]
        merged.append(x)
        prov.append(ids)
    return net_in, per_slice, merged, prov
