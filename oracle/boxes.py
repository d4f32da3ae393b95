"""Oracle: box rescale to source pixels and integer crop geometry.

TEST INFRASTRUCTURE ONLY.  ``scale_boxes``/``clip_boxes`` restate upstream
``ultralytics/utils/ops.py`` (ultralytics==8.3.176; parity unpinned).
``safe_crop_ref`` follows ``/root/reference/detect.py:100-113`` (pad=6 at ``detect.py:586``);
the consumers' ``int()`` truncation is ``detect.py:581`` / ``yolo.py:369`` / ``pipe.py:115``.
"""

from __future__ import annotations

import torch


def scale_boxes_ref(img1_shape, boxes: torch.Tensor, img0_shape, ratio_pad=None, padding=True):
    """Rescale xyxy boxes from letterboxed (img1) to source (img0) pixels, in place on a clone."""
    boxes = boxes.clone()
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
               round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    if padding:
        boxes[..., 0] -= pad[0]
        boxes[..., 1] -= pad[1]
        boxes[..., 2] -= pad[0]
        boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    return clip_boxes_ref(boxes, img0_shape)


def clip_boxes_ref(boxes: torch.Tensor, shape):
    boxes[..., 0].clamp_(0, shape[1])
    boxes[..., 1].clamp_(0, shape[0])
    boxes[..., 2].clamp_(0, shape[1])
    boxes[..., 3].clamp_(0, shape[0])
    return boxes


def safe_crop_box_ref(frame_hw, x1, y1, x2, y2, pad=6):
    """detect.py:100-113 on already int()-truncated coords.  Returns (x1,y1,x2,y2) or None."""
    h, w = frame_hw
    x1 = max(0, min(w - 1, int(x1 - pad)))
    x2 = max(0, min(w, int(x2 + pad)))
    y1 = max(0, min(h - 1, int(y1 - pad)))
    y2 = max(0, min(h, int(y2 + pad)))
    if x2 <= x1 or y2 <= y1:
        return None
    return x1, y1, x2, y2


def safe_crop_ref(frame, x1, y1, x2, y2, pad=6):
    box = safe_crop_box_ref(frame.shape[:2], x1, y1, x2, y2, pad)
    if box is None:
        return None
    x1, y1, x2, y2 = box
    return frame[y1:y2, x1:x2]
