"""Oracle: YOLOv8 Detect-head decode.  TEST INFRASTRUCTURE ONLY.

Reference entry: ``/root/reference/detect.py:541`` -> upstream
``ultralytics/nn/modules/head.py::Detect._inference``, ``block.py::DFL.forward``,
``utils/tal.py::make_anchors, dist2bbox`` (ultralytics==8.3.176; restated from SURVEY.md
Appendix A.5-A.7 with torch CPU ops in the same order -- parity unpinned).
"""

from __future__ import annotations

import torch
import torch.nn.functional as F

REG_MAX = 16


def level_shapes(in_h: int, in_w: int, strides=(8, 16, 32)):
    """Feature-map (h, w) per level for a letterboxed input of (in_h, in_w)."""
    return [(in_h // s, in_w // s) for s in strides]


def make_anchors_ref(level_hw, strides=(8, 16, 32), grid_cell_offset=0.5):
    """tal.make_anchors: returns anchor_points (A,2) [x,y] and stride_tensor (A,1)."""
    anchor_points, stride_tensor = [], []
    for (h, w), stride in zip(level_hw, strides):
        sx = torch.arange(end=w, dtype=torch.float32) + grid_cell_offset
        sy = torch.arange(end=h, dtype=torch.float32) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        anchor_points.append(torch.stack((sx, sy), -1).view(-1, 2))
        stride_tensor.append(torch.full((h * w, 1), stride, dtype=torch.float32))
    return torch.cat(anchor_points), torch.cat(stride_tensor)


def dfl_ref(box: torch.Tensor) -> torch.Tensor:
    """DFL.forward: (B, 64, A) -> (B, 4, A); softmax over 16 bins then 1x1 conv with arange weights."""
    b, _, a = box.shape
    w = torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)
    return F.conv2d(box.view(b, 4, REG_MAX, a).transpose(2, 1).softmax(1), w).view(b, 4, a)


def dist2bbox_ref(distance, anchor_points, xywh=True, dim=1):
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        c_xy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        return torch.cat((c_xy, wh), dim)
    return torch.cat((x1y1, x2y2), dim)


def detect_inference_ref(head: torch.Tensor, level_hw, strides=(8, 16, 32)) -> torch.Tensor:
    """Detect._inference: x_cat (B, 64+nc, A) fp32 -> y (B, 4+nc, A): xywh px + sigmoid scores."""
    head = head.float()
    nc = head.shape[1] - 4 * REG_MAX
    anchors, stride_t = make_anchors_ref(level_hw, strides, 0.5)
    anchors, stride_t = anchors.transpose(0, 1), stride_t.transpose(0, 1)
    box, cls = head.split((4 * REG_MAX, nc), 1)
    dbox = dist2bbox_ref(dfl_ref(box), anchors.unsqueeze(0), xywh=True, dim=1) * stride_t
    return torch.cat((dbox, cls.sigmoid()), 1)


def cat_levels(levels):
    """Detect._inference head: cat([xi.view(B, no, -1) for xi in x], 2)."""
    b, no = levels[0].shape[:2]
    return torch.cat([xi.reshape(b, no, -1) for xi in levels], 2)
