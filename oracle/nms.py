"""Oracle: candidate filter + class-aware NMS.  TEST INFRASTRUCTURE ONLY.

Reference entry: ``/root/reference/detect.py:541`` / ``yolo.py:361`` / ``pipe.py:179`` -> upstream
``ultralytics/utils/ops.py::non_max_suppression`` (ultralytics==8.3.176, ``requirements.txt:95``)
-> ``torchvision.ops.nms`` (``requirements.txt:89``).  The glue is restated from SURVEY.md Appendix
A.8; the suppression itself runs in the REAL ``torchvision.ops.nms`` CPU kernel.  Parity unpinned
(no golden detections exist in the reference).  Intentional divergence: the upstream wall-clock
``time_limit`` early break is not replicated.
"""

from __future__ import annotations

import numpy as np
import torch
import torchvision


def xywh2xyxy_ref(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression_ref(prediction, conf_thres=0.25, iou_thres=0.45, classes=None,
                            agnostic=False, multi_label=False, labels=(), max_det=300, nc=0,
                            max_nms=30000, max_wh=7680, return_idxs=False, nms_fn=None):
    """UL ``ops.non_max_suppression`` for the detect task on a (B, 4+nc, A) fp32 CPU tensor."""
    assert 0 <= conf_thres <= 1
    assert 0 <= iou_thres <= 1
    assert not multi_label and not labels, "oracle covers the reference's call pattern only"
    nms_fn = nms_fn or torchvision.ops.nms
    prediction = prediction.detach().cpu().float().clone()
    if classes is not None:
        classes = torch.tensor(classes)
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    extra = prediction.shape[1] - nc - 4
    mi = 4 + nc
    xc = prediction[:, 4:mi].amax(1) > conf_thres
    xinds = torch.stack([torch.arange(len(i)) for i in xc])[..., None]
    prediction = prediction.transpose(-1, -2)
    prediction[..., :4] = xywh2xyxy_ref(prediction[..., :4])
    output = [torch.zeros((0, 6 + extra))] * bs
    keepi = [torch.zeros((0,), dtype=torch.int64)] * bs
    for xi, (x, xk) in enumerate(zip(prediction, xinds)):
        filt = xc[xi]
        x, xk = x[filt], xk[filt]
        if not x.shape[0]:
            continue
        box, cls, mask = x.split((4, nc, extra), 1)
        conf, j = cls.max(1, keepdim=True)
        filt = conf.view(-1) > conf_thres
        x = torch.cat((box, conf, j.float(), mask), 1)[filt]
        xk = xk[filt]
        if classes is not None:
            filt = (x[:, 5:6] == classes).any(1)
            x, xk = x[filt], xk[filt]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            filt = x[:, 4].argsort(descending=True)[:max_nms]
            x, xk = x[filt], xk[filt]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        scores = x[:, 4]
        boxes = x[:, :4] + c
        i = nms_fn(boxes, scores, iou_thres)
        i = i[:max_det]
        output[xi], keepi[xi] = x[i], xk[i].reshape(-1)
    return (output, keepi) if return_idxs else output


def filter_candidates_ref(prediction, conf_thres=0.25, classes=None, nc=0):
    """The a7 stage alone: per image (n,6) xyxy/conf/cls candidates in anchor order + anchor idx."""
    prediction = prediction.detach().cpu().float().clone()
    nc = nc or (prediction.shape[1] - 4)
    out = []
    for p in prediction:
        cls = p[4:4 + nc].transpose(0, 1)
        conf, j = cls.max(1)
        keep = conf > conf_thres
        if classes is not None:
            keep &= (j[:, None] == torch.tensor(classes)[None]).any(1)
        idx = keep.nonzero().view(-1)
        box = xywh2xyxy_ref(p[:4].transpose(0, 1))[idx]
        out.append((torch.cat((box, conf[idx, None], j[idx, None].float()), 1), idx))
    return out


def nms_numpy_restated(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """fp32 restatement of torchvision's CPU ``nms_kernel``.

    Stable descending sort; greedy; ``ovr = inter / (ai + aj - inter)`` in fp32, compared with the
    *double* threshold.  This is the specification the CUDA NMS follows; tests check it
    index-for-index against ``torchvision.ops.nms``.
    """
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = (boxes[order, k] for k in range(4))
    areas = (x2 - x1) * (y2 - y1)
    suppressed = np.zeros(n, bool)
    keep = []
    thr = np.float64(iou_thres)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            if suppressed[i]:
                continue
            keep.append(order[i])
            if i + 1 == n:
                break
            xx1 = np.maximum(x1[i], x1[i + 1:])
            yy1 = np.maximum(y1[i], y1[i + 1:])
            xx2 = np.minimum(x2[i], x2[i + 1:])
            yy2 = np.minimum(y2[i], y2[i + 1:])
            w = np.maximum(np.float32(0), xx2 - xx1)
            h = np.maximum(np.float32(0), yy2 - yy1)
            inter = w * h
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
            suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, np.int64)
