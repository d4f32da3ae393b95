"""Oracle: letterbox + normalise.  TEST INFRASTRUCTURE ONLY.

Reference entry: ``/root/reference/detect.py:541`` ``model(frame)``,
``yolo.py:361``, ``pipe.py:179`` -> upstream ``ultralytics/data/augment.py::LetterBox.__call__``
and ``ultralytics/engine/predictor.py::BasePredictor.preprocess`` (ultralytics==8.3.176,
``requirements.txt:95``; restated from SURVEY.md Appendix A.3/A.4 -- parity unpinned,
see ``oracle/__init__.py``).  The pixel arithmetic is done by the real ``cv2`` leaf.
"""

from __future__ import annotations

import cv2
import numpy as np
import torch


def letterbox_geometry(shape_hw, new_shape=(640, 640), auto=False, scale_fill=False,
                       scaleup=True, center=True, stride=32):
    """LetterBox.__call__ geometry (Appendix A.3).  Returns dict with python ints."""
    h, w = int(shape_hw[0]), int(shape_hw[1])
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(w * r)), int(round(h * r))  # (w', h'), python round = half-to-even
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = dw % stride, dh % stride
    elif scale_fill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return dict(new_w=new_unpad[0], new_h=new_unpad[1], top=top, bottom=bottom, left=left,
                right=right, out_h=new_unpad[1] + top + bottom, out_w=new_unpad[0] + left + right,
                ratio=r)


def letterbox_ref(img, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True,
                  center=True, stride=32, padding_value=114):
    """HWC BGR u8 -> letterboxed HWC BGR u8 through the real cv2 leaves."""
    g = letterbox_geometry(img.shape[:2], new_shape, auto, scale_fill, scaleup, center, stride)
    h, w = img.shape[:2]
    if (w, h) != (g["new_w"], g["new_h"]):
        img = cv2.resize(img, (g["new_w"], g["new_h"]), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, g["top"], g["bottom"], g["left"], g["right"],
                              cv2.BORDER_CONSTANT, value=(padding_value,) * 3)


def preprocess_ref(frames, new_shape=(640, 640), auto=False, stride=32):
    """BasePredictor.preprocess (Appendix A.4): list/array of HWC BGR u8 -> (B,3,H,W) fp32."""
    im = np.stack([letterbox_ref(f, new_shape, auto=auto, stride=stride) for f in frames])
    im = im[..., ::-1].transpose((0, 3, 1, 2))
    im = np.ascontiguousarray(im)
    t = torch.from_numpy(im).float()
    t /= 255
    return t


# --------------------------------------------------------------------------------------------
# numpy restatement of cv2.resize(INTER_LINEAR) on uint8.
# It is the *specification* the CUDA kernel follows; tests check it bit-for-bit against cv2.
# --------------------------------------------------------------------------------------------

def cv2_linear_axis_table(ssize: int, dsize: int, vertical: bool = False):
    """Per-axis (index0, index1, a0, a1) of cv2's 8-bit INTER_LINEAR: 11-bit coefficients.

    Horizontal axis: taps outside the image are CLAMPED WITH THE FRACTION RESET (``fx = 0``).
    Vertical axis: cv::resize keeps the fractional weights and only clamps the source ROW indices
    (``clip(sy + k, 0, ssize)`` in the row loop), so on up-scales the first/last rows blend the same
    source row twice with two separately truncated products.  Identical for every down-scale."""
    scale = 1.0 / (dsize / ssize)  # double, exactly as cv::resize: inv_scale = dsize/ssize; scale = 1/inv
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0.0
        s[hi] = ssize - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int32)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int32)
    s0 = np.clip(s, 0, ssize - 1)
    s1 = np.clip(s + 1, 0, ssize - 1)
    return s0, s1, a0, a1


def cv2_resize_linear_restated(img, dsize_wh):
    """Bit-exact restatement of cv2.resize(img, dsize, INTER_LINEAR) on u8 HWC (down- and up-scales)."""
    H, W = img.shape[:2]
    dw, dh = int(dsize_wh[0]), int(dsize_wh[1])
    sx, sx1, ax0, ax1 = cv2_linear_axis_table(W, dw)
    sy, sy1, ay0, ay1 = cv2_linear_axis_table(H, dh, vertical=True)
    src = img.astype(np.int32)
    # horizontal pass on the two referenced rows
    r0 = src[sy]            # (dh, W, C)
    r1 = src[sy1]
    S0 = r0[:, sx] * ax0[None, :, None] + r0[:, sx1] * ax1[None, :, None]
    S1 = r1[:, sx] * ax0[None, :, None] + r1[:, sx1] * ax1[None, :, None]
    b0 = ay0[:, None, None]
    b1 = ay1[:, None, None]
    out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_restated(img, new_shape=(640, 640), auto=False, stride=32, padding_value=114):
    """letterbox_ref with the numpy fixed-point resize instead of cv2 (self-check helper)."""
    g = letterbox_geometry(img.shape[:2], new_shape, auto=auto, stride=stride)
    h, w = img.shape[:2]
    if (w, h) != (g["new_w"], g["new_h"]):
        img = cv2_resize_linear_restated(img, (g["new_w"], g["new_h"]))
    out = np.full((g["out_h"], g["out_w"], 3), padding_value, np.uint8)
    out[g["top"]:g["top"] + g["new_h"], g["left"]:g["left"] + g["new_w"]] = img
    return out
