"""Oracle: ROI crop -> 64x64 classifier batch.  TEST INFRASTRUCTURE ONLY.

Reference entry: ``/root/reference/detect.py:121`` ``rank_model(crop)`` -> upstream
``ultralytics/models/yolo/classify/predict.py::ClassificationPredictor.preprocess`` which applies the
checkpoint's stored transforms ``Compose(Resize(64, bilinear, antialias), CenterCrop(64), ToTensor(),
Normalize(0, 1))``.  ``classify_preprocess_ref`` runs the REAL
``torchvision.transforms`` + PIL leaves; ``pil_resize_restated`` is the numpy restatement of Pillow's
8-bit two-pass ``ImagingResample`` (Appendix B.4) that the CUDA kernel follows.  This leg is pinned by
the classifier known-answer test (63/67, ``runs/rank_classifier/results.csv:21``).
"""

from __future__ import annotations

import math

import cv2
import numpy as np
import torch
import torchvision.transforms as T
from PIL import Image

PRECISION_BITS = 32 - 8 - 2


def stored_transforms(size=64):
    """The Compose pickled inside rank_classifier.pt (verified equal in tests/test_oracle_kat.py)."""
    return T.Compose([
        T.Resize(size, interpolation=T.InterpolationMode.BILINEAR),
        T.CenterCrop((size, size)),
        T.ToTensor(),
        T.Normalize(mean=torch.tensor([0.0, 0.0, 0.0]), std=torch.tensor([1.0, 1.0, 1.0])),
    ])


_TF = None


def classify_preprocess_ref(crop_bgr: np.ndarray, size=64) -> torch.Tensor:
    """One BGR HWC u8 crop -> (3,size,size) fp32, through the real PIL/torchvision leaves."""
    global _TF
    if _TF is None or _TF[0] != size:
        _TF = (size, stored_transforms(size))
    return _TF[1](Image.fromarray(cv2.cvtColor(np.ascontiguousarray(crop_bgr), cv2.COLOR_BGR2RGB)))


def resize_target(w: int, h: int, size=64):
    """torchvision Resize(int) output (new_w, new_h): short side -> size, long -> int(size*long/short)."""
    if w <= h:
        return size, int(size * h / w)
    return int(size * w / h), size


def center_crop_offsets(w: int, h: int, size=64):
    """torchvision CenterCrop: (left, top) with python banker's rounding."""
    return int(round((w - size) / 2.0)), int(round((h - size) / 2.0))


def pil_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(xmax, np.float64)
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            v = -v if v < 0.0 else v
            wv = 1.0 - v if v < 1.0 else 0.0
            w[x] = wv
            ww += wv
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(xmax):
            kk[xx, x] = int(-0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] < 0 else int(0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample along axis 0 of an (N, ..., ) uint8 array with Pillow's 8-bit fixed point."""
    bounds, kk = pil_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(xmax):
            acc += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_restated(img_rgb: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """Image.resize((new_w,new_h), BILINEAR) on HWC u8: horizontal pass, round to u8, vertical pass."""
    h, w = img_rgb.shape[:2]
    out = img_rgb
    if new_w != w:
        out = _resample_axis0(out.transpose(1, 0, 2), new_w).transpose(1, 0, 2)
    if new_h != h:
        out = _resample_axis0(out, new_h)
    return np.ascontiguousarray(out)


def classify_preprocess_restated_u8(crop_bgr: np.ndarray, size=64) -> np.ndarray:
    """Restated pipeline up to (but excluding) /255: returns (size,size,3) RGB u8."""
    rgb = crop_bgr[..., ::-1]
    h, w = rgb.shape[:2]
    new_w, new_h = resize_target(w, h, size)
    if (new_w, new_h) != (w, h):
        rgb = pil_resize_restated(rgb, new_w, new_h)
    left, top = center_crop_offsets(new_w, new_h, size)
    return np.ascontiguousarray(rgb[top:top + size, left:left + size])


def classify_preprocess_restated(crop_bgr: np.ndarray, size=64) -> torch.Tensor:
    u8 = classify_preprocess_restated_u8(crop_bgr, size)
    return torch.from_numpy(u8).permute(2, 0, 1).contiguous().float().div(255)
