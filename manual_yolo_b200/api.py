"""Python host side of the B200 detection path: same names / kwargs as the Ultralytics functions the
reference reaches through ``model(frame)`` (``/root/reference/detect.py:541``, ``yolo.py:361``,
``pipe.py:179``) and ``rank_model(crop)`` (``detect.py:121``), bound to the hand-written sm_100a
kernels through the C ABI in ``include/b200yolo.h``.

PyTorch is only the owner of device memory and streams here.  Every function takes CUDA tensors and
raises on CPU tensors or unsupported modes -- there is no CPU fallback.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib, geometry

REG_MAX = 16


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (manual_yolo_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must have dtype {dtype}, got {t.dtype}")


def _frames_4d(frames: torch.Tensor, allow_pinned=False):
    """``allow_pinned``: K5 may read the frames straight from PINNED host memory (zero-copy over PCIe:
    under unified addressing the device uses the same pointer); every other stage needs CUDA tensors."""
    if allow_pinned and isinstance(frames, torch.Tensor) and not frames.is_cuda:
        if not frames.is_pinned() or frames.dtype != torch.uint8:
            raise ValueError("host frames must be a pinned uint8 tensor (tensor.pin_memory()); there is no CPU path")
    else:
        _require_cuda(frames, "frames", torch.uint8)
    squeeze = frames.dim() == 3
    if squeeze:
        frames = frames.unsqueeze(0)
    if frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("frames must be (H,W,3) or (B,H,W,3) uint8 BGR")
    if frames.stride(-1) != 1 or frames.stride(-2) != 3:
        if not frames.is_cuda:
            raise ValueError("pinned host frames must have interleaved BGR pixels (strides (.., 3, 1))")
        frames = frames.contiguous()
    return frames, squeeze


def _class_mask(classes, nc, device):
    """Allow-list -> (ceil(nc/32),) int32 device tensor of bit words (None = all classes)."""
    if classes is None:
        return None
    if isinstance(classes, torch.Tensor):      # prebuilt mask (Pipeline caches it: graph-capture safe)
        return classes
    words = geometry.class_mask_words(classes, nc)
    return torch.tensor([w - (1 << 32) if w >= (1 << 31) else w for w in words], dtype=torch.int32, device=device)


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
def letterbox(image, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True,
              stride=32, padding_value=114, out=None):
    """``LetterBox.__call__`` drop-in: (H,W,3) or (B,H,W,3) uint8 BGR -> letterboxed uint8, same layout."""
    frames, squeeze = _frames_4d(image)
    B, H, W, _ = frames.shape
    g = geometry.letterbox_geometry((H, W), new_shape, auto, scale_fill, scaleup, center, stride)
    if out is None:
        out = torch.empty((B, g["out_h"], g["out_w"], 3), dtype=torch.uint8, device=frames.device)
    lib = _lib.load()
    rc = lib.b200yolo_letterbox_u8(_ptr(frames), B, H, W, frames.stride(1), frames.stride(0), _ptr(out),
                                   g["out_h"], g["out_w"], g["new_w"], g["new_h"], g["top"], g["left"],
                                   int(padding_value), _stream())
    _lib.check(rc, "letterbox")
    return out[0] if squeeze else out


def stage_rows_h2d(frames_host: torch.Tensor, out: Optional[torch.Tensor] = None, new_shape=(640, 640), auto=False,
                   scale_fill=False, scaleup=True, center=True, stride=32, device="cuda"):
    """H2D copy of only the source rows the letterbox resize reads (``geometry.referenced_rows``): for
    1920x1200 -> 640x400 one row in three, i.e. a third of the PCIe bytes of ``frames.to(device)``.
    ``frames_host``: pinned (B,H,W,3) uint8.  Returns the (B,n_rows,W,3) device tensor to hand to
    ``preprocess(..., src_hw=(H,W))``; asynchronous on the current stream."""
    if not isinstance(frames_host, torch.Tensor) or frames_host.is_cuda or not frames_host.is_pinned():
        raise ValueError("frames_host must be a pinned host tensor")
    if frames_host.dtype != torch.uint8 or frames_host.dim() != 4 or frames_host.shape[-1] != 3 \
            or not frames_host.is_contiguous():
        raise ValueError("frames_host must be a contiguous (B,H,W,3) uint8 tensor")
    B, H, W, _ = frames_host.shape
    g = geometry.letterbox_geometry((H, W), new_shape, auto, scale_fill, scaleup, center, stride)
    row0, step, n = geometry.referenced_rows(H, g["new_h"])
    if out is None:
        out = torch.empty((B, n, W, 3), dtype=torch.uint8, device=device)
    elif tuple(out.shape) != (B, n, W, 3) or out.dtype != torch.uint8 or not out.is_cuda or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous CUDA uint8 tensor of shape {(B, n, W, 3)}")
    rc = _lib.load().b200yolo_stage_rows_h2d(_ptr(frames_host), B, H, W, frames_host.stride(1), frames_host.stride(0),
                                             row0, step, n, _ptr(out), out.stride(1), out.stride(0), _stream())
    _lib.check(rc, "stage_rows_h2d")
    return out


def preprocess(frames, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True,
               stride=32, padding_value=114, out=None, src_hw=None, half=False):
    """``BasePredictor.preprocess`` drop-in, fused: uint8 BGR frames -> (B,3,H,W) fp32 RGB in [0,1]
    (``half=True``: float16, the ``predict(half=True)`` form -- ``im.half(); im /= 255``).

    ``src_hw``: when given, ``frames`` holds only the rows ``geometry.referenced_rows(src_hw[0], new_h)``
    names (staged by ``stage_rows_h2d``); the letterbox geometry is that of the full (H,W) source and the
    kernel runs with a vertical scale of 1 -- bit-identical output (tests/test_gpu_letterbox.py)."""
    frames, _ = _frames_4d(frames)
    B, H, W, _ = frames.shape
    if src_hw is not None:
        if int(src_hw[1]) != W:
            raise ValueError("staged rows must keep the source width")
        g = geometry.letterbox_geometry(src_hw, new_shape, auto, scale_fill, scaleup, center, stride)
        n = geometry.referenced_rows(src_hw[0], g["new_h"])[2]
        if H != n:
            raise ValueError(f"staged rows: expected {n} rows per frame, got {H}")
        if n != src_hw[0] and n != g["new_h"]:
            raise ValueError("staged rows must be the full frame or exactly new_h rows")
    else:
        g = geometry.letterbox_geometry((H, W), new_shape, auto, scale_fill, scaleup, center, stride)
    dt = torch.float16 if half else torch.float32
    if out is None:
        out = torch.empty((B, 3, g["out_h"], g["out_w"]), dtype=dt, device=frames.device)
    elif tuple(out.shape) != (B, 3, g["out_h"], g["out_w"]) or out.dtype != dt or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous (B,3,out_h,out_w) {dt} tensor")
    lib = _lib.load()
    fn = lib.b200yolo_letterbox_u8_to_f16 if half else lib.b200yolo_letterbox_u8_to_f32
    rc = fn(_ptr(frames), B, H, W, frames.stride(1), frames.stride(0), _ptr(out), g["out_h"], g["out_w"], g["new_w"],
            g["new_h"], g["top"], g["left"], int(padding_value), 1, _stream())
    _lib.check(rc, "preprocess")
    return out


def _slice_table(slices):
    """[[x0,y0,x1,y1],...] (equal sizes) -> (ctypes int array of origins, n, slice_h, slice_w)."""
    if not slices:
        raise ValueError("no slices")
    sw, sh = slices[0][2] - slices[0][0], slices[0][3] - slices[0][1]
    if any(s[2] - s[0] != sw or s[3] - s[1] != sh for s in slices):
        raise ValueError("slices must all have the same size (geometry.slice_boxes guarantees it)")
    flat = (ctypes.c_int * (2 * len(slices)))(*[v for s in slices for v in (int(s[0]), int(s[1]))])
    return flat, len(slices), sh, sw


def preprocess_slices(frames, slices, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True,
                      stride=32, padding_value=114, out=None):
    """K1 in slice mode: the SAHI-style front end (``pipe.py:183-194``).  ``frames`` (F,H,W,3) uint8 BGR on the
    device, ``slices`` from ``geometry.slice_boxes``; item ``f * n_slices + s`` of the result is slice ``s`` of
    frame ``f`` letterboxed + normalised exactly as ``preprocess`` would treat that window as an image.
    Returns (F * n_slices, 3, out_h, out_w) fp32; one launch for the whole batch."""
    frames, _ = _frames_4d(frames)
    F, H, W, _ = frames.shape
    xy, ns, sh, sw = _slice_table(slices)
    g = geometry.letterbox_geometry((sh, sw), new_shape, auto, scale_fill, scaleup, center, stride)
    if out is None:
        out = torch.empty((F * ns, 3, g["out_h"], g["out_w"]), dtype=torch.float32, device=frames.device)
    elif tuple(out.shape) != (F * ns, 3, g["out_h"], g["out_w"]) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous (F*n_slices,3,out_h,out_w) float32 tensor")
    rc = _lib.load().b200yolo_letterbox_slices_u8_to_f32(_ptr(frames), F, H, W, frames.stride(1), frames.stride(0), xy,
                                                         ns, sh, sw, _ptr(out), g["out_h"], g["out_w"], g["new_w"],
                                                         g["new_h"], g["top"], g["left"], int(padding_value), 1,
                                                         _stream())
    _lib.check(rc, "preprocess_slices")
    return out


def gather_slice_detections(det: "Detections", slices, n_frames, out: Optional["Candidates"] = None,
                            full_det: Optional["Detections"] = None) -> "Candidates":
    """Merge input of sliced prediction: the kept detections of each frame's slices (``det`` over F * n_slices
    items, boxes in slice pixels), concatenated slice-major and shifted by the slice origins, as candidates
    for the merge (``greedy_nmm`` -- SAHI's default -- or one more NMS, ``nms_candidates``).  ``anchor`` of a candidate
    = slice * max_det + rank.  ``full_det``: detections of the full frames (frame pixels), appended after the slices
    with ``anchor`` = n_slices * max_det + rank -- SAHI's ``perform_standard_pred=True`` (its default, ``pipe.py:186``)."""
    xy, ns, _, _ = _slice_table(slices)
    max_det = det.rows.shape[1]
    if det.rows.shape[0] != n_frames * ns:
        raise ValueError("det must hold n_frames * n_slices items")
    if full_det is not None and tuple(full_det.rows.shape[:2]) != (n_frames, max_det):
        raise ValueError("full_det must hold n_frames items with the same max_det")
    cap = (ns + (1 if full_det is not None else 0)) * max_det
    cands = _alloc_candidates(n_frames, cap, det.rows.device, out)
    rc = _lib.load().b200yolo_gather_slice_detections(
        _ptr(det.rows), _ptr(det.count), n_frames, ns, max_det, xy, _ptr(full_det.rows if full_det is not None else None),
        _ptr(full_det.count if full_det is not None else None), _ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count), cap,
        _stream())
    _lib.check(rc, "gather_slice_detections")
    return cands


def greedy_nmm(cands: "Candidates", det: "Detections", match_metric="IOS", match_threshold=0.5, class_agnostic=False,
               roi_mask: Optional[torch.Tensor] = None, roi_nc=0, roi_cnt: Optional[torch.Tensor] = None) -> "Detections":
    """SAHI's default merge of sliced predictions (``postprocess_type="GREEDYNMM"``, ``match_metric="IOS"``,
    ``match_threshold=0.5``, class-aware -- what ``get_sliced_prediction`` does when called as ``pipe.py:186-193``
    calls it): greedy non-maximum MERGING.  ``cands`` from ``gather_slice_detections``; the merged rows (union box,
    max score) go to ``det`` in descending score of the kept boxes, ``det.anchor`` = provenance of each kept box."""
    metric = {"IOU": 0, "IOS": 1}.get(str(match_metric).upper())
    if metric is None:
        raise ValueError("match_metric must be 'IOU' or 'IOS'")
    if not 0 <= match_threshold <= 1:
        raise ValueError("match_threshold must be in [0, 1]")
    F = cands.rows.shape[0]
    rc = _lib.load().b200yolo_greedy_nmm(_ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count), F, cands.cap, metric,
                                         float(match_threshold), int(bool(class_agnostic)), det.rows.shape[1],
                                         _ptr(det.rows), _ptr(det.anchor), _ptr(det.count), _ptr(roi_mask), int(roi_nc),
                                         _ptr(roi_cnt), _stream())
    _lib.check(rc, "greedy_nmm")
    return det


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
@dataclass
class Candidates:
    """Fixed-capacity candidate arrays produced by K2 (slot order arbitrary)."""
    rows: torch.Tensor      # (B, cap, 6) f32  x1,y1,x2,y2,score,class  (letterboxed px)
    anchor: torch.Tensor    # (B, cap) i32
    count: torch.Tensor     # (B,) i32 (may exceed cap on overflow)
    cap: int


def _alloc_candidates(B, cap, device, out: Optional[Candidates], zero=True):
    if out is not None:
        if out.cap != cap or out.rows.shape[0] != B:
            raise ValueError("candidate buffers do not match (B, cap)")
        if zero:
            out.count.zero_()
        return out
    return Candidates(torch.empty((B, cap, 6), dtype=torch.float32, device=device),
                      torch.empty((B, cap), dtype=torch.int32, device=device),
                      torch.zeros((B,), dtype=torch.int32, device=device), cap)


def stage_head_classes_h2d(head_host: torch.Tensor, out: torch.Tensor):
    """H2D copy of the class channels only (``head[:, 64:, :]``) of a pinned host head tensor into the
    same-shaped device tensor ``out``: one strided DMA, (nc / (64+nc)) of the bytes.  The DFL channels of
    ``out`` are left untouched -- pair with ``postprocess_small(head=head_host)`` (zero-copy DFL reads)."""
    if head_host.is_cuda or not head_host.is_pinned() or head_host.dtype != torch.float32 or head_host.dim() != 3 \
            or not head_host.is_contiguous():
        raise ValueError("head_host must be a pinned contiguous (B,64+nc,A) float32 tensor")
    if not out.is_cuda or out.shape != head_host.shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("out must be a contiguous CUDA float32 tensor of the same shape")
    B, no, A = head_host.shape
    off = 4 * REG_MAX * A * 4
    rc = _lib.load().b200yolo_copy2d_h2d(ctypes.c_void_p(out.data_ptr() + off), no * A * 4,
                                         ctypes.c_void_p(head_host.data_ptr() + off), no * A * 4,
                                         (no - 4 * REG_MAX) * A * 4, B, _stream())
    _lib.check(rc, "stage_head_classes_h2d")
    return out


def _require_dev_accessible(t: torch.Tensor, name: str, dtype):
    """CUDA tensor, or pinned host tensor (read zero-copy by the kernel)."""
    if isinstance(t, torch.Tensor) and not t.is_cuda and t.is_pinned() and t.dtype == dtype:
        return
    _require_cuda(t, name, dtype)


def _head_levels(head, strides, in_hw, level_hw, allow_pinned=False):
    """C-ABI level descriptors for a concatenated (B,64+nc,A) head or a list of per-level tensors.
    Returns (levels, n_levels, B, nc, A, device, keepalive).  ``allow_pinned``: pinned host tensors are accepted
    (zero-copy reads by the kernel) besides CUDA tensors."""
    check = _require_dev_accessible if allow_pinned else _require_cuda
    levels = (_lib.Level * 3)()
    if isinstance(head, (list, tuple)):
        if len(head) > 3 or len(head) != len(strides):
            raise ValueError("need one stride per level, at most 3 levels")
        B, no = head[0].shape[:2]
        keep = []
        for l, (x, s) in enumerate(zip(head, strides)):
            check(x, "head level", torch.float32)
            if x.dim() != 4 or x.shape[0] != B or x.shape[1] != no:
                raise ValueError("levels must be (B, 64+nc, Hi, Wi) with equal B and channels")
            if not x.is_cuda and not x.is_contiguous():
                raise ValueError("a pinned host head level must be contiguous")
            x = x if x.is_contiguous() else x.contiguous()
            keep.append(x)
            levels[l] = _lib.Level(x.data_ptr(), x.stride(0), x.stride(1), x.shape[2], x.shape[3], float(s))
        n_levels, A, device = len(head), sum(x.shape[2] * x.shape[3] for x in head), head[0].device
    else:
        check(head, "head", torch.float32)
        if head.dim() != 3:
            raise ValueError("head must be (B, 64+nc, A)")
        if not head.is_cuda and not head.is_contiguous():
            raise ValueError("a pinned host head must be contiguous")
        head = head if head.is_contiguous() else head.contiguous()
        keep = [head]
        B, no, A = head.shape
        if level_hw is None:
            if in_hw is None:
                raise ValueError("give in_hw=(h,w) of the letterboxed input or level_hw")
            level_hw = geometry.level_shapes(in_hw[0], in_hw[1], strides)
        if sum(h * w for h, w in level_hw) != A:
            raise ValueError(f"level shapes {level_hw} do not add up to A={A}")
        off = 0
        for l, ((h, w), s) in enumerate(zip(level_hw, strides)):
            levels[l] = _lib.Level(head.data_ptr() + 4 * off, head.stride(0), head.stride(1), h, w, float(s))
            off += h * w
        n_levels, device = len(level_hw), head.device
    nc = no - 4 * REG_MAX
    if nc <= 0:
        raise ValueError("head needs 64 DFL channels + nc class channels")
    return levels, n_levels, B, nc, A, device, keep


def decode_and_filter(head, strides=(8, 16, 32), conf_thres=0.25, classes=None, in_hw=None, level_hw=None,
                      cap=None, out: Optional[Candidates] = None, defer_boxes=False, zero=True) -> Candidates:
    """Detect-head decode + confidence filter.

    ``head`` is either the concatenated (B, 64+nc, A) fp32 tensor (``Detect._inference``'s ``x_cat``;
    give ``in_hw`` = letterboxed input (h, w) or ``level_hw``) or the list of per-level
    (B, 64+nc, Hi, Wi) tensors straight from the Detect convolutions.  ``defer_boxes=True`` runs the
    class filter only: the survivors' boxes are decoded later by ``postprocess_small``.  ``zero=False``: ``out.count``
    is already zero (``postprocess_small(..., cand_seen=...)`` re-armed it): no memset launch.
    """
    if not 0 <= conf_thres <= 1:
        raise ValueError(f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0")
    levels, n_levels, B, nc, A, device, keep = _head_levels(head, strides, in_hw, level_hw)
    cap = int(cap or A)
    cands = _alloc_candidates(B, cap, device, out, zero)
    mask = _class_mask(classes, nc, device)
    fn = _lib.load().b200yolo_class_filter if defer_boxes else _lib.load().b200yolo_decode_filter
    rc = fn(levels, n_levels, B, nc, float(conf_thres), _ptr(mask), _ptr(cands.rows), _ptr(cands.anchor),
            _ptr(cands.count), cap, _stream())
    _lib.check(rc, "decode_and_filter")
    del keep
    return cands


def postprocess_small(cands: Candidates, det: "Detections", head=None, strides=(8, 16, 32), in_hw=None, level_hw=None,
                      iou_thres=0.45, agnostic=False, max_nms=30000, max_wh=7680, scale: Optional[torch.Tensor] = None,
                      roi_mask: Optional[torch.Tensor] = None, roi_nc=0, roi_cnt: Optional[torch.Tensor] = None,
                      cand_seen: Optional[torch.Tensor] = None):
    """Fused box decode + sort + NMS (+rescale) for ``cands.cap <= 1024``: one launch, one CTA per image.

    ``head``: the same head passed to ``decode_and_filter(..., defer_boxes=True)`` (None if the candidate
    rows already hold boxes).  Results are bit-identical to ``nms_candidates``.  An image whose candidate count
    exceeds ``cands.cap`` is processed on its first ``cap`` slots only, which is NOT the reference's result: the
    caller must check the counts (``Pipeline`` / ``HostRunner`` raise ``CandidateOverflow``).
    ``cand_seen``: optional (B,) int32 -- receives the unclamped candidate counts while ``cands.count`` is reset to
    zero inside the kernel, so the next ``decode_and_filter(..., out=cands, zero=False)`` needs no memset launch."""
    if not 0 <= iou_thres <= 1:
        raise ValueError(f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0")
    B = cands.rows.shape[0]
    max_det = det.rows.shape[1]
    levels, n_levels, keep = None, 0, None
    if head is not None:
        # a pinned host head is allowed here: only the 64 DFL values of each survivor are read (zero-copy)
        levels, n_levels, _, _, _, _, keep = _head_levels(head, strides, in_hw, level_hw, allow_pinned=True)
    rc = _lib.load().b200yolo_postprocess_small(levels, n_levels, _ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count),
                                                B, cands.cap, int(max_nms), float(iou_thres), float(max_wh),
                                                int(bool(agnostic)), int(max_det), _ptr(scale), _ptr(det.rows),
                                                _ptr(det.anchor), _ptr(det.count), _ptr(roi_mask), int(roi_nc),
                                                _ptr(roi_cnt), _ptr(cand_seen), _stream())
    _lib.check(rc, "postprocess_small")
    del keep
    return det


def filter_decoded(prediction, conf_thres=0.25, classes=None, nc=0, cap=None,
                   out: Optional[Candidates] = None) -> Candidates:
    """Candidate filter on an already decoded UL-format prediction (B, 4+nc(+extra), A)."""
    _require_cuda(prediction, "prediction", torch.float32)
    if prediction.dim() != 3:
        raise ValueError("prediction must be (B, 4+nc, A)")
    prediction = prediction if prediction.is_contiguous() else prediction.contiguous()
    B, ch, A = prediction.shape
    nc = nc or (ch - 4)
    cap = int(cap or A)
    cands = _alloc_candidates(B, cap, prediction.device, out)
    mask = _class_mask(classes, nc, prediction.device)
    rc = _lib.load().b200yolo_filter_decoded(_ptr(prediction), B, ch, nc, A, float(conf_thres), _ptr(mask),
                                             _ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count), cap, _stream())
    _lib.check(rc, "filter_decoded")
    return cands


# ------------------------------------------------------------------------------------------------
# K3 + K4
# ------------------------------------------------------------------------------------------------
@dataclass
class Detections:
    """Padded NMS output: rows [x1,y1,x2,y2,conf,cls]; only the first count[b] rows of image b are valid."""
    rows: torch.Tensor      # (B, max_det, 6) f32
    anchor: torch.Tensor    # (B, max_det) i32 (original anchor index: UL return_idxs)
    count: torch.Tensor     # (B,) i32

    def to_list(self, return_idxs=False):
        """Ragged ``list[Tensor(n_i, 6)]`` as ``ops.non_max_suppression`` returns (one D2H of counts)."""
        counts = self.count.tolist()
        out = [self.rows[b, :n] for b, n in enumerate(counts)]
        if return_idxs:
            return out, [self.anchor[b, :n].to(torch.int64) for b, n in enumerate(counts)]
        return out


class Workspace:
    """Caller-owned scratch for sort/NMS (order array + oversize-image spill)."""

    def __init__(self, B, cap, max_det, device, det: Optional["Detections"] = None):
        """``det``: write the detections into these (B, max_det, ..) buffers (e.g. slices of a larger batch)."""
        self.B, self.cap, self.max_det = B, cap, max_det
        self.order = torch.empty((B, cap), dtype=torch.int32, device=device)
        nbytes = int(_lib.load().b200yolo_workspace_bytes(B, cap))
        self.ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=device)
        if det is not None and (tuple(det.rows.shape) != (B, max_det, 6) or not det.rows.is_contiguous()):
            raise ValueError("det buffers do not match (B, max_det)")
        self.det = det if det is not None else Detections(
            torch.empty((B, max_det, 6), dtype=torch.float32, device=device),
            torch.empty((B, max_det), dtype=torch.int32, device=device),
            torch.empty((B,), dtype=torch.int32, device=device))


def sort_candidates(cands: Candidates, max_nms=30000, ws: Optional[Workspace] = None) -> Workspace:
    B = cands.rows.shape[0]
    ws = ws or Workspace(B, cands.cap, 300, cands.rows.device)
    rc = _lib.load().b200yolo_sort_topk(_ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count), B, cands.cap,
                                        int(max_nms), _ptr(ws.order), _ptr(ws.ws), ws.ws.numel(), _stream())
    _lib.check(rc, "sort_candidates")
    return ws


def nms_sorted(cands: Candidates, ws: Workspace, iou_thres=0.45, agnostic=False, max_det=300, max_nms=30000,
               max_wh=7680, scale: Optional[torch.Tensor] = None, roi_mask: Optional[torch.Tensor] = None,
               roi_nc=0, roi_cnt: Optional[torch.Tensor] = None) -> Detections:
    """K4 on candidates whose order array (``ws.order``) was filled by ``sort_candidates``.
    ``roi_mask``/``roi_cnt``: optional class allow-list mask and (B,) i32 output of per-image ROI counts."""
    if not 0 <= iou_thres <= 1:
        raise ValueError(f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0")
    B = cands.rows.shape[0]
    if ws.max_det != max_det or ws.cap != cands.cap or ws.B != B:
        raise ValueError("workspace does not match (B, cap, max_det)")
    rc = _lib.load().b200yolo_nms(_ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count), _ptr(ws.order), B,
                                  cands.cap, int(max_nms), float(iou_thres), float(max_wh), int(bool(agnostic)),
                                  int(max_det), _ptr(scale), _ptr(ws.det.rows), _ptr(ws.det.anchor),
                                  _ptr(ws.det.count), _ptr(roi_mask), int(roi_nc), _ptr(roi_cnt), _ptr(ws.ws),
                                  ws.ws.numel(), _stream())
    _lib.check(rc, "nms")
    return ws.det


def postprocess_dense(cands: Candidates, ws: Workspace, head, strides=(8, 16, 32), in_hw=None, level_hw=None,
                      iou_thres=0.45, agnostic=False, max_det=300, max_nms=30000, max_wh=7680,
                      scale: Optional[torch.Tensor] = None, roi_mask: Optional[torch.Tensor] = None, roi_nc=0,
                      roi_cnt: Optional[torch.Tensor] = None) -> Detections:
    """Dense-regime chain on the survivors of ``decode_and_filter(..., defer_boxes=True)``: select-sort of the best
    2048 entries per image -> DFL decode of exactly those -> windowed NMS (+ exact full fallback for images that
    need more).  Bit-identical to ``decode_and_filter`` + ``sort_candidates`` + ``nms_sorted``."""
    if not 0 <= iou_thres <= 1:
        raise ValueError(f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0")
    B = cands.rows.shape[0]
    if ws.max_det != max_det or ws.cap != cands.cap or ws.B != B:
        raise ValueError("workspace does not match (B, cap, max_det)")
    levels, n_levels, _, _, _, _, keep = _head_levels(head, strides, in_hw, level_hw)
    rc = _lib.load().b200yolo_postprocess_dense(levels, n_levels, _ptr(cands.rows), _ptr(cands.anchor), _ptr(cands.count),
                                                B, cands.cap, int(max_nms), float(iou_thres), float(max_wh),
                                                int(bool(agnostic)), int(max_det), _ptr(scale), _ptr(ws.det.rows),
                                                _ptr(ws.det.anchor), _ptr(ws.det.count), _ptr(roi_mask), int(roi_nc),
                                                _ptr(roi_cnt), _ptr(ws.order), _ptr(ws.ws), ws.ws.numel(), _stream())
    _lib.check(rc, "postprocess_dense")
    del keep
    return ws.det


class DenseChain:
    """The whole low-confidence (evaluation-regime) chain -- class filter -> select-sort -> box decode -> windowed NMS
    (``decode_and_filter(defer_boxes=True)`` + ``postprocess_dense``) -- on ``splits`` sub-batches of the images, each on
    its own stream.  Images are independent, and the chain alternates HBM-bound kernels (class filter, box decode) with
    latency-bound per-image ones (select-sort, NMS): with two or more sub-batches in flight the latency-bound kernels
    of one run underneath the streaming kernels of another.  Results are those of the single-stream chain, bit for bit.
    Fork/join with events on the caller's current stream: capturable into a CUDA graph.

    ``pipelined=True``: the class filters of the sub-batches run back to back on the caller's stream and sub-batch s's
    select-sort -> decode -> NMS start on a high-priority stream as soon as ITS filter has finished.  Measured SLOWER
    than launching every sub-batch's whole chain at once (config 3, one CUDA graph: 338 / 353 us with 2 / 4 sub-batches
    against 323 us forked and 329 us single-stream at the time): the per-image kernels are latency-bound but not light -- a
    select-sort CTA holds 1024 of an SM's 2048 thread slots and 84 KB of its shared memory -- so they take whole SMs
    away from the streaming filter instead of hiding under it.  Kept as an option; the default is the forked form."""

    def __init__(self, B, cap, max_det, device, splits=8, pipelined=False):
        splits = max(1, min(int(splits), B))
        self.pipelined = bool(pipelined) and splits > 1
        self.B, self.cap, self.max_det, self.device = B, cap, max_det, torch.device(device)
        self.det = Detections(torch.empty((B, max_det, 6), dtype=torch.float32, device=device),
                              torch.empty((B, max_det), dtype=torch.int32, device=device),
                              torch.zeros((B,), dtype=torch.int32, device=device))
        self.cand_count = torch.zeros((B,), dtype=torch.int32, device=device)
        self.parts = []
        for s in range(splits):
            lo, hi = (s * B) // splits, ((s + 1) * B) // splits
            n = hi - lo
            cands = Candidates(torch.empty((n, cap, 6), dtype=torch.float32, device=device),
                               torch.empty((n, cap), dtype=torch.int32, device=device), self.cand_count[lo:hi], cap)
            ws = Workspace(n, cap, max_det, device, det=Detections(self.det.rows[lo:hi], self.det.anchor[lo:hi],
                                                                    self.det.count[lo:hi]))
            self.parts.append((lo, hi, cands, ws))
        prio = -1 if self.pipelined else 0
        self.streams = [torch.cuda.Stream(device=device, priority=prio) for _ in self.parts] if splits > 1 else [None]

    def __call__(self, head, strides=(8, 16, 32), conf_thres=0.001, iou_thres=0.7, classes=None, in_hw=None, level_hw=None,
                 agnostic=False, max_nms=30000, max_wh=7680, scale: Optional[torch.Tensor] = None,
                 roi_mask: Optional[torch.Tensor] = None, roi_nc=0, roi_cnt: Optional[torch.Tensor] = None) -> "Detections":
        cur = torch.cuda.current_stream()
        if self.pipelined:
            self.cand_count.zero_()
            for (lo, hi, cands, ws), st in zip(self.parts, self.streams):
                h = [x[lo:hi] for x in head] if isinstance(head, (list, tuple)) else head[lo:hi]
                decode_and_filter(h, strides, conf_thres, classes, in_hw=in_hw, level_hw=level_hw, cap=self.cap, out=cands,
                                  defer_boxes=True, zero=False)
                st.wait_stream(cur)                                       # sub-batch s is filtered: fork its tail
                with torch.cuda.stream(st):
                    postprocess_dense(cands, ws, h, strides, in_hw=in_hw, level_hw=level_hw, iou_thres=iou_thres,
                                      agnostic=agnostic, max_det=self.max_det, max_nms=max_nms, max_wh=max_wh,
                                      scale=None if scale is None else scale[lo:hi], roi_mask=roi_mask, roi_nc=roi_nc,
                                      roi_cnt=None if roi_cnt is None else roi_cnt[lo:hi])
            for st in self.streams:
                cur.wait_stream(st)                                       # join
            return self.det
        for (lo, hi, cands, ws), st in zip(self.parts, self.streams):
            h = [x[lo:hi] for x in head] if isinstance(head, (list, tuple)) else head[lo:hi]
            if st is not None:
                st.wait_stream(cur)                                       # fork
            with torch.cuda.stream(st if st is not None else cur):
                decode_and_filter(h, strides, conf_thres, classes, in_hw=in_hw, level_hw=level_hw, cap=self.cap, out=cands,
                                  defer_boxes=True)
                postprocess_dense(cands, ws, h, strides, in_hw=in_hw, level_hw=level_hw, iou_thres=iou_thres,
                                  agnostic=agnostic, max_det=self.max_det, max_nms=max_nms, max_wh=max_wh,
                                  scale=None if scale is None else scale[lo:hi], roi_mask=roi_mask, roi_nc=roi_nc,
                                  roi_cnt=None if roi_cnt is None else roi_cnt[lo:hi])
        for st in self.streams:
            if st is not None:
                cur.wait_stream(st)                                       # join
        return self.det


def nms_candidates(cands: Candidates, iou_thres=0.45, agnostic=False, max_det=300, max_nms=30000, max_wh=7680,
                   scale: Optional[torch.Tensor] = None, ws: Optional[Workspace] = None) -> Detections:
    """K3 + K4 on K2's candidates.  ``scale``: optional (B,5) f32 {gain,pad_x,pad_y,w0,h0} -> source pixels."""
    B = cands.rows.shape[0]
    if ws is None or ws.max_det != max_det or ws.cap != cands.cap or ws.B != B:
        ws = Workspace(B, cands.cap, max_det, cands.rows.device)
    sort_candidates(cands, max_nms, ws)
    return nms_sorted(cands, ws, iou_thres, agnostic, max_det, max_nms, max_wh, scale)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000,
                        max_wh=7680, in_place=True, rotated=False, end2end=False, return_idxs=False):
    """Drop-in for ``ultralytics.utils.ops.non_max_suppression`` (detect task) on a CUDA prediction.

    Same signature and return value (``list[Tensor(n_i, 6)]``, or ``(output, keepi)`` with
    ``return_idxs``).  Differences, all documented in DESIGN.md: the input is never modified
    (``in_place`` is accepted and ignored), the upstream wall-clock ``max_time_img`` break is not
    replicated, and ``multi_label`` / ``labels`` / ``rotated`` / ``end2end`` / extra mask channels
    raise ``NotImplementedError`` instead of silently falling back.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if rotated or end2end or labels:
        raise NotImplementedError("rotated / end2end / labels are outside the reference's call pattern")
    _require_cuda(prediction, "prediction")
    nc = nc or (prediction.shape[1] - 4)
    if multi_label and nc > 1:
        raise NotImplementedError("multi_label=True is not implemented (the reference never sets it)")
    if prediction.shape[1] - nc - 4 != 0:
        raise NotImplementedError("extra mask channels (segment/pose) are outside this path")
    cands = filter_decoded(prediction.float(), conf_thres, classes, nc)
    det = nms_candidates(cands, iou_thres, agnostic, max_det, max_nms, max_wh)
    out = det.to_list(return_idxs)
    if return_idxs:
        return [o.clone() for o in out[0]], out[1]
    return [o.clone() for o in out]


# ------------------------------------------------------------------------------------------------
# a10
# ------------------------------------------------------------------------------------------------
def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """``ops.scale_boxes`` (+ ``clip_boxes``) drop-in on a CUDA (n, >=4) xyxy tensor, in place."""
    if xywh or not padding:
        raise NotImplementedError("xywh=True / padding=False are outside the reference's call pattern")
    _require_cuda(boxes, "boxes", torch.float32)
    if boxes.dim() != 2 or boxes.shape[1] < 4 or boxes.stride(1) != 1:
        raise ValueError("boxes must be (n, >=4) float32 with unit inner stride")
    gain, pad = geometry.scale_boxes_params(img1_shape, img0_shape, ratio_pad)
    if boxes.shape[0] == 0:              # a frame without detections (upstream construct_result calls this on (0,4) views)
        return boxes
    rc = _lib.load().b200yolo_scale_boxes(_ptr(boxes), boxes.shape[0], boxes.stride(0) if boxes.shape[0] else 4,
                                          float(gain), float(pad[0]), float(pad[1]), float(img0_shape[1]),
                                          float(img0_shape[0]), _stream())
    _lib.check(rc, "scale_boxes")
    return boxes


def scale_params_tensor(img1_shape, img0_shapes: Sequence, device):
    """(B,5) {gain, pad_x, pad_y, w0, h0} rows for the fused rescale inside the NMS kernel."""
    rows = []
    for s0 in img0_shapes:
        gain, pad = geometry.scale_boxes_params(img1_shape, s0)
        rows.append([gain, pad[0], pad[1], s0[1], s0[0]])
    return torch.tensor(rows, dtype=torch.float32, device=device)


# ------------------------------------------------------------------------------------------------
# K5
# ------------------------------------------------------------------------------------------------
def crop_resize_rois(frames, boxes_xyxy, batch_idx, pad=6, size=64, roi_count=None, out=None, valid=None):
    """Batched ``safe_crop`` + classifier preprocessing: returns ((N,3,size,size) fp32 RGB, valid (N,) i32).

    ``boxes_xyxy``: (N,4) fp32 source-pixel boxes (``int()`` truncation happens on the device, as
    ``detect.py:581`` does on the host); ``batch_idx``: (N,) int32 frame index.  ``valid`` is 1 (or 2: produced by the split
    large-ROI launch) for a valid crop, 0 where ``safe_crop`` would return None and -1 where the ROI is larger
    than the kernel's envelope (short side > 31 x size).
    """
    frames, _ = _frames_4d(frames, allow_pinned=True)
    _require_cuda(boxes_xyxy, "boxes_xyxy", torch.float32)
    _require_cuda(batch_idx, "batch_idx", torch.int32)
    boxes_xyxy = boxes_xyxy.contiguous()
    batch_idx = batch_idx.contiguous()
    N = boxes_xyxy.shape[0]
    if boxes_xyxy.dim() != 2 or boxes_xyxy.shape[1] != 4 or batch_idx.shape != (N,):
        raise ValueError("boxes_xyxy must be (N,4) and batch_idx (N,)")
    B, H, W, _ = frames.shape
    if out is None:
        out = torch.empty((N, 3, size, size), dtype=torch.float32, device=boxes_xyxy.device)
    if valid is None:
        valid = torch.empty((N,), dtype=torch.int32, device=boxes_xyxy.device)
    rc = _lib.load().b200yolo_roi_crop_resize(_ptr(frames), B, H, W, frames.stride(1), frames.stride(0),
                                              _ptr(boxes_xyxy), _ptr(batch_idx), _ptr(roi_count), N, int(pad),
                                              int(size), _ptr(out), _ptr(valid), _stream())
    _lib.check(rc, "crop_resize_rois")
    return out, valid


def classify_preprocess(crops, size=64, device="cuda"):
    """``ClassificationPredictor.preprocess`` drop-in for a list of host crops (``rank_model(crop)``,
    ``/root/reference/detect.py:121``: BGR HWC uint8 numpy arrays of any sizes): the crops are stacked into one pinned
    canvas, uploaded once, and resized by K5 in one launch (``pad=0``: every box is exactly its crop) -- BGR->RGB,
    Pillow-exact ``Resize(64)`` + ``CenterCrop(64)`` + ``ToTensor``.  Returns (N,3,size,size) fp32 on ``device``."""
    import numpy as np
    if not crops:
        return torch.empty((0, 3, size, size), dtype=torch.float32, device=device)
    for c in crops:
        if not isinstance(c, np.ndarray) or c.ndim != 3 or c.shape[2] != 3 or c.dtype != np.uint8 or c.size == 0:
            raise ValueError("crops must be non-empty HxWx3 uint8 numpy arrays (BGR)")
    wmax = max(c.shape[1] for c in crops)
    W = ((wmax + 8) * 3 + 15) // 16 * 16 // 3 + 16           # a few spare columns: no crop touches the buffer's edge
    H = sum(c.shape[0] for c in crops) + 2                    # one spare row above and below
    canvas = torch.zeros((1, H, W, 3), dtype=torch.uint8).pin_memory()
    cv = canvas.numpy()
    boxes, y = [], 1
    for c in crops:
        h, w = c.shape[:2]
        cv[0, y:y + h, 4:4 + w] = c
        boxes.append([4.0, float(y), float(4 + w), float(y + h)])
        y += h
    d = canvas.to(device, non_blocking=True)
    bx = torch.tensor(boxes, dtype=torch.float32).to(device, non_blocking=True)
    bidx = torch.zeros((len(crops),), dtype=torch.int32, device=device)
    out, valid = crop_resize_rois(d, bx, bidx, pad=0, size=size)
    if int((valid <= 0).sum()):
        raise ValueError("a crop is outside the K5 envelope (short side > 31 x 64 px)")
    return out


def rois_from_detections(frames, det: Detections, roi_cnt, roi_mask, nc, roi_cap, pad=6, size=64, out=None):
    """Pipeline form of K5: crops + resizes every detection whose class is in ``roi_mask`` straight from
    the NMS output (``roi_cnt`` = per-image counts written by ``nms_sorted``).  Returns
    (rois (roi_cap,3,size,size), roi_batch, roi_det, valid, roi_total (1,)); ``roi_total`` is the number of detections
    of the ROI classes in the batch, NOT clamped to ``roi_cap`` (only the first ``roi_cap`` are cropped)."""
    frames, _ = _frames_4d(frames, allow_pinned=True)     # pinned host frames: crops read zero-copy over PCIe
    B, H, W, _ = frames.shape
    max_det = det.rows.shape[1]
    dev = det.rows.device
    if out is None:
        out = (torch.empty((roi_cap, 3, size, size), dtype=torch.float32, device=dev),
               torch.empty((roi_cap,), dtype=torch.int32, device=dev),
               torch.empty((roi_cap,), dtype=torch.int32, device=dev),
               torch.empty((roi_cap,), dtype=torch.int32, device=dev),
               torch.empty((1,), dtype=torch.int32, device=dev))
    rc = _lib.load().b200yolo_roi_from_detections(_ptr(frames), B, H, W, frames.stride(1), frames.stride(0),
                                                  _ptr(det.rows), _ptr(det.count), _ptr(roi_cnt), max_det,
                                                  _ptr(roi_mask), int(nc), int(pad), int(size), _ptr(out[0]),
                                                  _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _ptr(out[4]),
                                                  int(roi_cap), _stream())
    _lib.check(rc, "rois_from_detections")
    return out


def select_rois(det: Detections, classes, nc, roi_cap, out=None):
    """Device-side gather of the detections whose class is in ``classes`` (the ``*_rank`` ids).

    Returns (roi_boxes (cap,4) f32, roi_batch (cap,) i32, roi_det (cap,) i32, roi_count (1,) i32); ``roi_count`` is not
    clamped to ``roi_cap`` (rows beyond it are dropped)."""
    B, max_det, _ = det.rows.shape
    dev = det.rows.device
    mask = _class_mask(classes, nc, dev)
    if out is None:
        out = (torch.empty((roi_cap, 4), dtype=torch.float32, device=dev),
               torch.empty((roi_cap,), dtype=torch.int32, device=dev),
               torch.empty((roi_cap,), dtype=torch.int32, device=dev),
               torch.zeros((1,), dtype=torch.int32, device=dev))
    else:
        out[3].zero_()
    rc = _lib.load().b200yolo_select_rois(_ptr(det.rows), _ptr(det.count), B, max_det, _ptr(mask), int(nc),
                                          _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), int(roi_cap),
                                          _stream())
    _lib.check(rc, "select_rois")
    return out


# ------------------------------------------------------------------------------------------------
# N2: tracker association costs
# ------------------------------------------------------------------------------------------------
def iou_cost_matrix(tracks_xyxy: torch.Tensor, track_count: torch.Tensor, det: Detections, fuse_score=False,
                    pad_cost=1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ByteTrack association costs for a batch of frames / streams (``detect.py:557`` ->
    supervision ``iou_distance`` [+ ``fuse_score``]): ``tracks_xyxy`` (B,T,4) fp32 predicted track boxes,
    ``track_count`` (B,) int32, ``det`` the padded NMS output.  Returns (B,T,max_det) fp32 costs ``1 - iou``
    (``1 - iou * score`` when fused); padding entries hold ``pad_cost``.  The linear assignment stays on the host
    (``handoff.associate``)."""
    _require_cuda(tracks_xyxy, "tracks_xyxy", torch.float32)
    _require_cuda(track_count, "track_count", torch.int32)
    tracks_xyxy = tracks_xyxy.contiguous()
    if tracks_xyxy.dim() != 3 or tracks_xyxy.shape[2] != 4 or tracks_xyxy.shape[0] != det.rows.shape[0]:
        raise ValueError("tracks_xyxy must be (B,T,4) with the same B as det")
    B, T, _ = tracks_xyxy.shape
    max_det = det.rows.shape[1]
    if out is None:
        out = torch.empty((B, T, max_det), dtype=torch.float32, device=tracks_xyxy.device)
    rc = _lib.load().b200yolo_iou_cost_matrix(_ptr(tracks_xyxy), _ptr(track_count), _ptr(det.rows), _ptr(det.count), B, T,
                                              max_det, int(bool(fuse_score)), float(pad_cost), _ptr(out), _stream())
    _lib.check(rc, "iou_cost_matrix")
    return out
