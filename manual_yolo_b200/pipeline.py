"""Whole hot path for a batch of frames: letterbox -> decode/filter -> sort -> NMS(+rescale) -> ROI crops.

This is the batched, device-resident form of the reference's per-frame loop
(``/root/reference/detect.py:535-600``): ``model(frame)`` (``detect.py:541``), the per-detection
``int()`` + ``safe_crop(pad=6)`` (``detect.py:581-586``) and the per-crop ``rank_model(crop)``
preprocessing (``detect.py:121``).  The backbone/neck stay torch modules outside this package: the
Detect-head tensor is an input here (``head``), the letterboxed network input an output (``net_in``).

All buffers are allocated once in ``__init__`` (the C ABI never allocates); a step is 6 launches of
this package's kernels plus one counter memset, capturable into one CUDA graph.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import api, geometry

RANK_CLASS_IDS = (6, 11, 16, 21, 26, 37, 43)  # *_rank classes, roadmap1.v3i.yolov8/data.yaml:6
FUSED_CAP_MAX = 1024                           # b200yolo_postprocess_small envelope


@dataclass
class PipelineResult:
    net_in: torch.Tensor        # (B,3,h,w) f32 letterboxed + normalised network input (K1)
    det: api.Detections         # boxes in SOURCE pixels (scale_boxes fused into the NMS epilogue)
    cand_count: torch.Tensor    # (B,) i32 candidates that passed conf (K2)
    rois: torch.Tensor          # (roi_cap,3,S,S) f32 classifier batch (K5); first roi_count rows valid
    roi_batch: torch.Tensor     # (roi_cap,) i32 frame index of each ROI
    roi_det: torch.Tensor       # (roi_cap,) i32 detection row of each ROI
    roi_valid: torch.Tensor     # (roi_cap,) i32
    roi_count: torch.Tensor     # (1,) i32


class Pipeline:
    def __init__(self, batch: int, src_hw, nc: int, imgsz=640, auto=False, conf=0.25, iou=0.7, max_det=300,
                 agnostic=False, classes=None, max_nms=30000, max_wh=7680, roi_classes: Sequence[int] = RANK_CLASS_IDS,
                 rois_per_frame=8, pad=6, roi_size=64, strides=(8, 16, 32), device="cuda", cap=None, overlap=False):
        if not torch.cuda.is_available():
            raise RuntimeError("manual_yolo_b200.Pipeline needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.B, self.src_hw, self.nc = int(batch), (int(src_hw[0]), int(src_hw[1])), int(nc)
        self.conf, self.iou, self.max_det, self.agnostic = conf, iou, max_det, agnostic
        self.classes, self.max_nms, self.max_wh = classes, max_nms, max_wh
        self.roi_classes, self.pad, self.roi_size, self.strides = tuple(roi_classes), pad, roi_size, tuple(strides)
        new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        self.new_shape, self.auto = new_shape, auto
        g = geometry.letterbox_geometry(self.src_hw, new_shape, auto=auto, stride=max(int(s) for s in strides))
        self.geom = g
        self.in_hw = (g["out_h"], g["out_w"])
        self.level_hw = geometry.level_shapes(g["out_h"], g["out_w"], strides)
        self.A = sum(h * w for h, w in self.level_hw)
        self.cap = int(cap or self.A)
        self.fused = self.cap <= FUSED_CAP_MAX       # sparse regime: fused per-image post-processing kernel
        dev = self.device
        B = self.B
        self.net_in = torch.empty((B, 3, g["out_h"], g["out_w"]), dtype=torch.float32, device=dev)
        self.cands = api.Candidates(torch.empty((B, self.cap, 6), dtype=torch.float32, device=dev),
                                    torch.empty((B, self.cap), dtype=torch.int32, device=dev),
                                    torch.zeros((B,), dtype=torch.int32, device=dev), self.cap)
        self.ws = api.Workspace(B, self.cap, max_det, dev)
        self.scale = api.scale_params_tensor(self.in_hw, [self.src_hw] * B, dev)
        self.roi_cap = max(1, B * int(rois_per_frame))
        self.cls_mask = api._class_mask(classes, self.nc, dev)
        self.roi_mask = api._class_mask(self.roi_classes, self.nc, dev)
        self.roi_cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.roi_out = (torch.zeros((self.roi_cap, 3, roi_size, roi_size), dtype=torch.float32, device=dev),
                        torch.zeros((self.roi_cap,), dtype=torch.int32, device=dev),
                        torch.zeros((self.roi_cap,), dtype=torch.int32, device=dev),
                        torch.zeros((self.roi_cap,), dtype=torch.int32, device=dev),
                        torch.zeros((1,), dtype=torch.int32, device=dev))
        self._graph = None
        self._static = None
        self._prof = None
        self._open = None
        # overlap=True: the letterbox kernel (HBM-bound, feeds the backbone) runs on a side stream,
        # concurrently with decode -> NMS -> ROI of the same batch (latency-bound, feed the classifier);
        # the two branches share no buffer.  Fork/join with events, capturable into one CUDA graph.
        self.overlap = bool(overlap)
        self._side = self._hi = None
        if self.overlap:
            self._make_streams()

    def _make_streams(self):
        # K1 on a normal-priority stream, the latency-bound branch on a HIGH-priority one: its small CTAs are
        # scheduled first whenever an SM frees resources, so they run in the shadow of the streaming kernel
        self._side = torch.cuda.Stream(device=self.device, priority=0)
        self._hi = torch.cuda.Stream(device=self.device, priority=-1)

    # -- one step on device-resident inputs ------------------------------------------------------
    def __call__(self, frames: torch.Tensor, head) -> PipelineResult:
        """frames: (B,H,W,3) uint8 BGR on the device; head: (B,64+nc,A) fp32 or list of level tensors."""
        if tuple(frames.shape) != (self.B, self.src_hw[0], self.src_hw[1], 3):
            raise ValueError(f"frames must be {(self.B, *self.src_hw, 3)}, got {tuple(frames.shape)}")
        t = self._tick
        if self.overlap:
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)                               # fork
            self._hi.wait_stream(main)
            with torch.cuda.stream(self._side):
                api.preprocess(frames, self.new_shape, auto=self.auto, stride=max(int(s) for s in self.strides),
                               out=self.net_in)
            with torch.cuda.stream(self._hi):
                res = self._post_branch(frames, head, t)
            main.wait_stream(self._side)                               # join
            main.wait_stream(self._hi)
            return res
        else:
            t("letterbox")
            api.preprocess(frames, self.new_shape, auto=self.auto, stride=max(int(s) for s in self.strides),
                           out=self.net_in)
        return self._post_branch(frames, head, t)

    def _post_branch(self, frames, head, t):
        """decode -> (sort ->) NMS -> ROI crops on the current stream."""
        if self.fused:
            # sparse regime (cap <= 1024): class filter, then ONE fused launch for decode + sort + NMS
            t("decode_filter")
            api.decode_and_filter(head, self.strides, self.conf, self.cls_mask, level_hw=self.level_hw,
                                  cap=self.cap, out=self.cands, defer_boxes=True)
            t("postprocess_small")
            det = api.postprocess_small(self.cands, self.ws.det, head, self.strides, level_hw=self.level_hw,
                                        iou_thres=self.iou, agnostic=self.agnostic, max_nms=self.max_nms,
                                        max_wh=self.max_wh, scale=self.scale, roi_mask=self.roi_mask, roi_nc=self.nc,
                                        roi_cnt=self.roi_cnt)
        else:
            t("decode_filter")
            api.decode_and_filter(head, self.strides, self.conf, self.cls_mask, level_hw=self.level_hw,
                                  cap=self.cap, out=self.cands)
            t("sort_topk")
            api.sort_candidates(self.cands, self.max_nms, self.ws)
            t("nms")
            det = api.nms_sorted(self.cands, self.ws, self.iou, self.agnostic, self.max_det, self.max_nms,
                                 self.max_wh, scale=self.scale, roi_mask=self.roi_mask, roi_nc=self.nc,
                                 roi_cnt=self.roi_cnt)
        t("roi_crop_resize")
        ro = api.rois_from_detections(frames, det, self.roi_cnt, self.roi_mask, self.nc, self.roi_cap, self.pad,
                                      self.roi_size, out=self.roi_out)
        t(None)
        return PipelineResult(self.net_in, det, self.cands.count, ro[0], ro[1], ro[2], ro[3], ro[4])

    # -- per-kernel CUDA-event timing (bench.py): events on the launching stream around each stage --
    def enable_profiling(self, on=True, external=False):
        """``external=True`` records the events as external nodes so they can sit inside a captured CUDA
        graph: after ``capture()`` every ``replay()`` + synchronize refreshes ``kernel_times_ms()``."""
        self._prof = [] if on else None
        self._open = None
        self._prof_external = bool(external)

    def _tick(self, name):
        if getattr(self, "_prof", None) is None:
            return
        ev = torch.cuda.Event(enable_timing=True, external=getattr(self, "_prof_external", False))
        ev.record()
        if self._open is not None:
            self._prof.append((self._open[0], self._open[1], ev))
        self._open = (name, ev) if name is not None else None

    def kernel_times_ms(self):
        """{stage: [ms per recorded launch]} -- call after torch.cuda.synchronize()."""
        out = {}
        for name, a, b in self._prof or []:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out

    # -- CUDA-graph form: one launch per batch ---------------------------------------------------
    def capture(self, frames: torch.Tensor, head: torch.Tensor):
        """Capture one step on static input buffers; afterwards ``replay()`` re-runs it (copy new data
        into the tensors passed here first)."""
        self._static = (frames, head)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self(frames, head)                       # warm-up: sets kernel attributes, touches buffers
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self._prof is not None:
            self._prof, self._open = [], None        # keep only the events recorded inside the graph
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._result = self(frames, head)
        return self._result

    def replay(self) -> PipelineResult:
        if self._graph is None:
            raise RuntimeError("capture() first")
        self._graph.replay()
        return self._result

    # -- host-facing entry: pinned host buffers in, host results out -----------------------------
    def run_host(self, frames_host: torch.Tensor, head_host: torch.Tensor, staging=None):
        """End-to-end call on HOST (pinned) buffers: H2D of the frames and the head tensor, the device
        path, D2H of detections/counts.  Returns (det_rows, det_count, roi_count) host tensors."""
        if staging is None:
            staging = self.make_staging()
        d_frames, d_head, h_rows, h_count, h_roi = staging
        d_frames.copy_(frames_host, non_blocking=True)
        d_head.copy_(head_host, non_blocking=True)
        res = self(d_frames, d_head)
        h_rows.copy_(res.det.rows, non_blocking=True)
        h_count.copy_(res.det.count, non_blocking=True)
        h_roi.copy_(res.roi_count, non_blocking=True)
        return h_rows, h_count, h_roi

    def launches_per_step(self):
        """Kernels of this package launched per step: letterbox, class filter, [fused post-processing | box
        decode, sort, nms], ROI crops, large-ROI pass."""
        return 5 if self.fused else 7

    def check_overflow(self):
        """Raise if any image of the last step had more candidates than ``cap`` (one D2H of the counts).
        With ``cap < A`` a candidate beyond ``cap`` is dropped, so results are only valid when this passes."""
        mx = int(self.cands.count.max())
        if mx > self.cap:
            raise RuntimeError(f"candidate overflow: {mx} candidates in one image > cap={self.cap}; "
                               "raise cap (cap=None uses all anchors) or the confidence threshold")
        return mx

    def make_staging(self):
        dev = self.device
        no = 64 + self.nc
        return (torch.empty((self.B, self.src_hw[0], self.src_hw[1], 3), dtype=torch.uint8, device=dev),
                torch.empty((self.B, no, self.A), dtype=torch.float32, device=dev),
                torch.empty((self.B, self.max_det, 6), dtype=torch.float32).pin_memory(),
                torch.empty((self.B,), dtype=torch.int32).pin_memory(),
                torch.empty((1,), dtype=torch.int32).pin_memory())

    def h2d_bytes_per_step(self):
        return self.B * self.src_hw[0] * self.src_hw[1] * 3 + self.B * (64 + self.nc) * self.A * 4

    def d2h_bytes_per_step(self):
        return self.B * self.max_det * 6 * 4 + self.B * 4 + 4


class HostRunner:
    """End-to-end driver on HOST buffers: every step copies that step's frames + head tensor from pinned
    host memory (copy stream), runs the device path (compute stream) and reads the detections back.
    Two staging sets let step i+1's H2D overlap step i's kernels."""

    def __init__(self, pipe: Pipeline, depth=2, head_resident=None):
        """``head_resident``: a device head tensor to use every step instead of copying one from the host
        (the deployment case: the Detect head is produced on the device by the backbone)."""
        self.pipe, self.depth, self.head_resident = pipe, depth, head_resident
        self.staging = [pipe.make_staging() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=pipe.device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.step = 0

    def submit(self, frames_host: torch.Tensor, head_host: torch.Tensor):
        """Enqueue one step; returns the (rows, count, roi_count) pinned host tensors it will fill."""
        s = self.step % self.depth
        d_frames, d_head, h_rows, h_count, h_roi = self.staging[s]
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if self.step >= self.depth:
                self.copy_stream.wait_event(self.free[s])
            d_frames.copy_(frames_host, non_blocking=True)
            if self.head_resident is None:
                d_head.copy_(head_host, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        compute.wait_event(self.ready[s])
        res = self.pipe(d_frames, d_head if self.head_resident is None else self.head_resident)
        h_rows.copy_(res.det.rows, non_blocking=True)
        h_count.copy_(res.det.count, non_blocking=True)
        h_roi.copy_(res.roi_count, non_blocking=True)
        self.free[s].record(compute)
        self.step += 1
        return h_rows, h_count, h_roi


def detections_to_records(det_rows, det_count, names=None, frame_offset=0):
    """Host-side gather into the reference's ``frame_data`` schema (``detect.py:590-598``), without the
    OCR/tracker fields the path does not produce: bbox ints are ``int()``-truncated as ``detect.py:581``."""
    recs = []
    rows = det_rows.tolist() if hasattr(det_rows, "tolist") else det_rows
    counts = det_count.tolist() if hasattr(det_count, "tolist") else det_count
    for b, n in enumerate(counts):
        for i in range(n):
            x1, y1, x2, y2, conf, cls = rows[b][i]
            cid = int(cls)
            recs.append({"frame": frame_offset + b, "tracker_id": -1, "class_id": cid,
                         "class_name": names.get(cid, f"class{cid}") if names else f"class{cid}",
                         "bbox": [int(x1), int(y1), int(x2), int(y2)], "conf": round(float(conf), 3)})
    return recs
