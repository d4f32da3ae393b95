"""Whole hot path for a batch of frames: letterbox -> decode/filter -> sort -> NMS(+rescale) -> ROI crops.

This is the batched, device-resident form of the reference's per-frame loop
(``/root/reference/detect.py:535-600``): ``model(frame)`` (``detect.py:541``), the per-detection
``int()`` + ``safe_crop(pad=6)`` (``detect.py:581-586``) and the per-crop ``rank_model(crop)``
preprocessing (``detect.py:121``).  The backbone/neck stay torch modules outside this package: the
Detect-head tensor is an input here (``head``), the letterboxed network input an output (``net_in``).

All buffers are allocated once in ``__init__`` (the C ABI never allocates); a step is 5 launches of this
package's kernels in the sparse regime and nothing else (the fused post-processing kernel re-arms the compaction
counters itself: no memset node), 7-10 launches plus one counter memset on the general / dense path; capturable
into one CUDA graph.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import api, geometry

RANK_CLASS_IDS = (6, 11, 16, 21, 26, 37, 43)  # *_rank classes, roadmap1.v3i.yolov8/data.yaml:6
FUSED_CAP_MAX = 1024                           # b200yolo_postprocess_small envelope


class CapacityOverflow(RuntimeError):
    """A fixed-capacity buffer of the step was too small for its input: the results of the step are NOT the
    reference's and must not be used (raise rather than truncate silently)."""


class CandidateOverflow(CapacityOverflow):
    """More candidates passed the confidence threshold in one image than ``cap`` slots exist."""


class RoiOverflow(CapacityOverflow):
    """More detections of the ROI classes in the batch than ``roi_cap`` crops exist."""


def check_counts(cand_count, cap, roi_total=None, roi_cap=None):
    """Host-side capacity check on counts that were already copied back (no extra device sync): raises
    ``CandidateOverflow`` / ``RoiOverflow``.  Returns the largest per-image candidate count."""
    mx = int(cand_count.max()) if cand_count.numel() else 0
    if mx > cap:
        raise CandidateOverflow(f"candidate overflow: {mx} candidates in one image > cap={cap}; the step's detections are "
                                "invalid -- raise cap (cap=None uses all anchors) or the confidence threshold")
    if roi_total is not None and roi_cap is not None and int(roi_total) > roi_cap:
        raise RoiOverflow(f"ROI overflow: {int(roi_total)} detections of the ROI classes in the batch > roi_cap={roi_cap}; "
                          "raise rois_per_frame")
    return mx


@dataclass
class PipelineResult:
    net_in: torch.Tensor        # (B,3,h,w) f32 letterboxed + normalised network input (K1)
    det: api.Detections         # boxes in SOURCE pixels (scale_boxes fused into the NMS epilogue)
    cand_count: torch.Tensor    # (B,) i32 candidates that passed conf (K2)
    rois: torch.Tensor          # (roi_cap,3,S,S) f32 classifier batch (K5); first roi_count rows valid
    roi_batch: torch.Tensor     # (roi_cap,) i32 frame index of each ROI
    roi_det: torch.Tensor       # (roi_cap,) i32 detection row of each ROI
    roi_valid: torch.Tensor     # (roi_cap,) i32
    roi_count: torch.Tensor     # (1,) i32: ROI-class detections in the batch (NOT clamped: > roi_cap = overflow)
    cap: int = 0                # candidate slots per image (cand_count[b] > cap = overflow: results invalid)

    def n_rois(self) -> int:
        """Valid rows of ``rois`` (one device->host read)."""
        return min(int(self.roi_count), int(self.rois.shape[0]))

    def check_overflow(self) -> int:
        """Raise ``CandidateOverflow`` / ``RoiOverflow`` if a capacity was exceeded in this step (device->host read of
        the counts).  Returns the largest per-image candidate count."""
        return check_counts(self.cand_count.cpu(), self.cap if self.cap > 0 else 1 << 30, self.roi_count.cpu(),
                            int(self.rois.shape[0]))


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
        self.rows_plan = geometry.referenced_rows(self.src_hw[0], g["new_h"])   # (row0, step, n_rows) K1 reads
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
        # fused path: the post-processing kernel hands the per-image candidate counts out here and re-arms
        # cands.count itself, so a step has no memset launch
        self.cand_seen = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.nvtx = False                            # NVTX range per stage (nsys / ncu --nvtx), see _tick
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
    def __call__(self, frames: torch.Tensor, head, roi_frames: Optional[torch.Tensor] = None,
                 dfl_head=None) -> PipelineResult:
        """frames: (B,H,W,3) uint8 BGR on the device; head: (B,64+nc,A) fp32 or list of level tensors.

        Host-fed form (``HostRunner``): ``frames`` may instead be the (B,n_rows,W,3) rows staged by
        ``api.stage_rows_h2d`` (only what the letterbox reads); the ROI crops are then taken from
        ``roi_frames`` -- the full frames, on the device or in PINNED host memory (zero-copy).  ``dfl_head``:
        tensor the survivors' DFL channels are read from (default ``head``; may be the pinned host head when
        only the class channels of ``head`` were staged)."""
        full = (self.B, self.src_hw[0], self.src_hw[1], 3)
        self._lb_src_hw = None
        if tuple(frames.shape) != full:
            if roi_frames is None or tuple(frames.shape) != (self.B, self.rows_plan[2], self.src_hw[1], 3):
                raise ValueError(f"frames must be {full} (or the staged rows {(self.B, self.rows_plan[2], self.src_hw[1], 3)} "
                                 f"together with roi_frames), got {tuple(frames.shape)}")
            self._lb_src_hw = self.src_hw
        if roi_frames is not None and tuple(roi_frames.shape) != full:
            raise ValueError(f"roi_frames must be {full}, got {tuple(roi_frames.shape)}")
        self._roi_frames = frames if roi_frames is None else roi_frames
        self._dfl_head = head if dfl_head is None else dfl_head
        if dfl_head is not None and not self.fused:
            raise ValueError("dfl_head needs the fused sparse-regime path (cap <= 1024)")
        t = self._tick
        if self.overlap:
            main = torch.cuda.current_stream()
            self._side.wait_stream(main)                               # fork
            self._hi.wait_stream(main)
            with torch.cuda.stream(self._side):
                api.preprocess(frames, self.new_shape, auto=self.auto, stride=max(int(s) for s in self.strides),
                               out=self.net_in, src_hw=self._lb_src_hw)
            with torch.cuda.stream(self._hi):
                res = self._post_branch(frames, head, t)
            main.wait_stream(self._side)                               # join
            main.wait_stream(self._hi)
            return res
        else:
            t("letterbox")
            api.preprocess(frames, self.new_shape, auto=self.auto, stride=max(int(s) for s in self.strides),
                           out=self.net_in, src_hw=self._lb_src_hw)
        return self._post_branch(frames, head, t)

    def _post_branch(self, frames, head, t):
        """decode -> (sort ->) NMS -> ROI crops on the current stream."""
        if self.fused:
            # sparse regime (cap <= 1024): class filter, then ONE fused launch for decode + sort + NMS
            t("decode_filter")
            api.decode_and_filter(head, self.strides, self.conf, self.cls_mask, level_hw=self.level_hw,
                                  cap=self.cap, out=self.cands, defer_boxes=True, zero=False)
            t("postprocess_small")
            det = api.postprocess_small(self.cands, self.ws.det, self._dfl_head, self.strides, level_hw=self.level_hw,
                                        iou_thres=self.iou, agnostic=self.agnostic, max_nms=self.max_nms,
                                        max_wh=self.max_wh, scale=self.scale, roi_mask=self.roi_mask, roi_nc=self.nc,
                                        roi_cnt=self.roi_cnt, cand_seen=self.cand_seen)
            cand_count = self.cand_seen
        else:
            # general / dense regime: class filter, then ONE host call for select-sort -> decode of the ordered
            # prefix -> windowed NMS (+ exact fallback)
            t("decode_filter")
            api.decode_and_filter(head, self.strides, self.conf, self.cls_mask, level_hw=self.level_hw,
                                  cap=self.cap, out=self.cands, defer_boxes=True)
            t("postprocess_dense")
            det = api.postprocess_dense(self.cands, self.ws, head, self.strides, level_hw=self.level_hw,
                                        iou_thres=self.iou, agnostic=self.agnostic, max_det=self.max_det,
                                        max_nms=self.max_nms, max_wh=self.max_wh, scale=self.scale,
                                        roi_mask=self.roi_mask, roi_nc=self.nc, roi_cnt=self.roi_cnt)
            cand_count = self.cands.count
        if not getattr(self, "roi_stage", True):     # detections only (SlicedPipeline's full-frame pass: the ROIs are cut
            t(None)                                   # from the MERGED detections); the ROI outputs keep their last contents
            return PipelineResult(self.net_in, det, cand_count, self.roi_out[0], self.roi_out[1], self.roi_out[2],
                                  self.roi_out[3], self.roi_out[4], self.cap)
        t("roi_crop_resize")
        ro = api.rois_from_detections(self._roi_frames, det, self.roi_cnt, self.roi_mask, self.nc, self.roi_cap, self.pad,
                                      self.roi_size, out=self.roi_out)
        t(None)
        return PipelineResult(self.net_in, det, cand_count, ro[0], ro[1], ro[2], ro[3], ro[4], self.cap)

    # -- per-kernel CUDA-event timing (bench.py): events on the launching stream around each stage --
    def enable_profiling(self, on=True, external=False):
        """``external=True`` records the events as external nodes so they can sit inside a captured CUDA
        graph: after ``capture()`` every ``replay()`` + synchronize refreshes ``kernel_times_ms()``."""
        self._prof = [] if on else None
        self._open = None
        self._prof_external = bool(external)

    def _tick(self, name):
        if getattr(self, "nvtx", False):             # one NVTX range per stage (tracing hooks)
            if getattr(self, "_nvtx_open", False):
                torch.cuda.nvtx.range_pop()
            self._nvtx_open = name is not None
            if name is not None:
                torch.cuda.nvtx.range_push(f"b200yolo:{name}")
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
    def capture(self, frames: torch.Tensor, head: torch.Tensor, key=0):
        """Capture one step on static input buffers; afterwards ``replay(key)`` re-runs it (copy new data into the
        tensors passed here first).  Several graphs can be captured on one ``Pipeline`` (one per ``key``: e.g. one per
        resident input batch of a stream): they share the output buffers, so replay them one after another."""
        self._static = (frames, head)
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self(frames, head)                       # warm-up: sets kernel attributes, touches buffers
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self._prof is not None:
            self._prof, self._open = [], None        # keep only the events recorded inside the graph
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            result = self(frames, head)
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        self._graphs[key] = (graph, result)
        self._graph, self._result = graph, result
        return result

    def replay(self, key=None) -> PipelineResult:
        if self._graph is None:
            raise RuntimeError("capture() first")
        if key is None:
            self._graph.replay()
            return self._result
        graph, result = self._graphs[key]
        graph.replay()
        return result

    # -- host-facing entry: pinned host buffers in, host results out -----------------------------
    def run_host(self, frames_host: torch.Tensor, head_host: torch.Tensor, staging=None):
        """End-to-end call on HOST (pinned) buffers: H2D of the frames and the head tensor, the device
        path, D2H of detections/counts.  Returns (det_rows, det_count, roi_count) host tensors."""
        if staging is None:
            staging = self.make_staging()
        d_frames, d_head, h_rows, h_count, h_roi, h_cand = staging
        d_frames.copy_(frames_host, non_blocking=True)
        d_head.copy_(head_host, non_blocking=True)
        res = self(d_frames, d_head)
        h_rows.copy_(res.det.rows, non_blocking=True)
        h_count.copy_(res.det.count, non_blocking=True)
        h_roi.copy_(res.roi_count, non_blocking=True)
        h_cand.copy_(res.cand_count, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # host results are complete when this returns ...
        check_counts(h_cand, self.cap, h_roi, self.roi_cap)   # ... and valid: raises on capacity overflow
        return h_rows, h_count, h_roi

    def launches_per_step(self):
        """Kernels of this package launched per step: letterbox, class filter, [fused post-processing | box
        decode, sort, nms], ROI crops, large-ROI pass."""
        return 5 if self.fused else (10 if self.cap > 2048 else 7)

    def check_overflow(self):
        """Raise ``CandidateOverflow`` / ``RoiOverflow`` if the last step exceeded a capacity (one D2H of the counts).
        With ``cap < A`` a candidate beyond ``cap`` is dropped, so results are only valid when this passes.
        ``run_host`` / ``HostRunner`` perform this check themselves on the counts they read back anyway."""
        cc = self.cand_seen if self.fused else self.cands.count
        return check_counts(cc.cpu(), self.cap, self.roi_out[4].cpu(), self.roi_cap)

    def make_staging(self, rows_only=False, head=True):
        dev = self.device
        no = 64 + self.nc
        rows = self.rows_plan[2] if rows_only else self.src_hw[0]
        return (torch.empty((self.B, rows, self.src_hw[1], 3), dtype=torch.uint8, device=dev),
                torch.zeros((self.B, no, self.A), dtype=torch.float32, device=dev) if head else None,
                torch.empty((self.B, self.max_det, 6), dtype=torch.float32).pin_memory(),
                torch.empty((self.B,), dtype=torch.int32).pin_memory(),
                torch.empty((1,), dtype=torch.int32).pin_memory(),
                torch.zeros((self.B,), dtype=torch.int32).pin_memory())

    def h2d_bytes_per_step(self):
        return self.B * self.src_hw[0] * self.src_hw[1] * 3 + self.B * (64 + self.nc) * self.A * 4

    def d2h_bytes_per_step(self):
        return self.B * self.max_det * 6 * 4 + 2 * self.B * 4 + 4       # detections, det counts, candidate counts, ROI total


class HostRunner:
    """End-to-end driver on HOST buffers: every step moves that step's inputs from pinned host memory (copy
    stream), runs the device path (compute stream) and reads the detections back.  ``depth`` staging sets let
    step i+1's H2D overlap step i's kernels.  PCIe is the bound of this path, so the bytes moved are the design:

    ``stage="full"``  copies whole frames (B*H*W*3) and the whole head tensor.
    ``stage="rows"``  copies only the source rows the letterbox reads (``geometry.referenced_rows``: a third of a
                      1920x1200 frame) with one strided DMA, and the ROI kernel takes its crops straight from the
                      pinned host frames (zero-copy, a few KB per ROI) -- the full-resolution frame never crosses.
    ``dfl_zero_copy`` additionally copies only the class channels of the head; the 64 DFL values of each survivor
                      are read zero-copy by the fused post-processing kernel (sparse regime only).

    With zero-copy reads the caller must leave a submitted step's host buffers untouched until that step is done
    (``wait(slot)`` / stream synchronisation), exactly as for the asynchronous copies."""

    def __init__(self, pipe: Pipeline, depth=2, head_resident=None, stage="full", dfl_zero_copy=False):
        """``head_resident``: a device head tensor to use every step instead of copying one from the host
        (the deployment case: the Detect head is produced on the device by the backbone)."""
        if stage not in ("full", "rows"):
            raise ValueError("stage must be 'full' or 'rows'")
        if dfl_zero_copy and (not pipe.fused or head_resident is not None):
            raise ValueError("dfl_zero_copy needs the fused sparse-regime path and a host head")
        self.pipe, self.depth, self.head_resident = pipe, depth, head_resident
        self.stage, self.dfl_zero_copy = stage, bool(dfl_zero_copy)
        self.staging = [pipe.make_staging(rows_only=(stage == "rows"), head=head_resident is None)
                        for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=pipe.device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.step = 0

    def submit(self, frames_host: torch.Tensor, head_host: torch.Tensor):
        """Enqueue one step; returns the (rows, count, roi_count) pinned host tensors it will fill.  Before a staging
        slot is re-used the counts of the step that last used it are checked (they are complete by then: the copy
        stream waits on that step): a capacity overflow raises ``CandidateOverflow`` / ``RoiOverflow`` here, at the
        latest in ``wait()``."""
        s = self.step % self.depth
        if self.step >= self.depth:
            self.free[s].synchronize()
            self._check(s)
        d_frames, d_head, h_rows, h_count, h_roi, h_cand = self.staging[s]
        p = self.pipe
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if self.step >= self.depth:
                self.copy_stream.wait_event(self.free[s])
            if self.stage == "rows":
                api.stage_rows_h2d(frames_host, d_frames, p.new_shape, auto=p.auto,
                                   stride=max(int(x) for x in p.strides))
            else:
                d_frames.copy_(frames_host, non_blocking=True)
            if self.head_resident is None:
                if self.dfl_zero_copy:
                    api.stage_head_classes_h2d(head_host, d_head)
                else:
                    d_head.copy_(head_host, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        compute.wait_event(self.ready[s])
        res = p(d_frames, d_head if self.head_resident is None else self.head_resident,
                roi_frames=frames_host if self.stage == "rows" else None,
                dfl_head=head_host if self.dfl_zero_copy else None)
        h_rows.copy_(res.det.rows, non_blocking=True)
        h_count.copy_(res.det.count, non_blocking=True)
        h_roi.copy_(res.roi_count, non_blocking=True)
        h_cand.copy_(res.cand_count, non_blocking=True)
        self.free[s].record(compute)
        self.step += 1
        return h_rows, h_count, h_roi

    def _check(self, s):
        st = self.staging[s]
        check_counts(st[5], self.pipe.cap, st[4], self.pipe.roi_cap)

    def wait(self, slot=None):
        """Block until the step last submitted into ``slot`` (default: the most recent) has finished; raises
        ``CandidateOverflow`` / ``RoiOverflow`` if that step exceeded a capacity (its results are invalid)."""
        s = (self.step - 1) % self.depth if slot is None else slot
        self.free[s].synchronize()
        self._check(s)

    def h2d_bytes_per_step(self):
        """Bytes moved by the H2D DMAs of a step (zero-copy reads are extra: see ``zero_copy_bytes``)."""
        p = self.pipe
        rows = p.rows_plan[2] if self.stage == "rows" else p.src_hw[0]
        n = p.B * rows * p.src_hw[1] * 3
        if self.head_resident is None:
            n += p.B * ((p.nc if self.dfl_zero_copy else 64 + p.nc)) * p.A * 4
        return n

    def zero_copy_bytes(self, det_rows, det_count):
        """Bytes the ROI kernel pulls from the pinned host frames for one step, from that step's results: the crop
        rectangles of the ROI-class detections (``stage='rows'`` only).  The DFL part of ``dfl_zero_copy`` (64 values
        per candidate, one 32-byte sector each) is added by the caller from the candidate counts."""
        if self.stage != "rows":
            return 0
        p, H, W = self.pipe, self.pipe.src_hw[0], self.pipe.src_hw[1]
        total = 0
        rows, counts = det_rows.tolist(), det_count.tolist()
        for b, n in enumerate(counts):
            for i in range(n):
                x1, y1, x2, y2, _, c = rows[b][i]
                if int(c) in p.roi_classes:
                    cx1, cy1 = max(0, min(W - 1, int(x1 - p.pad))), max(0, min(H - 1, int(y1 - p.pad)))
                    cx2, cy2 = max(0, min(W, int(x2 + p.pad))), max(0, min(H, int(y2 + p.pad)))
                    total += max(0, cx2 - cx1) * max(0, cy2 - cy1) * 3
        return total


class BatchStream:
    """``depth`` batches in flight on one GPU.  Slot s owns a ``Pipeline`` (its own output buffers), a stream and
    one captured CUDA graph; ``submit()`` replays the next slot's graph on that slot's stream, so the
    latency-bound tail of batch i (NMS, ROI crops) runs underneath the bandwidth-bound head of batch i+1
    (letterbox, class filter) instead of leaving the SMs idle.  The reference processes one frame at a time
    (``detect.py:530-600``); a production stream of frames has no such dependency between batches."""

    def __init__(self, pipes: Sequence[Pipeline]):
        self.pipes = list(pipes)
        self.depth = len(self.pipes)
        dev = self.pipes[0].device
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.pipes]
        self.step = 0

    def capture(self, inputs):
        """``inputs``: per slot, one (frames, head) pair of static device tensors (may be the same pair for every slot)
        or a list of such pairs -- the resident batches that slot will process, graph ``k`` = its k-th pair."""
        if len(inputs) != self.depth:
            raise ValueError("one (frames, head) pair (or list of pairs) per slot")
        for p, item in zip(self.pipes, inputs):
            pairs = item if isinstance(item, list) else [item]
            for k, (f, h) in enumerate(pairs):
                p.capture(f, h, key=k)

    def submit(self, key=None) -> PipelineResult:
        """Replay the next slot's graph (``key``: which of that slot's captured batches) on the slot's stream."""
        s = self.step % self.depth
        st = self.streams[s]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            res = self.pipes[s].replay(key)
        self.step += 1
        return res

    def join(self):
        """Make the current stream wait for every slot."""
        cur = torch.cuda.current_stream()
        for st in self.streams:
            cur.wait_stream(st)

    def check_overflow(self):
        """Synchronise every slot and raise ``CandidateOverflow`` / ``RoiOverflow`` if its last batch exceeded a capacity."""
        for st, p in zip(self.streams, self.pipes):
            st.synchronize()
            p.check_overflow()


class SlicedPipeline:
    """SAHI-style sliced prediction of a batch of frames (the reference calls
    ``get_sliced_prediction(frame, slice_height=640, slice_width=640, overlap ratio 0.2)``, ``pipe.py:183-194``, with
    SAHI's defaults for everything else: ``perform_standard_pred=True``, ``postprocess_type="GREEDYNMM"``,
    ``postprocess_match_metric="IOS"``, ``postprocess_match_threshold=0.5``, class-aware):

        K1 slice mode  every window of ``geometry.slice_boxes`` letterboxed + normalised -> the backbone's batch
                       (F * n_slices items; one launch); with ``standard_pred`` also the F full frames (K1)
        K2..K4         per slice (and per full frame), exactly as for a frame (boxes scaled to slice / frame pixels)
        gather         per frame: slices' detections shifted by the slice origins + the full-frame detections
        merge          ``merge="greedy_nmm"`` (SAHI's default: matched boxes are MERGED into the kept one) with
                       ``match_metric`` / ``merge_iou``; or ``merge="nms"``: one more class-aware NMS (K3 + K4)
        K5             ROI crops of the merged rank-class detections from the full frames

    The Detect-head tensors of the slices (and of the full frames) are inputs (the backbone stays torch)."""

    def __init__(self, n_frames: int, frame_hw, nc: int, slice_hw=(640, 640), overlap=(0.2, 0.2), imgsz=640, conf=0.25,
                 iou=0.7, merge_iou=0.5, max_det=300, agnostic=False, max_nms=30000, max_wh=7680,
                 roi_classes: Sequence[int] = RANK_CLASS_IDS, rois_per_frame=8, pad=6, roi_size=64, strides=(8, 16, 32),
                 device="cuda", cap=1024, merge="nms", match_metric="IOS", standard_pred=False):
        if not torch.cuda.is_available():
            raise RuntimeError("manual_yolo_b200.SlicedPipeline needs a CUDA device (no CPU fallback)")
        if merge not in ("nms", "greedy_nmm"):
            raise ValueError("merge must be 'nms' or 'greedy_nmm'")
        self.device = dev = torch.device(device)
        self.F, self.frame_hw, self.nc = int(n_frames), (int(frame_hw[0]), int(frame_hw[1])), int(nc)
        self.slices = geometry.slice_boxes(self.frame_hw[0], self.frame_hw[1], slice_hw[0], slice_hw[1], overlap[0], overlap[1])
        self.S = len(self.slices)
        self.slice_hw = (self.slices[0][3] - self.slices[0][1], self.slices[0][2] - self.slices[0][0])
        self.new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        self.strides, self.conf, self.iou, self.merge_iou = tuple(strides), conf, iou, merge_iou
        self.max_det, self.agnostic, self.max_nms, self.max_wh = max_det, agnostic, max_nms, max_wh
        self.pad, self.roi_size, self.roi_classes = pad, roi_size, tuple(roi_classes)
        self.merge, self.match_metric, self.standard_pred = merge, match_metric, bool(standard_pred)
        g = geometry.letterbox_geometry(self.slice_hw, self.new_shape, stride=max(int(s) for s in strides))
        self.in_hw = (g["out_h"], g["out_w"])
        self.level_hw = geometry.level_shapes(g["out_h"], g["out_w"], strides)
        self.A = sum(h * w for h, w in self.level_hw)
        n_items = self.F * self.S
        self.cap = int(cap or self.A)
        self.fused = self.cap <= FUSED_CAP_MAX
        self.net_in = torch.empty((n_items, 3, g["out_h"], g["out_w"]), dtype=torch.float32, device=dev)
        self.cands = api.Candidates(torch.empty((n_items, self.cap, 6), dtype=torch.float32, device=dev),
                                    torch.empty((n_items, self.cap), dtype=torch.int32, device=dev),
                                    torch.zeros((n_items,), dtype=torch.int32, device=dev), self.cap)
        self.ws = api.Workspace(n_items, self.cap, max_det, dev)               # per-slice sort/NMS
        self.scale = api.scale_params_tensor(self.in_hw, [self.slice_hw] * n_items, dev)
        # full-frame prediction (SAHI perform_standard_pred): an ordinary Pipeline over the F frames, ROI stage unused
        self.full = Pipeline(self.F, self.frame_hw, nc, imgsz=imgsz, conf=conf, iou=iou, max_det=max_det, agnostic=agnostic,
                             max_nms=max_nms, max_wh=max_wh, roi_classes=roi_classes, rois_per_frame=1, strides=strides,
                             device=dev, cap=cap) if self.standard_pred else None
        if self.full is not None:
            self.full.roi_stage = False                                        # its ROI kernels would be wasted work
            self._full_stream = torch.cuda.Stream(device=dev)                  # the full-frame pass runs beside the slices
        mcap = (self.S + (1 if self.standard_pred else 0)) * max_det
        self.mcands = api.Candidates(torch.empty((self.F, mcap, 6), dtype=torch.float32, device=dev),
                                     torch.empty((self.F, mcap), dtype=torch.int32, device=dev),
                                     torch.zeros((self.F,), dtype=torch.int32, device=dev), mcap)
        self.mws = api.Workspace(self.F, mcap, max_det, dev)                   # merge sort/NMS (or the NMM output)
        self.roi_cap = max(1, self.F * int(rois_per_frame))
        self.roi_mask = api._class_mask(self.roi_classes, self.nc, dev)
        self.roi_cnt = torch.zeros((self.F,), dtype=torch.int32, device=dev)
        self.cand_seen = torch.zeros((n_items,), dtype=torch.int32, device=dev)
        self.roi_out = (torch.zeros((self.roi_cap, 3, roi_size, roi_size), dtype=torch.float32, device=dev),
                        torch.zeros((self.roi_cap,), dtype=torch.int32, device=dev),
                        torch.zeros((self.roi_cap,), dtype=torch.int32, device=dev),
                        torch.zeros((self.roi_cap,), dtype=torch.int32, device=dev),
                        torch.zeros((1,), dtype=torch.int32, device=dev))

    def preprocess(self, frames: torch.Tensor) -> torch.Tensor:
        """K1 slice mode alone: the (F*S,3,h,w) batch the backbone consumes."""
        if tuple(frames.shape) != (self.F, self.frame_hw[0], self.frame_hw[1], 3):
            raise ValueError(f"frames must be {(self.F, *self.frame_hw, 3)}, got {tuple(frames.shape)}")
        return api.preprocess_slices(frames, self.slices, self.new_shape, stride=max(int(s) for s in self.strides),
                                     out=self.net_in)

    def __call__(self, frames: torch.Tensor, head, head_full=None) -> PipelineResult:
        """frames (F,H,W,3) uint8 BGR on the device; head (F*S, 64+nc, A): item f*S+s = slice s of frame f;
        ``head_full`` (F, 64+nc, A_full): the Detect head of the letterboxed full frames (``standard_pred=True``).
        Returns the MERGED per-frame detections (frame pixels); ``det.anchor`` = slice * max_det + rank of the kept
        box (slice index S = the full-frame prediction)."""
        full_det = None
        if self.standard_pred:
            if head_full is None:
                raise ValueError("standard_pred=True needs head_full: the Detect head of the letterboxed full frames")
            cur = torch.cuda.current_stream()
            self._full_stream.wait_stream(cur)                                 # fork: independent of the slice branch
            with torch.cuda.stream(self._full_stream):
                self.full_result = self.full(frames, head_full)
            full_det = self.full_result.det
        self.preprocess(frames)
        if self.fused:
            api.decode_and_filter(head, self.strides, self.conf, level_hw=self.level_hw, cap=self.cap, out=self.cands,
                                  defer_boxes=True, zero=False)
            det = api.postprocess_small(self.cands, self.ws.det, head, self.strides, level_hw=self.level_hw,
                                        iou_thres=self.iou, agnostic=self.agnostic, max_nms=self.max_nms,
                                        max_wh=self.max_wh, scale=self.scale, cand_seen=self.cand_seen)
            cand_count = self.cand_seen
        else:
            api.decode_and_filter(head, self.strides, self.conf, level_hw=self.level_hw, cap=self.cap, out=self.cands)
            api.sort_candidates(self.cands, self.max_nms, self.ws)
            det = api.nms_sorted(self.cands, self.ws, self.iou, self.agnostic, self.max_det, self.max_nms, self.max_wh,
                                 scale=self.scale)
            cand_count = self.cands.count
        self.slice_det = det
        if self.standard_pred:
            torch.cuda.current_stream().wait_stream(self._full_stream)         # join
        api.gather_slice_detections(det, self.slices, self.F, out=self.mcands, full_det=full_det)
        if self.merge == "greedy_nmm":
            merged = api.greedy_nmm(self.mcands, self.mws.det, self.match_metric, self.merge_iou, self.agnostic,
                                    roi_mask=self.roi_mask, roi_nc=self.nc, roi_cnt=self.roi_cnt)
        else:
            api.sort_candidates(self.mcands, self.max_nms, self.mws)
            merged = api.nms_sorted(self.mcands, self.mws, self.merge_iou, self.agnostic, self.max_det, self.max_nms,
                                    self.max_wh, roi_mask=self.roi_mask, roi_nc=self.nc, roi_cnt=self.roi_cnt)
        ro = api.rois_from_detections(frames, merged, self.roi_cnt, self.roi_mask, self.nc, self.roi_cap, self.pad,
                                      self.roi_size, out=self.roi_out)
        return PipelineResult(self.net_in, merged, cand_count, ro[0], ro[1], ro[2], ro[3], ro[4], self.cap)

    def check_overflow(self):
        """Raise ``CandidateOverflow`` / ``RoiOverflow`` if the last call exceeded a capacity (one D2H of the counts)."""
        cc = self.cand_seen if self.fused else self.cands.count
        if self.standard_pred:
            check_counts((self.full.cand_seen if self.full.fused else self.full.cands.count).cpu(), self.full.cap)
        return check_counts(cc.cpu(), self.cap, self.roi_out[4].cpu(), self.roi_cap)


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
