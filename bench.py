#!/usr/bin/env python
"""bench.py -- post-processed frames/s of the detection hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the whole path (letterbox -> class filter -> decode + sort + NMS + scale_boxes -> ROI crops)
over one batch of 64 synthetic 1920x1200 BGR frames + the matching (64,128,8400) Detect-head tensor (BASELINE.json
configs[1]; the detector weights are absent from the reference, so the head is the label-derived synthetic one of
manual_yolo_b200/synth.py).  Every rank owns a resident stream of 16 DISTINCT batches (configs[4]: 1024 frames per
GPU; batch k of step i is k = i mod 16) -- frames shard across ranks with no data-path collective (weak scaling).

One JSON line on stdout (rank 0):
  value          frames/s, inputs resident in HBM: MEDIAN over R repetitions of the K-step block, each block bracketed
                 by barrier + synchronize and timed with CUDA events on the launching stream, max over ranks
  parity         the GPU results of rank 0's batch 0 compared with the CPU oracle on the same bits (exit 3 on a
                 kept-set mismatch)
  e2e            the same metric through HostRunner with pinned HOST buffers (H2D of frames + head, D2H of results)
  roofline       the dominant kernel (K1 letterbox) against the measured HBM peak
  configs        BASELINE configs[0], [2], [3] timed in the same run (graph us; class filter / select-sort / NMS us per
                 image and per batch, chain us + GB/s; ROI us + GB/s) and configs[4] (1024 frames per GPU from a host
                 barrier to the last count read-back, then the host-side columnar gather)
  cpu_baseline   the oracle port timed on this box's host cores (all threads, and one thread)
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "post-processed frames/sec"
UNIT = "frames/s"
BATCH, SRC_HW, NC, IMGSZ = 64, (1200, 1920), 64, 640
CONF, IOU, MAX_DET = 0.25, 0.45, 300
STREAM_BATCHES = 16                       # configs[4]: 1024 frames per GPU = 16 batches of 64
WORKLOAD = ("configs[1]: synthetic 1920x1200 BGR frames, batch=64, 8400 anchors, nc=64 poker head "
            "(label-derived synthetic, poker_model.pt absent), conf=0.25 iou=0.45 max_det=300, "
            "letterbox+decode+sort+NMS+scale_boxes+ROI crops 64x64")


def _config(n_gpus):
    """Identical keys and values in both arms (ours / --impl reference)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "frame_hw": list(SRC_HW),
            "anchors": 8400, "nc": NC, "conf": CONF, "iou": IOU, "max_det": MAX_DET,
            "parallelism": f"frame-sharded x{n_gpus}, no collective",
            "l2_policy": "inputs larger than L2: every step reads its own 717 MB batch (16 distinct resident batches per "
                         "GPU, 126 MB L2)"}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---- clocks sampled DURING the timed region -------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except Exception:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference's CPU path (oracle port) ---------------------------------------------------------
def oracle_pass(frames, head, level_hw, in_hw, keep_outputs=False):
    """One pass of the oracle (real cv2 / torch CPU / torchvision.ops.nms / PIL leaves + restated UL glue) over
    `frames` (numpy (n,H,W,3)) + `head` (torch (n,128,A)).  keep_outputs: return what the parity check compares."""
    from oracle import boxes as oboxes
    from oracle import head as ohead
    from oracle import letterbox as olb
    from oracle import nms as onms
    from oracle import roi as oroi
    from manual_yolo_b200.pipeline import RANK_CLASS_IDS
    H, W = frames.shape[1:3]
    net_in = olb.preprocess_ref(list(frames), (IMGSZ, IMGSZ))
    pred = ohead.detect_inference_ref(head, level_hw)
    out, idx = onms.non_max_suppression_ref(pred, CONF, IOU, max_det=MAX_DET, return_idxs=True)
    dets, rois, where = [], [], []
    for b, o in enumerate(out):
        o = o.clone()
        o[:, :4] = oboxes.scale_boxes_ref(in_hw, o[:, :4], (H, W))
        dets.append(o)
        for i, row in enumerate(o):
            if int(row[5]) in RANK_CLASS_IDS:
                crop = oboxes.safe_crop_ref(frames[b], *[int(v) for v in row[:4]], pad=6)
                if crop is not None:
                    t = oroi.classify_preprocess_ref(crop)
                    if keep_outputs:
                        rois.append(t)
                        where.append((b, i))
    if keep_outputs:
        return {"net_in": net_in, "dets": dets, "idx": idx, "rois": rois, "where": where}
    return None


def cpu_path_frames_per_s(frames, head, level_hw, in_hw, n_timed_batches=1, warm=True, threads=None):
    """Times oracle_pass on the host cores: `threads` torch + cv2 threads (default: all)."""
    import cv2
    cores = threads or (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    if warm:
        oracle_pass(frames, head, level_hw, in_hw)
    t0 = time.perf_counter()
    for _ in range(n_timed_batches):
        oracle_pass(frames, head, level_hw, in_hw)
    dt = time.perf_counter() - t0
    return frames.shape[0] * n_timed_batches / dt, cores, dt


def parity_report(res, ref, frames_shape):
    """GPU step result vs the oracle outputs of the same batch: kept sets (anchor indices, in order) and class ids must
    be bit-equal; boxes/scores within 1e-4; ROI tensors within 1/255; the letterboxed network input bit-equal."""
    B = frames_shape[0]
    counts = res.det.count.cpu().tolist()
    rows, anchors = res.det.rows.cpu(), res.det.anchor.cpu()
    kept_equal = class_equal = True
    max_box = max_score = 0.0
    n_det = 0
    for b in range(B):
        exp, idx = ref["dets"][b], ref["idx"][b]
        k = counts[b]
        n_det += exp.shape[0]
        if k != exp.shape[0] or not torch.equal(anchors[b, :k].long(), idx):
            kept_equal = False
            continue
        got = rows[b, :k]
        class_equal = class_equal and bool(torch.equal(got[:, 5], exp[:, 5]))
        if k:
            max_box = max(max_box, float((got[:, :4] - exp[:, :4]).abs().max()))
            max_score = max(max_score, float((got[:, 4] - exp[:, 4]).abs().max()))
    n = res.n_rois()
    where = list(zip(res.roi_batch[:n].cpu().tolist(), res.roi_det[:n].cpu().tolist()))
    roi_sets_equal = where == ref["where"][:n] and int(res.roi_count) == len(ref["where"])
    roi_max = 0.0
    if roi_sets_equal and n:
        got = res.rois[:n].cpu()
        roi_max = max(float((got[i] - ref["rois"][i]).abs().max()) for i in range(n))
    letterbox_equal = bool(torch.equal(res.net_in.cpu(), ref["net_in"]))
    ok = (kept_equal and class_equal and roi_sets_equal and letterbox_equal and max_box <= 1e-4 and max_score <= 1e-4
          and roi_max <= 1.0 / 255.0)
    return {"checked": "rank 0, batch 0 (64 frames, seed 0): GPU step vs the CPU oracle on the same bits",
            "ok": ok, "kept_sets_equal": kept_equal, "class_ids_equal": class_equal, "letterbox_bit_equal": letterbox_equal,
            "max_abs_box": max_box, "max_abs_score": max_score, "roi_sets_equal": roi_sets_equal, "roi_max_abs": roi_max,
            "n_det": n_det, "n_roi": len(ref["where"]),
            "tolerances": {"kept/class": "bit-exact", "box/score": 1e-4, "roi": 1.0 / 255.0},
            "oracle": "parity unpinned for rows a1-a11 (no golden detections exist upstream); ROI leg pinned by the 63/67 KAT"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (Ultralytics is not installable
    here, so the oracle port is what runs), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from manual_yolo_b200 import geometry, synth
    sample = BATCH                               # frames per step: the full 64-frame batch (~0.25-0.5 s of CPU work)
    in_hw = (IMGSZ, IMGSZ)
    level_hw = geometry.level_shapes(*in_hw)
    frames = synth.synth_frames(sample, *SRC_HW, seed=0).numpy()
    head, _ = synth.synth_head_from_labels(sample, NC, in_hw=in_hw, src_hw=SRC_HW, seed=0, conf_thres=CONF)
    for _ in range(max(1, args.warmup)):
        cpu_path_frames_per_s(frames, head, level_hw, in_hw, 1, warm=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fps, cores, _ = cpu_path_frames_per_s(frames, head, level_hw, in_hw, 1, warm=False)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args.gpus), "sample_frames_per_step": sample,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                             "sample": f"{sample} frames per step (the full batch), {args.steps} steps, oracle port "
                                       "(cv2/torch-CPU/torchvision.nms/PIL leaves), all host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _events(n):
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def run_ours(args, rank, world, local):
    import manual_yolo_b200 as m
    from manual_yolo_b200 import multigpu, synth
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = multigpu.bind_to_gpu_numa_node(local) if world > 1 else None   # pinned host buffers local to the GPU
    K, W, R = args.steps, max(3, args.warmup), max(1, args.reps)
    mk = lambda **kw: m.Pipeline(BATCH, SRC_HW, NC, imgsz=IMGSZ, conf=CONF, iou=IOU, max_det=MAX_DET, device=dev,
                                 cap=args.cap or None, **kw)
    pipe = mk()

    # ---- inputs: batch 0 is generated on the CPU (seed base + rank) so the oracle sees the same bits; the other
    #      15 batches of the rank's 1024-frame stream are device-generated noise frames (seeded) + the two CPU-made
    #      heads rolled along the batch axis (every frame slot sees a different head in every batch) ----
    seed = args.seed + rank
    frames_h = synth.synth_frames(BATCH, *SRC_HW, seed=seed).pin_memory()
    head_h, _ = synth.synth_head_from_labels(BATCH, NC, in_hw=pipe.in_hw, src_hw=SRC_HW, seed=seed, conf_thres=CONF)
    head_h = head_h.pin_memory()
    head2_h, _ = synth.synth_head_from_labels(BATCH, NC, in_hw=pipe.in_hw, src_hw=SRC_HW, seed=seed + 1000, conf_thres=CONF)
    NB = STREAM_BATCHES
    frames_all = torch.empty((NB, BATCH) + SRC_HW + (3,), dtype=torch.uint8, device=dev)
    heads_all = torch.empty((NB, BATCH, 64 + NC, pipe.A), dtype=torch.float32, device=dev)
    frames_all[0].copy_(frames_h)
    g = torch.Generator(device=dev)
    g.manual_seed(1_000_003 * (seed + 1))
    for k in range(1, NB):
        frames_all[k].copy_(torch.randint(0, 256, (BATCH,) + SRC_HW + (3,), dtype=torch.uint8, device=dev, generator=g))
    h0, h1 = head_h.to(dev), head2_h.to(dev)
    for k in range(NB):
        heads_all[k].copy_(torch.roll(h0 if k % 2 == 0 else h1, shifts=k // 2, dims=0))
    del h1
    frames_d, head_d = frames_all[0], heads_all[0]
    torch.cuda.synchronize()

    # ---- parity: this batch through the GPU path vs the oracle (rank 0) ----
    parity, ref = None, None
    res = pipe(frames_d, head_d)
    torch.cuda.synchronize()
    max_cand = pipe.check_overflow()             # raises CandidateOverflow / RoiOverflow: the run would be invalid
    if rank == 0 and not args.no_parity:
        ref = oracle_pass(frames_h.numpy(), head_h.clone(), pipe.level_hw, pipe.in_hw, keep_outputs=True)
        parity = parity_report(res, ref, frames_h.shape)
    n_det = int(res.det.count.sum())

    # ---- per-kernel timing (a): eager launches with CUDA events around every kernel ----
    for _ in range(W):
        pipe(frames_d, head_d)
    torch.cuda.synchronize()
    pipe.enable_profiling(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    multigpu.barrier()
    torch.cuda.synchronize()
    e0, e1 = _events(2)
    t0 = time.time()
    e0.record()
    for i in range(K):
        pipe(frames_all[i % NB], heads_all[i % NB])
    e1.record()
    torch.cuda.synchronize()
    multigpu.barrier()
    ms_eager = multigpu.max_over_ranks(e0.elapsed_time(e1), dev)
    kt_eager = pipe.kernel_times_ms()
    pipe.enable_profiling(False)
    # (a2) the dominant kernel (K1) alone: R blocks of K back-to-back launches between two events on the launching
    #      stream, a different 442 MB input batch per launch (inputs + outputs 757 MB > L2)
    net_alt = torch.empty_like(pipe.net_in)
    for i in range(3):
        m.preprocess(frames_all[i], (IMGSZ, IMGSZ), out=pipe.net_in)
    torch.cuda.synchronize()
    k1_blocks = []
    for r in range(R):
        a, b = _events(2)
        a.record()
        for i in range(K):
            m.preprocess(frames_all[(r * K + i) % NB], (IMGSZ, IMGSZ), out=pipe.net_in if i % 2 == 0 else net_alt)
        b.record()
        torch.cuda.synchronize()
        k1_blocks.append(a.elapsed_time(b) / K)
    del net_alt
    k1_alone_ms, k1_alone_min = statistics.median(k1_blocks), min(k1_blocks)
    # (b) instrumented single-stream graph (external event nodes around every kernel) for the per-kernel breakdown
    pipe.enable_profiling(True, external=True)
    pipe.capture(frames_d, head_d)
    kt = {}
    for it in range(W + K):
        pipe.replay()
        torch.cuda.synchronize()
        if it >= W:
            for k, v in pipe.kernel_times_ms().items():
                kt.setdefault(k, []).extend(v)
    pipe.enable_profiling(False)

    def blocks(submit, join, reps=R):
        """reps repetitions of the K-step block; each block: barrier + synchronize, event, K steps, event,
        synchronize + barrier.  Returns per-block ms (max over ranks)."""
        out = []
        for r in range(reps):
            torch.cuda.synchronize()
            multigpu.barrier()
            a, b = _events(2)
            a.record()
            for i in range(K):
                submit(r * K + i)
            join()
            b.record()
            torch.cuda.synchronize()
            multigpu.barrier()
            out.append(a.elapsed_time(b))
        return multigpu.max_over_ranks_list(out, dev)

    # serial (single-stream) graph: value_serial_graph
    pipe.capture(frames_d, head_d)
    for _ in range(W):
        pipe.replay()
    ms_serial = statistics.median(blocks(lambda i: pipe.replay(), lambda: None, reps=min(R, 5)))
    # production graphs: the letterbox branch (HBM-bound) forked onto a side stream, concurrent with the
    # decode -> NMS -> ROI branch (latency-bound); `depth` batches in flight (BatchStream: slot s = its own Pipeline
    # buffers, stream and one graph per resident batch of its parity); step i replays batch i mod 16
    def timed_stream(depth, reps=R):
        pipes = [mk(overlap=True) for _ in range(depth)]
        bs = m.BatchStream(pipes)
        per_slot = [[(frames_all[k], heads_all[k]) for k in range(s, NB, depth)] for s in range(depth)] \
            if NB % depth == 0 else [[(frames_all[s], heads_all[s])] for s in range(depth)]
        nkeys = len(per_slot[0])
        bs.capture(per_slot)
        for i in range(W * depth):
            bs.submit((i // depth) % nkeys)
        bs.join()
        ms = blocks(lambda i: bs.submit((i // depth) % nkeys), bs.join, reps)
        bs.check_overflow()
        return ms, bs
    ms1_blocks, bs1 = timed_stream(1, reps=min(R, 5))
    del bs1
    ms2_blocks, bs2 = timed_stream(2)
    t1 = time.time()
    # results of the in-flight run for batch 0 must equal the single-pipeline result
    bs2.step = 0
    r0 = bs2.submit(0)
    bs2.join()
    torch.cuda.synchronize()
    cnt0 = res.det.count.cpu().tolist()
    assert torch.equal(r0.det.count, res.det.count)
    assert all(torch.equal(r0.det.rows[b, :k], res.det.rows[b, :k]) for b, k in enumerate(cnt0))

    # ---- configs[4]: the rank's 1024-frame stream from a host barrier to the last count read-back, then the
    #      host-side (columnar) gather of the detection records ----
    h_rows = torch.empty((NB, BATCH, MAX_DET, 6), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((NB, BATCH), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()
    multigpu.barrier()
    w0 = time.perf_counter()
    a, b = _events(2)
    a.record()
    bs2.step = 0
    for k in range(NB):
        s = k % 2
        r_k = bs2.submit(k // 2)
        with torch.cuda.stream(bs2.streams[s]):      # read-back on the slot's stream: overlaps the other slot's kernels
            h_rows[k].copy_(r_k.det.rows, non_blocking=True)
            h_cnt[k].copy_(r_k.det.count, non_blocking=True)
    bs2.join()
    b.record()
    torch.cuda.synchronize()
    c5_wall = time.perf_counter() - w0
    c5_wall_max = multigpu.max_over_ranks(c5_wall, dev)
    c5_dev_ms = multigpu.max_over_ranks(a.elapsed_time(b), dev)
    multigpu.gather_columnar(multigpu.detections_columnar(h_rows[0, :1].numpy(), h_cnt[0, :1].numpy()), dst=0)  # warm-up
    multigpu.barrier()
    g0 = time.perf_counter()
    col = multigpu.detections_columnar(h_rows.view(NB * BATCH, MAX_DET, 6).numpy(), h_cnt.view(-1).numpy(),
                                       frame_offset=rank * NB * BATCH)
    merged = multigpu.gather_columnar(col, dst=0)
    g_s = multigpu.max_over_ranks(time.perf_counter() - g0, dev)
    n_records = int(len(merged)) if merged is not None else 0
    del bs2, h_rows, h_cnt

    # ---- end to end on host buffers (H2D of the inputs, D2H of the detections, every step) ----
    def timed_e2e(p, fh, hh, reps, **kw):
        runner = m.HostRunner(p, depth=2, **kw)
        resident = kw.get("head_resident") is not None
        for _ in range(3):
            o = runner.submit(fh, None if resident else hh)
        runner.wait()
        ms = blocks(lambda i: runner.submit(fh, None if resident else hh), lambda: None, reps)
        runner.wait()
        return statistics.median(ms), min(ms), runner, o
    e2e_modes = {}
    Re = max(1, min(R, args.e2e_reps))
    for name, kw in (("full", dict(stage="full")), ("rows", dict(stage="rows")),
                     ("rows_dfl_zero_copy", dict(stage="rows", dfl_zero_copy=True))):
        if kw.get("dfl_zero_copy") and not pipe.fused:
            continue
        ms_e, ms_e_min, runner, out = timed_e2e(pipe, frames_h, head_h, Re if name == args.e2e_mode else 1, **kw)
        zc = runner.zero_copy_bytes(out[0], out[1])
        if kw.get("dfl_zero_copy"):
            zc += int(pipe.cand_seen.clamp(max=pipe.cap).sum()) * 64 * 32      # 64 DFL values, one 32-B sector each
        e2e_modes[name] = {"value": BATCH * K * world / (ms_e / 1e3), "ms_per_step": ms_e / K,
                           "ms_per_step_min": ms_e_min / K, "h2d_bytes_per_step": runner.h2d_bytes_per_step(),
                           "zero_copy_bytes_per_step": zc}
    e2e_pick = args.e2e_mode if args.e2e_mode in e2e_modes else "rows"  # dense regime (cap > 1024): no zero-copy DFL
    # the deployment case: frames from the host, the head already on the device (a backbone produces it there)
    fo_ms, fo_min, runner2, _ = timed_e2e(pipe, frames_h, head_h, Re, stage="rows", head_resident=head_d)
    # concurrent plain H2D calibration: every rank copies a 256 MB pinned buffer with one contiguous cudaMemcpyAsync per
    # repetition at the same time -- the box's host->device ceiling next to what the e2e path moves
    cal_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    cal_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cal_d.copy_(cal_h, non_blocking=True)
    torch.cuda.synchronize()
    multigpu.barrier()
    a, b = _events(2)
    a.record()
    for _ in range(8):
        cal_d.copy_(cal_h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    h2d_cal = 8 * (256 << 20) / (a.elapsed_time(b) / 1e3) / 1e9
    h2d_cal_min = -multigpu.max_over_ranks(-h2d_cal, dev)
    h2d_cal_sum = multigpu.sum_over_ranks(h2d_cal, dev)
    del cal_h, cal_d
    # e2e at the other production geometry, 1600x900 (scale 2.5: every source row is referenced, nothing to skip)
    e2e_other = None
    if not args.no_extra and world == 1:
        p9 = m.Pipeline(BATCH, (900, 1600), NC, imgsz=IMGSZ, conf=CONF, iou=IOU, max_det=MAX_DET, device=dev, cap=args.cap or None)
        f9 = synth.synth_frames(BATCH, 900, 1600, seed=seed).pin_memory()
        h9, _ = synth.synth_head_from_labels(BATCH, NC, in_hw=p9.in_hw, src_hw=(900, 1600), seed=seed, conf_thres=CONF)
        h9 = h9.pin_memory()
        h9d = h9.to(dev)
        ms9, _, r9, _ = timed_e2e(p9, f9, h9, 1, stage="rows", dfl_zero_copy=p9.fused)
        ms9f, _, r9f, _ = timed_e2e(p9, f9, h9, 1, stage="rows", head_resident=h9d)
        e2e_other = {"frame_hw": [900, 1600], "value": BATCH * K / (ms9 / 1e3), "h2d_bytes_per_step": r9.h2d_bytes_per_step(),
                     "frames_only": {"value": BATCH * K / (ms9f / 1e3), "h2d_bytes_per_step": r9f.h2d_bytes_per_step()},
                     "note": "1600x900 -> 640x360 is scale 2.5: all 900 rows are staged (no row skipping as at 1920x1200)"}
        del p9, f9, h9, h9d, r9, r9f
    t2 = time.time()
    clocks = sampler.stop(t0, t2) if rank == 0 else None

    if rank != 0:
        return 0
    # ---- the other BASELINE configs, same run (rank 0) ----
    configs = {"config5_stream": {
        "frames_per_gpu": NB * BATCH, "batches": NB, "n_gpus": world,
        "wall_s_barrier_to_last_readback": c5_wall_max, "device_ms": c5_dev_ms,
        "frames_per_s": NB * BATCH * world / c5_wall_max,
        "gather": {"records": n_records, "seconds": g_s, "records_per_s": n_records / g_s if g_s > 0 else None,
                   "format": "numpy structured array per rank (frame, x1, y1, x2, y2, conf, class_id), one gather_object"}}}
    if not args.no_extra and world == 1:
        del frames_all, heads_all
        torch.cuda.empty_cache()
        import bench_extra
        configs["config1_single_frame"] = bench_extra.config1(dev)
        c3 = bench_extra.config3(dev, B=256, cpu_images=0)
        configs["config3_nms_heavy"] = c3
        configs["config4_roi_4096"] = bench_extra.config4(dev)
        # what plain torch fill / read kernels reach on THIS box for the step's bytes (context for the roofline fractions)
        cal = bench_extra.calibration(dev)
        wr, rd = cal["fill_315MB"]["GBps"], cal["read_reduce_1GiB_f32"]["GBps"]
        written = BATCH * 3 * IMGSZ * IMGSZ * 4
        calibration = {"fill_GBps": wr, "read_reduce_GBps": rd, "copy_GBps": cal["copy_1GiB"]["GBps"],
                       "what": "torch fill_ of 315 MB / sum of 1 GiB / copy_ of 1 GiB on this GPU, cold L2",
                       "bytes_written_per_step": written}
    else:
        calibration = None

    # ---- roofline of the dominant kernel (K1 letterbox), algorithmic bytes per launch ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gm = pipe.geom
    k1_bytes_frame = gm["new_h"] * SRC_HW[1] * 3 + 3 * gm["out_h"] * gm["out_w"] * 4     # 7 219 200 B
    k1_graph_ms = statistics.mean(kt["letterbox"])            # between external event nodes inside a serialised graph
    achieved = BATCH * k1_bytes_frame / (k1_alone_ms / 1e3) / 1e9
    achieved_graph = BATCH * k1_bytes_frame / (k1_graph_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("letterbox_bytes_per_launch")
    except Exception:
        pass
    ms_med, ms_min = statistics.median(ms2_blocks), min(ms2_blocks)
    ms1_med = statistics.median(ms1_blocks)
    total_frames = BATCH * K * world
    value = total_frames / (ms_med / 1e3)
    survivors = float(sum(min(c, pipe.cap) for c in res.cand_count.cpu().tolist())) / BATCH
    # bytes a step must move per frame.  K2 is lazy by design: it reads the nc class channels of every anchor and the
    # 64 DFL channels (one 32-byte sector each = 256 B... x 8: 64 values in 64 different channel rows) of the survivors only
    k2_required = NC * pipe.A * 4 + survivors * 64 * 32
    k2_definition = (64 + NC) * pipe.A * 4
    rois_per_frame = len(ref["where"]) / BATCH if ref is not None else res.n_rois() / BATCH
    k5_required = rois_per_frame * (49152 + 12000)
    required = k1_bytes_frame + k2_required + 10_000 + k5_required
    cal_floor = None
    if calibration:
        # the step's bytes at the rates plain streaming kernels reach on this box: net_in written at the fill rate, all
        # other bytes (frames rows, class channels, survivors' sectors, ROI traffic) at the read rate
        wbytes = calibration["bytes_written_per_step"]
        floor_s = wbytes / (calibration["fill_GBps"] * 1e9) + (required * BATCH - wbytes) / (calibration["read_reduce_GBps"] * 1e9)
        cal_floor = {"us_per_step": floor_s * 1e6, "frames_per_s_per_gpu": BATCH / floor_s}
    definition = k1_bytes_frame + k2_definition + 10_000 + 4 * 60_000                          # 11.77 MB
    kernels = {}
    bytes_per_launch = {"letterbox": BATCH * k1_bytes_frame, "decode_filter": BATCH * k2_required}
    for name, v in kt.items():
        kernels[name] = {"us": 1e3 * statistics.mean(v), "us_eager_launch": 1e3 * statistics.mean(kt_eager[name])}
        if name in bytes_per_launch:
            kernels[name]["required_GBps"] = bytes_per_launch[name] / (statistics.mean(v) / 1e3) / 1e9
            kernels[name]["frac_of_measured_peak"] = kernels[name]["required_GBps"] / peak
    kernels["decode_filter"]["note"] = ("K2 is lazy: it reads the nc class channels of every anchor + the DFL channels of the "
                                        "survivors only; bytes counted = nc*A*4 + 64 sectors x 32 B per survivor")
    # ---- CPU baseline (oracle port) on a bounded sample of the same workload ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        n, passes = BATCH, 12
        fps, cores, dt = cpu_path_frames_per_s(frames_h[:n].numpy(), head_h[:n].clone(), pipe.level_hw, pipe.in_hw, passes)
        fps1, _, dt1 = cpu_path_frames_per_s(frames_h[:16].numpy(), head_h[:16].clone(), pipe.level_hw, pipe.in_hw, 2,
                                             threads=1)
        cpu_path_frames_per_s(frames_h[:2].numpy(), head_h[:2].clone(), pipe.level_hw, pipe.in_hw, 1, warm=False)  # threads back
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
               "one_thread": {"value": fps1, "cores": 1, "sample": f"16 frames, 1 warm-up + 2 timed passes ({dt1:.1f} s)"},
               "sample": f"the rank-0 batch of {n} frames, 1 warm-up + {passes} timed passes ({dt:.1f} s), oracle port "
                         "(cv2 / torch CPU / torchvision.ops.nms / PIL), torch+cv2 threads = all cores"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": _config(world),
            "timing": {"repetitions": len(ms2_blocks), "block": f"{K} steps between barrier+synchronize, CUDA events, max over ranks",
                       "ms_per_step_median": ms_med / K, "ms_per_step_min": ms_min / K,
                       "ms_per_step_max": max(ms2_blocks) / K, "timed_ms_total": sum(ms2_blocks),
                       "value_best_block": total_frames / (ms_min / 1e3)},
            "launch_mode": "one CUDA-graph replay per step, two batches in flight (BatchStream: one stream + output buffers "
                           "per slot, one graph per resident batch); inside a graph the letterbox is forked onto a side "
                           "stream, concurrent with class filter -> fused post-processing -> ROI crops",
            "value_one_in_flight": total_frames / (ms1_med / 1e3),
            "value_eager_launches": total_frames / (ms_eager / 1e3), "value_serial_graph": total_frames / (ms_serial / 1e3),
            "parity": parity,
            "e2e": {"value": e2e_modes[e2e_pick]["value"], "unit": UNIT,
                    "h2d_bytes_per_step": e2e_modes[e2e_pick]["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": pipe.d2h_bytes_per_step(), "ms_per_step": e2e_modes[e2e_pick]["ms_per_step"],
                    "mode": e2e_pick, "zero_copy_bytes_per_step": e2e_modes[e2e_pick]["zero_copy_bytes_per_step"],
                    "frames_only": {"value": total_frames / (fo_ms / 1e3), "ms_per_step": fo_ms / K,
                                    "h2d_bytes_per_step": runner2.h2d_bytes_per_step(),
                                    "note": "the deployment case: the Detect head is produced on the device by the backbone; only "
                                            "frames cross PCIe (stage='rows')"},
                    "note": "HostRunner: pinned host frames+head -> device path -> detections back, double-buffered; "
                            "'rows' stages only the source rows the letterbox reads (one strided DMA: at 1920x1200 -> 640x400 "
                            "one row in three) and the ROI kernel crops zero-copy from the pinned frames; 'rows_dfl_zero_copy' "
                            "also stages only the class channels of the head, the survivors' DFL values being read zero-copy by "
                            "the fused post-processing kernel; 'full' copies whole frames + head.  The headline mode benefits "
                            "from the 1:3 row skip of this geometry and from a host-resident head: see frames_only and "
                            "other_geometry for the numbers without them",
                    "modes": e2e_modes, "other_geometry": e2e_other,
                    "h2d_calibration": {"per_rank_GBps_min": h2d_cal_min, "aggregate_GBps": h2d_cal_sum, "ranks": world,
                                        "what": "all ranks concurrently: contiguous 256 MB pinned -> device cudaMemcpyAsync x8",
                                        "achieved_GBps_aggregate": (e2e_modes[e2e_pick]["h2d_bytes_per_step"] +
                                                                    e2e_modes[e2e_pick]["zero_copy_bytes_per_step"]) * world
                                        / (e2e_modes[e2e_pick]["ms_per_step"] / 1e3) / 1e9}},
            "gpu_launches": pipe.launches_per_step() * K * len(ms2_blocks) * world,   # timed (graph) blocks only
            "gpu_launches_per_step": pipe.launches_per_step(),
            "roofline": {"kernel": "letterbox_kernel<float> (K1)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": BATCH * k1_bytes_frame, "avg_launch_us": 1e3 * k1_alone_ms,
                         "min_block_launch_us": 1e3 * k1_alone_min,
                         "timed_in": f"median over {R} blocks of {K} back-to-back launches of the kernel alone between two CUDA "
                                     "events on the launching stream, a different input batch per launch (757 MB per launch > L2)",
                         "in_graph": {"avg_launch_us": 1e3 * k1_graph_ms, "achieved": achieved_graph, "frac": achieved_graph / peak,
                                      "timed_in": "external event nodes around the kernel inside a single-stream CUDA graph "
                                                  "(includes the event-record nodes' own latency)"}},
            "kernels": kernels,
            "calibration": calibration,
            "pipeline_roofline": {"required_bytes_per_frame": required,
                                  "roofline_frames_per_s_per_gpu": peak * 1e9 / required,
                                  "frac": (value / world) / (peak * 1e9 / required),
                                  "calibrated_floor": None if cal_floor is None else
                                  {**cal_floor, "frac": (value / world) / cal_floor["frames_per_s_per_gpu"],
                                   "what": "the required bytes at this box's torch fill (writes) and read-reduce (reads) "
                                           "rates from `calibration`, reads and writes serialised on the DRAM bus"},
                                  "what": "bytes the step must move: K1 7.22 MB + K2 nc*A*4 + 2 KB per survivor + K3/K4 ~10 KB + "
                                          "K5 ~61 KB per ROI",
                                  "survey_definition": {"bytes_per_frame": definition,
                                                        "frac": (value / world) / (peak * 1e9 / definition),
                                                        "note": "SURVEY 8(d) counts all (64+nc)*A*4 head bytes; K2 never reads "
                                                                "the DFL channels of non-survivors, so this fraction can exceed 1 "
                                                                "-- it is not a roofline fraction"}},
            "configs": configs,
            "cpu_baseline": cpu, "clocks": clocks, "detections_batch0": n_det,
            "candidates": {"cap": pipe.cap, "max_per_image": max_cand, "fused_postprocess": pipe.fused},
            "numa_node_rank0": numa}
    print(json.dumps(line), flush=True)
    if parity is not None and not (parity["kept_sets_equal"] and parity["class_ids_equal"]):
        print("PARITY FAILURE: kept sets / class ids differ from the oracle", file=sys.stderr)
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reps", type=int, default=25, help="repetitions of the K-step block (median / min are reported)")
    ap.add_argument("--e2e-reps", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip configs 1/3/4 and the 1600x900 e2e case")
    ap.add_argument("--seed", type=int, default=0, help="base seed of the synthetic inputs (rank is added)")
    ap.add_argument("--e2e-mode", default="rows_dfl_zero_copy", choices=["full", "rows", "rows_dfl_zero_copy"],
                    help="which HostRunner staging mode is reported as e2e (all are measured and listed)")
    ap.add_argument("--cap", type=int, default=1024,
                    help="candidate capacity per image (<=1024 selects the fused post-processing kernel; 0 = all anchors)")
    args = ap.parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    from manual_yolo_b200 import multigpu
    rank, world, local = multigpu.init_from_env(os.environ.get("B200_DIST_BACKEND", "nccl"))
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    rc = 0
    try:
        rc = run_ours(args, rank, world, local)
    finally:
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()
    sys.exit(rc or 0)


if __name__ == "__main__":
    main()
