#!/usr/bin/env python
"""bench.py -- post-processed frames/s of the detection hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 16 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # the reference's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the whole path (letterbox -> decode/filter -> sort -> NMS+rescale -> ROI crops)
over one batch of 64 synthetic 1920x1200 BGR frames + the matching (64,128,8400) Detect-head tensor
(BASELINE.json configs[1]; the detector weights are absent from the reference, so the head is the
label-derived synthetic one of manual_yolo_b200/synth.py).  Frames shard across ranks with no
data-path collective (weak scaling: every rank runs its own 64-frame batches).

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same
metric through Pipeline/HostRunner with pinned HOST buffers (H2D of frames + head and D2H of the
detections inside the timed region); `roofline` = the dominant kernel (K1 letterbox) against the
measured HBM peak; `cpu_baseline` = the oracle port timed on this box's host cores.
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
WORKLOAD = ("configs[1]: synthetic 1920x1200 BGR frames, batch=64, 8400 anchors, nc=64 poker head "
            "(label-derived synthetic, poker_model.pt absent), conf=0.25 iou=0.45 max_det=300, "
            "letterbox+decode+sort+NMS+scale_boxes+ROI crops 64x64")


def _config(n_gpus, extra=None):
    c = {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "frame_hw": list(SRC_HW),
         "anchors": 8400, "nc": NC, "conf": CONF, "iou": IOU, "max_det": MAX_DET,
         "parallelism": f"frame-sharded x{n_gpus}, no collective",
         "l2_policy": "inputs larger than L2 (717 MB of frames+head per step vs 126 MB L2); every batch in flight reads its own copy"}
    if extra:
        c.update(extra)
    return c


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
def cpu_path_frames_per_s(frames, head, level_hw, in_hw, n_timed_batches=1, warm=True):
    """Times the oracle (real cv2 / torch CPU / torchvision.ops.nms / PIL leaves + restated UL glue) on the
    host cores over `frames` (numpy (n,H,W,3)) + `head` (torch (n,128,A))."""
    import cv2
    from oracle import boxes as oboxes
    from oracle import head as ohead
    from oracle import letterbox as olb
    from oracle import nms as onms
    from oracle import roi as oroi
    from manual_yolo_b200.pipeline import RANK_CLASS_IDS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cv2.setNumThreads(cores)
    H, W = frames.shape[1:3]

    def one_pass():
        olb.preprocess_ref(list(frames), (IMGSZ, IMGSZ))
        pred = ohead.detect_inference_ref(head, level_hw)
        out = onms.non_max_suppression_ref(pred, CONF, IOU, max_det=MAX_DET)
        n_roi = 0
        for b, o in enumerate(out):
            o = o.clone()
            o[:, :4] = oboxes.scale_boxes_ref(in_hw, o[:, :4], (H, W))
            for row in o:
                if int(row[5]) in RANK_CLASS_IDS:
                    crop = oboxes.safe_crop_ref(frames[b], *[int(v) for v in row[:4]], pad=6)
                    if crop is not None:
                        oroi.classify_preprocess_ref(crop)
                        n_roi += 1
        return n_roi

    if warm:
        one_pass()
    t0 = time.perf_counter()
    for _ in range(n_timed_batches):
        one_pass()
    dt = time.perf_counter() - t0
    return frames.shape[0] * n_timed_batches / dt, cores, dt


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
            "config": _config(args.gpus, {"sample_frames_per_step": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} frames per step (the full batch), {args.steps} steps, oracle port "
                                       "(cv2/torch-CPU/torchvision.nms/PIL leaves), all host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local):
    import manual_yolo_b200 as m
    from manual_yolo_b200 import multigpu, synth
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = multigpu.bind_to_gpu_numa_node(local) if world > 1 else None   # pinned host buffers local to the GPU
    pipe = m.Pipeline(BATCH, SRC_HW, NC, imgsz=IMGSZ, conf=CONF, iou=IOU, max_det=MAX_DET, device=dev,
                      cap=args.cap or None)

    # inputs: seeded per rank (config 5: seeds = base + rank); generated on the CPU so the oracle sees the same bits
    frames_h = synth.synth_frames(BATCH, *SRC_HW, seed=args.seed + rank).pin_memory()
    head_h, _ = synth.synth_head_from_labels(BATCH, NC, in_hw=pipe.in_hw, src_hw=SRC_HW, seed=args.seed + rank,
                                             conf_thres=CONF)
    head_h = head_h.pin_memory()
    frames_d, head_d = frames_h.to(dev), head_h.to(dev)

    # ---- device-resident timing: W warm-up steps, then exactly K steps between barriers ----
    # (a) eager launches with CUDA events around every kernel (per-kernel breakdown + roofline)
    for _ in range(max(3, args.warmup)):
        res = pipe(frames_d, head_d)
    torch.cuda.synchronize()
    pipe.enable_profiling(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    multigpu.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        res = pipe(frames_d, head_d)
    e1.record()
    torch.cuda.synchronize()
    multigpu.barrier()
    ms_eager = multigpu.max_over_ranks(e0.elapsed_time(e1), dev)
    kt_eager = pipe.kernel_times_ms()
    pipe.enable_profiling(False)
    # (a2) the dominant kernel (K1) alone: K back-to-back launches between two events on the launching stream
    #      (its 757 MB of inputs + outputs exceed L2, so consecutive launches do not feed each other)
    for _ in range(3):
        m.preprocess(frames_d, (IMGSZ, IMGSZ), out=pipe.net_in)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        m.preprocess(frames_d, (IMGSZ, IMGSZ), out=pipe.net_in)
    e1.record()
    torch.cuda.synchronize()
    k1_alone_ms = e0.elapsed_time(e1) / args.steps
    # (b) the same step captured once into a CUDA graph and replayed: this is how the path is meant to be
    #     driven (one launch per batch).  First an instrumented graph (external event nodes around every
    #     kernel; one synchronize per replay to read them) for the per-kernel breakdown and the roofline ...
    pipe.enable_profiling(True, external=True)
    pipe.capture(frames_d, head_d)
    kt = {}
    for it in range(max(3, args.warmup) + args.steps):
        pipe.replay()
        torch.cuda.synchronize()
        if it >= max(3, args.warmup):
            for k, v in pipe.kernel_times_ms().items():
                kt.setdefault(k, []).extend(v)
    pipe.enable_profiling(False)
    # serial (single-stream) graph, K replays: reported as value_serial_graph
    pipe.capture(frames_d, head_d)
    for _ in range(max(3, args.warmup)):
        pipe.replay()
    torch.cuda.synchronize()
    multigpu.barrier()
    e0.record()
    for _ in range(args.steps):
        pipe.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_serial = multigpu.max_over_ranks(e0.elapsed_time(e1), dev)
    # ... then the production graph: the letterbox branch (HBM-bound) forked onto a side stream so that it
    # runs concurrently with the decode -> NMS -> ROI branch (latency-bound); K replays back to back between
    # barriers: this is `value`
    pipe.overlap = True
    pipe._make_streams()
    pipe.capture(frames_d, head_d)
    for _ in range(max(3, args.warmup)):
        pipe.replay()
    torch.cuda.synchronize()
    multigpu.barrier()
    e0.record()
    for _ in range(args.steps):
        res = pipe.replay()
    e1.record()
    torch.cuda.synchronize()
    multigpu.barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    ms_max = multigpu.max_over_ranks(ms, dev)
    print(f"[rank {rank}] timed region: {ms:.4f} ms for {args.steps} steps (max over ranks {ms_max:.4f})", file=sys.stderr)
    total_frames = BATCH * args.steps * world
    value_eager = total_frames / (ms_eager / 1e3)
    value_serial = total_frames / (ms_serial / 1e3)

    # ... and the same graph with TWO batches in flight (BatchStream: slot s has its own Pipeline buffers, stream
    # and graph), so the latency-bound tail of batch i runs underneath the bandwidth-bound head of batch i+1
    def timed_stream(depth):
        pipes = [pipe] + [m.Pipeline(BATCH, SRC_HW, NC, imgsz=IMGSZ, conf=CONF, iou=IOU, max_det=MAX_DET, device=dev,
                                     cap=args.cap or None, overlap=True) for _ in range(depth - 1)]
        bs = m.BatchStream(pipes)
        # every slot reads its OWN copy of the inputs (same bits): batches in flight cannot feed each other through L2
        bs.capture([(frames_d, head_d)] + [(frames_d.clone(), head_d.clone()) for _ in range(depth - 1)])
        for _ in range(max(3, args.warmup) * depth):
            bs.submit()
        bs.join()
        torch.cuda.synchronize()
        multigpu.barrier()
        e0.record()
        for _ in range(args.steps):
            r = bs.submit()
        bs.join()
        e1.record()
        torch.cuda.synchronize()
        multigpu.barrier()
        return multigpu.max_over_ranks(e0.elapsed_time(e1), dev), pipes
    ms_one = ms_max
    ms_max, pipes2 = timed_stream(2)
    ms_three, pipes3 = timed_stream(3)
    print(f"[rank {rank}] in flight 1/2/3: {ms_one:.4f} / {ms_max:.4f} / {ms_three:.4f} ms for {args.steps} steps", file=sys.stderr)
    # every slot must have produced the same detections as the single-pipeline run
    for p2 in pipes2[1:] + pipes3[1:]:
        assert torch.equal(p2.ws.det.count, pipe.ws.det.count) and torch.equal(p2.ws.det.rows, pipe.ws.det.rows)
    del pipes2, pipes3
    value_one = total_frames / (ms_one / 1e3)
    value = total_frames / (ms_max / 1e3)
    value_three = total_frames / (ms_three / 1e3)
    t1 = time.time()

    # ---- end to end on host buffers (H2D of the inputs, D2H of the detections, every step) ----
    def timed_e2e(**kw):
        runner = m.HostRunner(pipe, depth=2, **kw)
        for _ in range(3):
            o = runner.submit(frames_h, None if kw.get("head_resident") is not None else head_h)
        torch.cuda.synchronize()
        multigpu.barrier()
        e0.record()
        for _ in range(args.steps):
            o = runner.submit(frames_h, None if kw.get("head_resident") is not None else head_h)
        e1.record()
        torch.cuda.synchronize()
        multigpu.barrier()
        return multigpu.max_over_ranks(e0.elapsed_time(e1), dev), runner, o
    e2e_modes = {}
    for name, kw in (("full", dict(stage="full")), ("rows", dict(stage="rows")),
                     ("rows_dfl_zero_copy", dict(stage="rows", dfl_zero_copy=True))):
        if kw.get("dfl_zero_copy") and not pipe.fused:
            continue
        ms_e, runner, out = timed_e2e(**kw)
        zc = runner.zero_copy_bytes(out[0], out[1])
        if kw.get("dfl_zero_copy"):
            zc += int(pipe.cands.count.clamp(max=pipe.cap).sum()) * 64 * 32      # 64 DFL values, one 32-B sector each
        e2e_modes[name] = {"value": total_frames / (ms_e / 1e3), "ms_per_step": ms_e / args.steps,
                           "h2d_bytes_per_step": runner.h2d_bytes_per_step(), "zero_copy_bytes_per_step": zc}
    e2e_pick = args.e2e_mode if args.e2e_mode in e2e_modes else "rows"  # dense regime (cap > 1024): no zero-copy DFL
    e2e_ms = e2e_modes[e2e_pick]["ms_per_step"] * args.steps
    t2 = time.time()
    clocks = sampler.stop(t0, t2) if rank == 0 else None
    e2e_value = e2e_modes[e2e_pick]["value"]
    # context only: the deployment case, frames from the host but the head already on the device
    e2e2_ms, runner2, _ = timed_e2e(stage="rows", head_resident=head_d)
    n_det = int(out[1].sum())
    max_cand = pipe.check_overflow()           # cap < A drops candidates past cap: the run is valid only if none were

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (K1 letterbox), algorithmic bytes per launch ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    g = pipe.geom
    k1_bytes_frame = g["new_h"] * SRC_HW[1] * 3 + 3 * g["out_h"] * g["out_w"] * 4     # 7 219 200 B
    k1_graph_ms = statistics.mean(kt["letterbox"])            # between external event nodes inside a serialised graph
    k1_ms = k1_alone_ms                                       # back-to-back launches of the kernel alone
    achieved = BATCH * k1_bytes_frame / (k1_ms / 1e3) / 1e9
    achieved_graph = BATCH * k1_bytes_frame / (k1_graph_ms / 1e3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("letterbox_bytes_per_launch")
    except Exception:
        pass
    kernels = {}
    bytes_per_launch = {
        "letterbox": BATCH * k1_bytes_frame,
        "decode_filter": BATCH * (64 + NC) * pipe.A * 4,
    }
    for name, v in kt.items():
        kernels[name] = {"us": 1e3 * statistics.mean(v), "us_eager_launch": 1e3 * statistics.mean(kt_eager[name])}
        if name in bytes_per_launch:
            kernels[name]["algo_GBps"] = bytes_per_launch[name] / (statistics.mean(v) / 1e3) / 1e9
            kernels[name]["frac_of_measured_peak"] = kernels[name]["algo_GBps"] / peak
    pipeline_bytes_frame = k1_bytes_frame + (64 + NC) * pipe.A * 4 + 10_000 + 4 * 60_000   # 11.77 MB
    # ---- CPU baseline (oracle port) on a bounded sample of the same workload ----
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        n, passes = BATCH, 24
        fps, cores, dt = cpu_path_frames_per_s(frames_h[:n].numpy(), head_h[:n].clone(), pipe.level_hw, pipe.in_hw, passes)
        cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"the rank-0 batch of {n} frames, 1 warm-up + {passes} timed passes ({dt:.1f} s), oracle port "
                         "(cv2 / torch CPU / torchvision.ops.nms / PIL), torch+cv2 threads = all cores"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(world, {"launch_mode": "one CUDA-graph replay per step, two batches in flight (BatchStream: one stream + graph + output buffers per slot); inside a graph the letterbox is forked onto a side stream, concurrent with decode->NMS->ROI"}),
            "value_one_in_flight": value_one, "value_three_in_flight": value_three,
            "value_eager_launches": value_eager, "value_serial_graph": value_serial,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_modes[e2e_pick]["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": pipe.d2h_bytes_per_step(), "ms_per_step": e2e_ms / args.steps,
                    "mode": e2e_pick, "zero_copy_bytes_per_step": e2e_modes[e2e_pick]["zero_copy_bytes_per_step"],
                    "note": "HostRunner: pinned host frames+head -> device path -> detections back, double-buffered; "
                            "'rows' stages only the source rows the letterbox reads (one strided DMA) and the ROI kernel "
                            "crops zero-copy from the pinned frames; 'rows_dfl_zero_copy' also stages only the class channels of the head, the "
                            "survivors' DFL values being read zero-copy by the fused post-processing kernel; 'full' copies whole frames + head",
                    "modes": e2e_modes,
                    "frames_only": {"value": total_frames / (e2e2_ms / 1e3), "h2d_bytes_per_step": runner2.h2d_bytes_per_step(),
                                    "note": "context: head resident on the device (as when a backbone produces it), stage='rows'"}},
            "gpu_launches": pipe.launches_per_step() * args.steps * world,   # timed (graph) region only
            "roofline": {"kernel": "letterbox_kernel<float> (K1)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": BATCH * k1_bytes_frame, "avg_launch_us": 1e3 * k1_ms,
                         "timed_in": "K back-to-back launches of the kernel alone between two CUDA events on the launching stream "
                                     "(inputs + outputs 757 MB > L2); ncu isolated launch: 75.0 us (profiles/ncu_full_r01_v3_summary.csv)",
                         "in_graph": {"avg_launch_us": 1e3 * k1_graph_ms, "achieved": achieved_graph, "frac": achieved_graph / peak,
                                      "timed_in": "external event nodes around the kernel inside a single-stream CUDA graph "
                                                  "(includes the event-record nodes' own latency)"}},
            "kernels": kernels,
            "pipeline_roofline": {"bytes_per_frame": pipeline_bytes_frame,
                                  "roofline_frames_per_s_per_gpu": peak * 1e9 / pipeline_bytes_frame,
                                  "frac": (value / world) / (peak * 1e9 / pipeline_bytes_frame)},
            "cpu_baseline": cpu, "clocks": clocks, "detections_last_step": n_det,
            "candidates": {"cap": pipe.cap, "max_per_image": max_cand, "fused_postprocess": pipe.fused},
            "numa_node_rank0": numa}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    try:
        run_ours(args, rank, world, local)
    finally:
        if torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
