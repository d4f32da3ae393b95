#!/usr/bin/env python
"""bench_extra.py -- the BASELINE.json configs that are not the headline line (parity cases with timings):

  config 3  NMS-heavy eval regime: (256,144,8400) head, nc=80, conf=0.001, max_det=300, iou 0.7 / 0.45
            -> decode us, sort us, NMS us per batch and per image
  config 4  ROI chain: 4096 rank-box crops -> (4096,3,64,64) fp32
  config 1  B=1 test2.png-shaped frame (1600x900), both letterbox modes, nc=64, conf 0.25

Prints one JSON object (also written to --out).  CUDA events on the launching stream, 3 warm-ups,
inputs re-used (these are latency numbers; the L2 state is stated per entry).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import manual_yolo_b200 as m  # noqa: E402
from manual_yolo_b200 import geometry, synth  # noqa: E402


def timed(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return {"us_median": statistics.median(ts), "us_min": min(ts), "iters": iters}


def config3(dev, B=256, nc=80, cpu_images=4):
    lv = geometry.level_shapes(640, 640)
    parts = [synth.synth_head_dense(64, nc, seed=s) for s in range(B // 64)]
    head = torch.cat(parts).to(dev)
    out = {"shape": list(head.shape), "conf": 0.001, "max_det": 300,
           "l2": "head 1.24 GB streams from HBM; candidate arrays (57 MB) are L2-resident"}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    cands = m.decode_and_filter(head, conf_thres=0.001, level_hw=lv)
    ws = m.Workspace(B, cands.cap, 300, dev)
    out["candidates_per_image_mean"] = float(cands.count.float().mean())

    def dec():
        cands.count.zero_()
        m.decode_and_filter(head, conf_thres=0.001, level_hw=lv, out=cands)
    out["decode_filter"] = timed(dec, flush=flush)
    out["decode_filter"]["algo_GBps"] = head.numel() * 4 / (out["decode_filter"]["us_median"] * 1e-6) / 1e9
    out["sort_topk"] = timed(lambda: m.sort_candidates(cands, 30000, ws))
    for iou in (0.7, 0.45):
        r = timed(lambda: m.nms_sorted(cands, ws, iou, max_det=300))
        r["us_per_image"] = r["us_median"] / B
        out[f"nms_iou{iou}"] = r
    out["kept_mean"] = float(ws.det.count.float().mean())
    # the production chain for this regime: class filter (scores only) + b200yolo_postprocess_dense
    # (select-sort of the best 2048 -> DFL decode of exactly those -> windowed NMS, exact fallback)
    cands2 = m.decode_and_filter(head, conf_thres=0.001, level_hw=lv, defer_boxes=True)
    ws2 = m.Workspace(B, cands2.cap, 300, dev)

    def cf():
        cands2.count.zero_()
        m.decode_and_filter(head, conf_thres=0.001, level_hw=lv, out=cands2, defer_boxes=True)
    out["class_filter"] = timed(cf, flush=flush)
    out["class_filter"]["algo_GBps"] = head.numel() * 4 / (out["class_filter"]["us_median"] * 1e-6) / 1e9
    out["postprocess_dense_iou0.7"] = timed(lambda: m.postprocess_dense(cands2, ws2, head, level_hw=lv, iou_thres=0.7))

    def chain():
        cf()
        m.postprocess_dense(cands2, ws2, head, level_hw=lv, iou_thres=0.7)
    r = timed(chain, flush=flush)
    r["us_per_image"] = r["us_median"] / B
    r["algo_GBps"] = head.numel() * 4 / (r["us_median"] * 1e-6) / 1e9
    out["dense_chain_total"] = r
    m.nms_sorted(cands, ws, 0.7, max_det=300)
    out["dense_chain_equals_stagewise"] = bool(torch.equal(ws2.det.count, ws.det.count) and
                                               torch.equal(ws2.det.anchor[:, :1], ws.det.anchor[:, :1]))
    # the same chain on sub-batches of the images, each on its own stream (api.DenseChain): the latency-bound
    # per-image kernels (select-sort, NMS) of one sub-batch run underneath the HBM-bound ones of another
    def graphed(fn):
        """fn captured once into a CUDA graph (how a deployment drives the chain: one launch per batch)."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g
    best = None
    g1 = graphed(chain)
    r = timed(lambda: g1.replay(), flush=flush)
    r["us_per_image"] = r["us_median"] / B
    r["algo_GBps"] = head.numel() * 4 / (r["us_median"] * 1e-6) / 1e9
    out["dense_chain_total_cuda_graph"] = r
    del g1
    # forked: every sub-batch's whole chain on its own stream; pipelined: class filters back to back on one stream, each
    # sub-batch's select-sort -> decode -> NMS forked onto a high-priority stream as soon as its own filter is done
    for splits, pipelined in ((2, True), (4, False), (8, False)):
        dc = m.DenseChain(B, cands2.cap, 300, dev, splits=splits, pipelined=pipelined)
        gd = graphed(lambda: dc(head, conf_thres=0.001, iou_thres=0.7, level_hw=lv))
        r = timed(lambda: gd.replay(), flush=flush)
        r["us_per_image"] = r["us_median"] / B
        r["algo_GBps"] = head.numel() * 4 / (r["us_median"] * 1e-6) / 1e9
        r["equals_single_stream"] = bool(torch.equal(dc.det.count, ws2.det.count) and torch.equal(dc.det.anchor, ws2.det.anchor)
                                         and torch.equal(dc.det.rows, ws2.det.rows))
        out[f"dense_chain_{'pipelined' if pipelined else 'forked'}_{splits}_cuda_graph"] = r
        if best is None or r["us_median"] < best[1]:
            best = (splits, r["us_median"], "pipelined" if pipelined else "forked")
        del gd, dc
    out["dense_chain_best"] = {"splits": best[0], "us_median": best[1], "us_per_image": best[1] / B,
                               "algo_GBps": head.numel() * 4 / (best[1] * 1e-6) / 1e9,
                               "what": f"api.DenseChain(splits={best[0]}, {best[2]}) as one CUDA graph"}
    if not cpu_images:
        return out
    # CPU oracle on a sub-sample, scaled (flagged)
    from oracle import head as ohead
    from oracle import nms as onms
    h = head[:cpu_images].cpu()
    t0 = time.perf_counter()
    pred = ohead.detect_inference_ref(h, lv)
    t1 = time.perf_counter()
    onms.non_max_suppression_ref(pred, 0.001, 0.7, max_det=300)
    t2 = time.perf_counter()
    out["cpu_oracle"] = {"images": cpu_images, "decode_ms_per_image": 1e3 * (t1 - t0) / cpu_images,
                         "nms_ms_per_image": 1e3 * (t2 - t1) / cpu_images, "cores": os.cpu_count(),
                         "note": f"sub-sample of {cpu_images} images, per-image figures (not scaled to B)"}
    return out


def config4(dev, N=4096, B=64):
    frames = synth.synth_frames(B, 1200, 1920, seed=0).to(dev)
    boxes, bidx = synth.synth_rois(N, B, seed=0)
    boxes, bidx = boxes.to(dev), bidx.to(dev)
    dst = torch.empty((N, 3, 64, 64), dtype=torch.float32, device=dev)
    valid = torch.empty((N,), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    r = timed(lambda: m.crop_resize_rois(frames, boxes, bidx, pad=6, out=dst, valid=valid), flush=flush)
    from oracle import boxes as oboxes
    crop_bytes = 0
    for i in range(N):
        c = oboxes.safe_crop_box_ref((1200, 1920), *[int(v) for v in boxes[i].cpu()], pad=6)
        crop_bytes += (c[2] - c[0]) * (c[3] - c[1]) * 3
    algo = crop_bytes + N * 49152
    r.update(rois=N, rois_per_s=N / (r["us_median"] * 1e-6), algorithmic_bytes=algo,
             algo_GBps=algo / (r["us_median"] * 1e-6) / 1e9, l2="L2 flushed between iterations")
    return r


def config_n3(dev, F=16):
    """SAHI-style sliced prediction of 1920x1200 frames (12 slices of 640x640 per frame)."""
    sp = m.SlicedPipeline(F, (1200, 1920), 64, conf=0.25, iou=0.7, merge_iou=0.5, device=dev, cap=1024, rois_per_frame=64)
    frames = synth.synth_frames(F, 1200, 1920, seed=0).to(dev)
    base, _ = synth.synth_head_from_labels(sp.S, 64, in_hw=sp.in_hw, src_hw=sp.slice_hw, seed=0, conf_thres=0.25)
    head = base.repeat(F, 1, 1).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    k1 = timed(lambda: sp.preprocess(frames), flush=flush)
    out_bytes = F * sp.S * 3 * sp.in_hw[0] * sp.in_hw[1] * 4
    in_bytes = F * sp.S * sp.slice_hw[0] * sp.slice_hw[1] * 3
    k1.update(algorithmic_bytes=in_bytes + out_bytes, algo_GBps=(in_bytes + out_bytes) / (k1["us_median"] * 1e-6) / 1e9)
    step = timed(lambda: sp(frames, head), flush=flush)
    res = sp(frames, head)
    torch.cuda.synchronize()
    out = {"frames": F, "slices_per_frame": sp.S, "letterbox_slices": k1, "sliced_step": step,
           "frames_per_s": F / (step["us_median"] * 1e-6), "slices_per_s": F * sp.S / (step["us_median"] * 1e-6),
           "merged_detections_per_frame": float(res.det.count.float().mean()),
           "slice_detections_per_frame": float(sp.slice_det.count.float().sum() / F),
           "l2": "L2 flushed between iterations; eager launches"}
    # the form the reference's call has (SAHI defaults): + full-frame standard prediction, GREEDYNMM / IOS 0.5 merge
    del sp
    sd = m.SlicedPipeline(F, (1200, 1920), 64, conf=0.25, iou=0.7, merge_iou=0.5, device=dev, cap=1024, rois_per_frame=64,
                          merge="greedy_nmm", match_metric="IOS", standard_pred=True)
    hf, _ = synth.synth_head_from_labels(F, 64, in_hw=sd.full.in_hw, src_hw=(1200, 1920), seed=3, conf_thres=0.25)
    hf = hf.to(dev)
    step2 = timed(lambda: sd(frames, head, head_full=hf), flush=flush)
    res2 = sd(frames, head, head_full=hf)
    torch.cuda.synchronize()
    try:                                                   # the same step as one CUDA graph (how a service would drive it)
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            sd(frames, head, head_full=hf)
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            sd(frames, head, head_full=hf)
        step2g = timed(lambda: g.replay(), flush=flush)
        step2g["frames_per_s"] = F / (step2g["us_median"] * 1e-6)
    except Exception as e:                                 # noqa: BLE001 -- reported, not hidden
        step2g = {"error": repr(e)[:200]}
    out["sahi_defaults"] = {"sliced_step": step2, "sliced_step_cuda_graph": step2g, "frames_per_s": F / (step2["us_median"] * 1e-6),
                            "merged_detections_per_frame": float(res2.det.count.float().mean()),
                            "what": "12 slices + the full frame per frame, GREEDYNMM (IOS 0.5) merge, ROI crops"}
    return out


def config_n2(dev, n_frames=200, n_obj=30):
    """ByteTrack (device Kalman + IoU costs, host association) on a synthetic sequence of padded NMS
    outputs: tracker updates per second (host-bound: a few small launches and two small reads per frame)."""
    import numpy as np
    from manual_yolo_b200 import tracking
    rng = np.random.default_rng(0)
    pos = rng.uniform(100, 1700, (n_obj, 2)); vel = rng.uniform(-5, 5, (n_obj, 2)); size = rng.uniform(40, 120, (n_obj, 2))
    max_det = 300
    dets = []
    for f in range(n_frames):
        c = pos + vel * f + rng.normal(0, 1.0, pos.shape)
        keep = rng.random(n_obj) > 0.1
        rows = np.concatenate([c - size / 2, c + size / 2, rng.uniform(0.3, 0.95, (n_obj, 1)), rng.integers(0, 5, (n_obj, 1))], 1)[keep]
        pad = torch.zeros((1, max_det, 6)); pad[0, :rows.shape[0]] = torch.from_numpy(rows.astype(np.float32))
        dets.append(m.Detections(pad.to(dev), torch.zeros((1, max_det), dtype=torch.int32, device=dev),
                                 torch.tensor([rows.shape[0]], dtype=torch.int32, device=dev)))
    trk = tracking.ByteTrack(device=dev, capacity=256, max_det=max_det)
    for d in dets[:20]:
        trk.update(d, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids = 0
    for d in dets[20:]:
        ids += int((trk.update(d, 0) >= 0).sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"frames": n_frames - 20, "objects": n_obj, "updates_per_s": (n_frames - 20) / dt, "ms_per_update": 1e3 * dt / (n_frames - 20),
            "tracked_detections": ids, "live_tracks": len(trk.tracked), "what": "wall clock, one frame per update, host-bound"}


def config1(dev):
    """BASELINE configs[0]: one 1600x900 frame (test2.png shape), imgsz 640, conf 0.25, iou 0.45.  cap=1024 is the
    fused sparse-regime path (what this confidence threshold calls for); cap=None the general path."""
    out = {}
    frame = synth.synth_frames(1, 900, 1600, seed=0)
    for auto in (False, True):
        for cap in (1024, None):
            pipe = m.Pipeline(1, (900, 1600), 64, imgsz=640, auto=auto, conf=0.25, iou=0.45, device=dev, cap=cap)
            head, _ = synth.synth_head_from_labels(1, 64, in_hw=pipe.in_hw, src_hw=(900, 1600), seed=0)
            f, h = frame.to(dev), head.to(dev)
            eager = timed(lambda: pipe(f, h), iters=20)
            pipe.capture(f, h)
            graph = timed(lambda: pipe.replay(), iters=20)
            key = f"auto={auto}" + ("" if cap else ",general_path")
            out[key] = {"in_hw": list(pipe.in_hw), "anchors": pipe.A, "cap": pipe.cap, "fused": pipe.fused, "eager": eager,
                        "cuda_graph": graph, "detections": int(pipe.ws.det.count[0])}
    return out


def k1_half(dev, B=64):
    """K1 with fp16 output (predict(half=True) form): half the bytes written."""
    frames = synth.synth_frames(B, 1200, 1920, seed=0).to(dev)
    out16 = torch.empty((B, 3, 640, 640), dtype=torch.float16, device=dev)
    out32 = torch.empty((B, 3, 640, 640), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    res = {}
    for name, o, half, nbytes in (("f32", out32, False, 4), ("f16", out16, True, 2)):
        r = timed(lambda: m.preprocess(frames, (640, 640), out=o, half=half), flush=flush)
        algo = B * (400 * 1920 * 3 + 3 * 640 * 640 * nbytes)
        r.update(algorithmic_bytes=algo, algo_GBps=algo / (r["us_median"] * 1e-6) / 1e9)
        res[name] = r
    return res


def calibration(dev):
    """What plain torch copy / fill kernels reach on this GPU for K1's traffic mix (147 MB read, 315 MB written)."""
    out = {}
    src = torch.empty(147_456_000, dtype=torch.uint8, device=dev)
    dst = torch.empty(78_643_200, dtype=torch.float32, device=dev)
    big_a = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device=dev)
    big_b = torch.empty_like(big_a)
    r = timed(lambda: dst.fill_(0.5), iters=10)
    out["fill_315MB"] = {**r, "GBps": dst.numel() * 4 / (r["us_median"] * 1e-6) / 1e9}
    r = timed(lambda: big_b.copy_(big_a), iters=10)
    out["copy_1GiB"] = {**r, "GBps": 2 * big_a.numel() * 4 / (r["us_median"] * 1e-6) / 1e9}
    r = timed(lambda: src.sum(), iters=10)
    out["read_reduce_147MB_u8"] = {**r, "GBps": src.numel() / (r["us_median"] * 1e-6) / 1e9}
    r = timed(lambda: big_a.sum(), iters=10)
    out["read_reduce_1GiB_f32"] = {**r, "GBps": big_a.numel() * 4 / (r["us_median"] * 1e-6) / 1e9}
    return out


def cpu_threads(dev, n=16):
    """The CPU oracle on the headline workload with all host threads and with one thread."""
    import cv2
    from bench import BATCH, CONF, IMGSZ, NC, SRC_HW, cpu_path_frames_per_s
    frames = synth.synth_frames(n, *SRC_HW, seed=0).numpy()
    lv = geometry.level_shapes(IMGSZ, IMGSZ)
    head, _ = synth.synth_head_from_labels(n, NC, in_hw=(IMGSZ, IMGSZ), src_hw=SRC_HW, seed=0, conf_thres=CONF)
    out = {}
    fps, cores, dt = cpu_path_frames_per_s(frames, head, lv, (IMGSZ, IMGSZ), 2)
    out["all_threads"] = {"frames_per_s": fps, "cores": cores, "seconds": dt}
    torch.set_num_threads(1)
    cv2.setNumThreads(1)
    t0 = time.perf_counter()
    from oracle import head as ohead, letterbox as olb, nms as onms
    for _ in range(2):
        olb.preprocess_ref(list(frames), (IMGSZ, IMGSZ))
        onms.non_max_suppression_ref(ohead.detect_inference_ref(head, lv), CONF, 0.45)
    out["one_thread"] = {"frames_per_s": 2 * n / (time.perf_counter() - t0), "cores": 1,
                         "note": "letterbox + decode + NMS only (ROI stage excluded)"}
    torch.set_num_threads(os.cpu_count())
    cv2.setNumThreads(os.cpu_count())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_extra.json"))
    ap.add_argument("--b3", type=int, default=256)
    ap.add_argument("--only", default="", help="comma list of: calibration,config3,config4,config1,n3,n2,k1half,cpu (default: all)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    parts = {"calibration": ("calibration", lambda: calibration(dev)),
             "config3": ("config3_nms_heavy", lambda: config3(dev, B=args.b3)),
             "config4": ("config4_roi_4096", lambda: config4(dev)),
             "config1": ("config1_single_frame", lambda: config1(dev)),
             "n3": ("n3_sliced_prediction", lambda: config_n3(dev)),
             "n2": ("n2_bytetrack", lambda: config_n2(dev)),
             "k1half": ("k1_letterbox_f32_vs_f16", lambda: k1_half(dev)),
             "cpu": ("cpu_oracle_threads", lambda: cpu_threads(dev))}
    only = [x for x in args.only.split(",") if x] or list(parts)
    res = {parts[k][0]: parts[k][1]() for k in only}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
