"""Whole path (rows a1-a13 chained) at production shapes vs the oracle chain; CUDA-graph replay; e2e host entry."""
import os

import cv2
import numpy as np
import pytest
import torch

import manual_yolo_b200 as m
from manual_yolo_b200 import synth
from oracle import boxes as oboxes
from oracle import head as ohead
from oracle import letterbox as olb
from oracle import nms as onms
from oracle import roi as oroi

pytestmark = pytest.mark.gpu


def _oracle_chain(frames, head, pipe, conf, iou):
    H, W = frames.shape[1:3]
    net_in = olb.preprocess_ref(list(frames), pipe.new_shape, auto=pipe.auto)
    pred = ohead.detect_inference_ref(head, pipe.level_hw)
    out, idx = onms.non_max_suppression_ref(pred, conf, iou, return_idxs=True)
    dets, rois, where = [], [], []
    for b, o in enumerate(out):
        o = o.clone()
        o[:, :4] = oboxes.scale_boxes_ref(pipe.in_hw, o[:, :4], (H, W))
        dets.append(o)
        for i, row in enumerate(o):
            if int(row[5]) in m.pipeline.RANK_CLASS_IDS:
                crop = oboxes.safe_crop_ref(frames[b], *[int(v) for v in row[:4]], pad=6)
                rois.append(None if crop is None else oroi.classify_preprocess_ref(crop))
                where.append((b, i))
    return net_in, dets, idx, rois, where


def _assert_matches(res, ref, B):
    net_in, dets, idx, rois, where = ref
    assert torch.equal(res.net_in.cpu(), net_in)
    counts = res.det.count.cpu().tolist()
    for b in range(B):
        assert counts[b] == dets[b].shape[0]
        assert torch.equal(res.det.anchor[b, :counts[b]].cpu().long(), idx[b])          # kept sets bit-exact
        got = res.det.rows[b, :counts[b]].cpu()
        assert torch.equal(got[:, 5], dets[b][:, 5])                                     # class ids bit-exact
        if counts[b]:
            assert (got[:, :5] - dets[b][:, :5]).abs().max().item() <= 1e-4
    n = int(res.roi_count.cpu())
    assert n == len(rois)
    assert list(zip(res.roi_batch[:n].cpu().tolist(), res.roi_det[:n].cpu().tolist())) == where
    for i, r in enumerate(rois):
        if r is None:
            assert int(res.roi_valid[i]) == 0
        else:
            assert (res.rois[i].cpu() - r).abs().max().item() <= 1 / 255
    return sum(counts), n


@pytest.mark.parametrize("cap", [None, 1024])          # general kernels / fused sparse-regime kernel
@pytest.mark.parametrize("src_hw,auto,conf,iou", [((1200, 1920), False, 0.25, 0.45), ((900, 1600), True, 0.25, 0.7),
                                                  ((1200, 1920), False, 0.25, 0.7)])
def test_pipeline_matches_oracle_chain(cuda_dev, golden_dir, src_hw, auto, conf, iou, cap):
    B, nc = 4, 64
    frames = synth.synth_frames(B, *src_hw, seed=11).numpy().copy()
    real = cv2.imread(os.path.join(golden_dir, "frames", f"frame_{src_hw[1]}x{src_hw[0]}.jpg"))
    frames[0] = real                                                   # one real dataset frame per batch
    pipe = m.Pipeline(B, src_hw, nc, imgsz=640, auto=auto, conf=conf, iou=iou, device=cuda_dev, cap=cap)
    assert pipe.fused == (cap is not None)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe.in_hw, src_hw=src_hw, seed=11, conf_thres=conf)
    res = pipe(torch.from_numpy(frames).to(cuda_dev), head.to(cuda_dev))
    ndet, nroi = _assert_matches(res, _oracle_chain(frames, head, pipe, conf, iou), B)
    assert ndet > 40 and nroi > 4
    assert pipe.check_overflow() <= pipe.cap


def test_graph_replay_and_host_entry(cuda_dev):
    B, nc, src_hw = 2, 64, (600, 960)
    pipe = m.Pipeline(B, src_hw, nc, imgsz=320, conf=0.25, iou=0.7, device=cuda_dev)
    frames = synth.synth_frames(B, *src_hw, seed=5)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe.in_hw, src_hw=src_hw, seed=5)
    ref = _oracle_chain(frames.numpy(), head, pipe, 0.25, 0.7)
    d_frames, d_head = torch.zeros_like(frames, device=cuda_dev), torch.zeros_like(head, device=cuda_dev)
    pipe.capture(d_frames, d_head)
    d_frames.copy_(frames); d_head.copy_(head)
    res = pipe.replay()
    torch.cuda.synchronize()
    _assert_matches(res, ref, B)
    # overlap mode: letterbox on a side stream (fork/join), eager and captured -- same results
    pipe2 = m.Pipeline(B, src_hw, nc, imgsz=320, conf=0.25, iou=0.7, device=cuda_dev, cap=1024, overlap=True)
    _assert_matches(pipe2(d_frames, d_head), ref, B)
    pipe2.capture(d_frames, d_head)
    pipe2.net_in.zero_()
    res2 = pipe2.replay()
    torch.cuda.synchronize()
    _assert_matches(res2, ref, B)
    # host-facing entry: pinned host buffers in, host results out
    rows, count, nroi = pipe.run_host(frames.pin_memory(), head.pin_memory())
    torch.cuda.synchronize()
    for b in range(B):
        assert int(count[b]) == ref[1][b].shape[0]
        assert (rows[b, :int(count[b]), :5] - ref[1][b][:, :5]).abs().max().item() <= 1e-4
    assert int(nroi) == len(ref[3])
    recs = m.pipeline.detections_to_records(rows, count)
    assert len(recs) == int(count.sum()) and all(len(r["bbox"]) == 4 for r in recs)


def test_fused_postprocess_equals_general_kernels(cuda_dev):
    """postprocess_small (one launch) must be bit-identical to sort_topk + nms, incl. the radix path (n > 512),
    ties, agnostic mode, max_det cut and the per-level head form; overflow past cap is detectable."""
    from oracle import head as ohead
    g = torch.Generator().manual_seed(7)
    A = 8400
    for n_obj, kw in [(40, {}), (511, {}), (513, {}), (1000, {}), (1024, dict(agnostic=True)), (900, dict(max_det=25))]:
        pred = torch.zeros((2, 4 + 16, A))
        for b in range(2):
            pred[b, 0] = torch.rand(A, generator=g) * 600
            pred[b, 1] = torch.rand(A, generator=g) * 600
            pred[b, 2] = torch.rand(A, generator=g) * 90 + 5
            pred[b, 3] = torch.rand(A, generator=g) * 90 + 5
            sel = torch.randperm(A, generator=g)[:n_obj]
            sc = torch.rand(n_obj, generator=g) * 0.7 + 0.3
            sc[::5] = sc[0]                                   # score ties -> anchor order decides
            pred[b, 4 + torch.randint(0, 16, (n_obj,), generator=g), sel] = sc
        pd = pred.to(cuda_dev)
        max_det = kw.get("max_det", 300)
        ref = m.nms_candidates(m.filter_decoded(pd, 0.25), 0.45, agnostic=kw.get("agnostic", False), max_det=max_det)
        ref_rows, ref_anchor, ref_count = ref.rows.clone(), ref.anchor.clone(), ref.count.clone()
        cands = m.filter_decoded(pd, 0.25, cap=1024)
        ws = m.Workspace(2, 1024, max_det, cuda_dev)
        det = m.postprocess_small(cands, ws.det, None, iou_thres=0.45, agnostic=kw.get("agnostic", False))
        assert torch.equal(det.count, ref_count)
        for b in range(2):
            k = int(ref_count[b])
            assert torch.equal(det.anchor[b, :k], ref_anchor[b, :k])
            assert torch.equal(det.rows[b, :k], ref_rows[b, :k])
    # raw head, per-level tensors, deferred boxes
    head, _ = synth.synth_head_from_labels(3, 64, seed=21)
    lv = m.geometry.level_shapes(640, 640)
    levels, off = [], 0
    for h, w in lv:
        levels.append(head[:, :, off:off + h * w].reshape(3, 128, h, w).contiguous().to(cuda_dev))
        off += h * w
    ref = m.nms_candidates(m.decode_and_filter(levels, conf_thres=0.25), 0.7)
    ref_rows, ref_count = ref.rows.clone(), ref.count.clone()
    cands = m.decode_and_filter(levels, conf_thres=0.25, cap=512, defer_boxes=True)
    ws = m.Workspace(3, 512, 300, cuda_dev)
    det = m.postprocess_small(cands, ws.det, levels, iou_thres=0.7)
    assert torch.equal(det.count, ref_count)
    for b in range(3):
        assert torch.equal(det.rows[b, :int(ref_count[b])], ref_rows[b, :int(ref_count[b])])
    # overflow: a cap below the candidate count must be reported, not silently accepted
    pipe = m.Pipeline(3, (600, 960), 64, imgsz=640, conf=0.25, device=cuda_dev, cap=8)
    pipe(synth.synth_frames(3, 600, 960, seed=1).to(cuda_dev), head.to(cuda_dev))
    with pytest.raises(RuntimeError):
        pipe.check_overflow()


def test_host_runner_stage_modes_identical(cuda_dev):
    """HostRunner: 'full' staging, referenced-rows staging + zero-copy ROI crops, and class-channels-only head
    staging + zero-copy DFL reads must give the oracle's results and bit-identical outputs to one another."""
    B, nc, src_hw = 3, 64, (1200, 1920)                    # 1200 -> 400 rows: odd scale 3, one row in three staged
    conf, iou = 0.25, 0.45
    frames = synth.synth_frames(B, *src_hw, seed=31)
    pipe = m.Pipeline(B, src_hw, nc, imgsz=640, conf=conf, iou=iou, device=cuda_dev, cap=1024)
    assert pipe.rows_plan == (1, 3, 400)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe.in_hw, src_hw=src_hw, seed=31, conf_thres=conf)
    ref = _oracle_chain(frames.numpy(), head, pipe, conf, iou)
    fh, hh = frames.pin_memory(), head.pin_memory()
    outs = []
    for kw in (dict(stage="full"), dict(stage="rows"), dict(stage="rows", dfl_zero_copy=True)):
        runner = m.HostRunner(pipe, depth=2, **kw)
        for _ in range(3):                                   # exercises both staging slots and the free/ready events
            rows, count, nroi = runner.submit(fh, hh)
        runner.wait()
        torch.cuda.synchronize()
        res = m.PipelineResult(pipe.net_in, pipe.ws.det, pipe.cands.count, *pipe.roi_out)
        _assert_matches(res, ref, B)
        outs.append((rows.clone(), count.clone(), int(nroi), pipe.net_in.clone(), pipe.roi_out[0].clone()))
        assert runner.h2d_bytes_per_step() > 0
    full = B * 1200 * 1920 * 3 + B * 128 * 8400 * 4
    assert m.HostRunner(pipe, stage="rows").h2d_bytes_per_step() == B * 400 * 1920 * 3 + B * 128 * 8400 * 4 < full
    for o in outs[1:]:
        assert torch.equal(o[1], outs[0][1]) and o[2] == outs[0][2]
        for b in range(B):
            assert torch.equal(o[0][b, :int(o[1][b])], outs[0][0][b, :int(o[1][b])])
        assert torch.equal(o[3], outs[0][3]) and torch.equal(o[4][:o[2]], outs[0][4][:o[2]])
    # a scale that is not an odd integer stages the whole frame (plan = identity) and still works
    pipe2 = m.Pipeline(2, (900, 1600), nc, imgsz=640, conf=conf, iou=iou, device=cuda_dev, cap=1024)
    assert pipe2.rows_plan == (0, 1, 900)
    f2 = synth.synth_frames(2, 900, 1600, seed=32)
    h2, _ = synth.synth_head_from_labels(2, nc, in_hw=pipe2.in_hw, src_hw=(900, 1600), seed=32, conf_thres=conf)
    r2 = m.HostRunner(pipe2, stage="rows")
    r2.submit(f2.pin_memory(), h2.pin_memory())
    torch.cuda.synchronize()
    _assert_matches(m.PipelineResult(pipe2.net_in, pipe2.ws.det, pipe2.cands.count, *pipe2.roi_out),
                    _oracle_chain(f2.numpy(), h2, pipe2, conf, iou), 2)
    # unpinned host memory is refused (no silent synchronous copy, no CPU path)
    with pytest.raises(ValueError):
        m.stage_rows_h2d(frames)
    with pytest.raises(ValueError):
        m.rois_from_detections(frames, pipe.ws.det, pipe.roi_cnt, pipe.roi_mask, nc, pipe.roi_cap)


def test_batch_stream_two_in_flight(cuda_dev):
    """BatchStream: two batches in flight on their own streams/graphs give the same results as one Pipeline."""
    B, nc, src_hw = 2, 64, (600, 960)
    data = []
    for seed in (41, 42):
        f = synth.synth_frames(B, *src_hw, seed=seed)
        pipe0 = m.Pipeline(B, src_hw, nc, imgsz=320, conf=0.25, iou=0.7, device=cuda_dev, cap=1024)
        h, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe0.in_hw, src_hw=src_hw, seed=seed)
        data.append((f, h, _oracle_chain(f.numpy(), h, pipe0, 0.25, 0.7)))
    pipes = [m.Pipeline(B, src_hw, nc, imgsz=320, conf=0.25, iou=0.7, device=cuda_dev, cap=1024, overlap=True)
             for _ in range(2)]
    bs = m.BatchStream(pipes)
    static = [(torch.zeros_like(f, device=cuda_dev), torch.zeros_like(h, device=cuda_dev)) for f, h, _ in data]
    bs.capture(static)
    for (df, dh), (f, h, _) in zip(static, data):
        df.copy_(f); dh.copy_(h)
    for p in pipes:
        p.net_in.zero_()
    torch.cuda.synchronize()
    results = [bs.submit() for _ in range(6)]               # slots alternate: 0,1,0,1,...
    bs.join()
    torch.cuda.synchronize()
    for s in range(2):
        _assert_matches(results[4 + s], data[s][2], B)


def test_plain_c_host_runs_the_path(cuda_dev, tmp_path):
    """The C ABI is a boundary for non-Python hosts too: examples/c_host/main.c (gcc, cudaMalloc, no torch) runs
    letterbox -> class filter -> fused post-processing -> ROI crops and must find its three planted objects."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_host")
    lib_dir = os.path.join(root, "manual_yolo_b200")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(root, "examples", "c_host", "main.c"), "-L", lib_dir, "-lb200yolo", "-L", "/usr/local/cuda/lib64",
                    "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
    assert "3 detections, 3 ROIs" in r.stdout


def test_capacity_overflow_raises_on_every_host_path(cuda_dev):
    """'raises rather than truncating silently'.  One image of a batch with more candidates than cap: the
    fused kernel reports the unclamped count (and re-arms the counters itself: no memset between steps), and every host
    entry that reads results back -- run_host, HostRunner, BatchStream.check_overflow, PipelineResult -- raises."""
    B, nc, src_hw = 3, 64, (600, 960)
    frames = synth.synth_frames(B, *src_hw, seed=2)
    pipe = m.Pipeline(B, src_hw, nc, imgsz=640, conf=0.25, device=cuda_dev, cap=256, rois_per_frame=64)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe.in_hw, src_hw=src_hw, seed=2)
    res = pipe(frames.to(cuda_dev), head.to(cuda_dev))
    seen = res.cand_count.cpu().tolist()
    assert max(seen) <= 256 and res.check_overflow() == max(seen)          # this batch fits
    assert int(pipe.cands.count.abs().sum()) == 0                            # counters re-armed inside the kernel
    res2 = pipe(frames.to(cuda_dev), head.to(cuda_dev))                      # second step without any memset: same counts
    assert res2.cand_count.cpu().tolist() == seen
    # image 1 gets 400 extra high-score anchors: more candidates than cap
    hot = head.clone()
    hot[1, 64 + 5, 100:500] = 4.0
    res3 = pipe(frames.to(cuda_dev), hot.to(cuda_dev))
    seen3 = res3.cand_count.cpu().tolist()
    assert seen3[1] >= 400 > 256 and seen3[0] == seen[0] and seen3[2] == seen[2]
    with pytest.raises(m.pipeline.CandidateOverflow):
        res3.check_overflow()
    with pytest.raises(m.pipeline.CandidateOverflow):
        pipe.check_overflow()
    with pytest.raises(m.pipeline.CandidateOverflow):
        pipe.run_host(frames.pin_memory(), hot.pin_memory())
    runner = m.HostRunner(pipe, depth=2, stage="full")
    runner.submit(frames.pin_memory(), hot.pin_memory())
    with pytest.raises(m.pipeline.CandidateOverflow):
        runner.wait()
    # the steps after the overflowing one are clean again (the kernel re-armed the counters)
    runner2 = m.HostRunner(pipe, depth=2, stage="full")
    rows, count, nroi = runner2.submit(frames.pin_memory(), head.pin_memory())
    runner2.wait()
    assert int(count.sum()) == int(res.det.count.sum())
    # ROI capacity: fewer crop slots than rank-class detections -> RoiOverflow, count not clamped
    small = m.Pipeline(B, src_hw, nc, imgsz=640, conf=0.25, device=cuda_dev, cap=1024, rois_per_frame=1)
    r = small(frames.to(cuda_dev), head.to(cuda_dev))
    assert int(r.roi_count) > small.roi_cap == 3 and r.n_rois() == 3
    with pytest.raises(m.pipeline.RoiOverflow):
        r.check_overflow()
    bs = m.BatchStream([small])
    bs.capture([(frames.to(cuda_dev), head.to(cuda_dev))])
    bs.submit()
    with pytest.raises(m.pipeline.RoiOverflow):
        bs.check_overflow()


def test_empty_inputs_scale_boxes_and_nms(cuda_dev):
    """an empty CUDA tensor has a NULL data pointer -- scale_boxes on (0,4) and non_max_suppression on a
    prediction without candidates must work (upstream construct_result calls scale_boxes on every frame)."""
    empty = torch.zeros((0, 6), device=cuda_dev)
    out = m.scale_boxes((640, 640), empty[:, :4], (1200, 1920))
    assert out.shape == (0, 4)
    assert m.scale_boxes((640, 640), torch.zeros((0, 4), device=cuda_dev), (1200, 1920)).shape == (0, 4)
    pred = torch.zeros((2, 4 + 8, 8400), device=cuda_dev)                    # nothing above conf
    dets = m.non_max_suppression(pred, 0.25, 0.45)
    assert [tuple(d.shape) for d in dets] == [(0, 6), (0, 6)]
    for d in dets:
        assert m.scale_boxes((640, 640), d[:, :4], (1200, 1920)).shape == (0, 4)
    # too many anchors for the 16-bit anchor field of the sort key: refused, not mis-sorted
    with pytest.raises(ValueError):
        m.filter_decoded(torch.zeros((1, 4 + 2, 70000), device=cuda_dev), 0.25)
