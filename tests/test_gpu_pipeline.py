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
