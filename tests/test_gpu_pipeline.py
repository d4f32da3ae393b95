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


@pytest.mark.parametrize("src_hw,auto,conf,iou", [((1200, 1920), False, 0.25, 0.45), ((900, 1600), True, 0.25, 0.7),
                                                  ((1200, 1920), False, 0.25, 0.7)])
def test_pipeline_matches_oracle_chain(cuda_dev, golden_dir, src_hw, auto, conf, iou):
    B, nc = 4, 64
    frames = synth.synth_frames(B, *src_hw, seed=11).numpy().copy()
    real = cv2.imread(os.path.join(golden_dir, "frames", f"frame_{src_hw[1]}x{src_hw[0]}.jpg"))
    frames[0] = real                                                   # one real dataset frame per batch
    pipe = m.Pipeline(B, src_hw, nc, imgsz=640, auto=auto, conf=conf, iou=iou, device=cuda_dev)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe.in_hw, src_hw=src_hw, seed=11, conf_thres=conf)
    res = pipe(torch.from_numpy(frames).to(cuda_dev), head.to(cuda_dev))
    ndet, nroi = _assert_matches(res, _oracle_chain(frames, head, pipe, conf, iou), B)
    assert ndet > 40 and nroi > 4


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
    # host-facing entry: pinned host buffers in, host results out
    rows, count, nroi = pipe.run_host(frames.pin_memory(), head.pin_memory())
    torch.cuda.synchronize()
    for b in range(B):
        assert int(count[b]) == ref[1][b].shape[0]
        assert (rows[b, :int(count[b]), :5] - ref[1][b][:, :5]).abs().max().item() <= 1e-4
    assert int(nroi) == len(ref[3])
    recs = m.pipeline.detections_to_records(rows, count)
    assert len(recs) == int(count.sum()) and all(len(r["bbox"]) == 4 for r in recs)
