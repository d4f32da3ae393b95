"""BASELINE.json configurations at their FULL sizes, checked through size-independent properties (the oracle takes
minutes at these sizes): agreement of independent kernel chains, sortedness, idempotence of NMS on its own output,
exact reproduction of constants by the resamplers, and oracle spot checks on sub-samples."""
import numpy as np
import pytest
import torch

import manual_yolo_b200 as m
from manual_yolo_b200 import geometry, synth
from oracle import boxes as oboxes
from oracle import roi as oroi

pytestmark = pytest.mark.gpu


def _nms_properties(det, conf, max_det):
    counts = det.count.cpu().tolist()
    rows, anchor = det.rows.cpu(), det.anchor.cpu()
    for b, k in enumerate(counts):
        assert 0 <= k <= max_det
        r, a = rows[b, :k], anchor[b, :k]
        if k == 0:
            continue
        s = r[:, 4]
        assert bool((s > conf).all())
        assert bool((s[:-1] >= s[1:]).all())                                        # score order
        tie = s[:-1] == s[1:]
        assert bool((a[:-1][tie] < a[1:][tie]).all())                               # ties: anchor order
        assert a.unique().numel() == k                                              # an anchor is kept once
        assert bool((r[:, 5] == r[:, 5].floor()).all())
    return counts


def _nms_idempotent(det, iou, dev, agnostic=False):
    """NMS applied to its own (unscaled) output must return it unchanged: no kept pair can suppress each other."""
    B, max_det, _ = det.rows.shape
    cands = m.Candidates(det.rows.clone(), torch.arange(max_det, dtype=torch.int32, device=dev).repeat(B, 1).contiguous(),
                         det.count.clone(), max_det)
    again = m.nms_candidates(cands, iou, agnostic=agnostic, max_det=max_det)
    assert torch.equal(again.count, det.count)
    for b, k in enumerate(det.count.cpu().tolist()):
        assert torch.equal(again.rows[b, :k], det.rows[b, :k])


def test_config2_full_batch_properties(cuda_dev):
    """configs[1]: 64 synthetic 1920x1200 frames, 8400 anchors, nc=64, conf 0.25, iou 0.45."""
    B, nc, src_hw, conf, iou = 64, 64, (1200, 1920), 0.25, 0.45
    frames = synth.synth_frames(B, *src_hw, seed=0)
    frames[1] = 77                                                                  # a constant frame
    fused = m.Pipeline(B, src_hw, nc, conf=conf, iou=iou, device=cuda_dev, cap=1024)
    general = m.Pipeline(B, src_hw, nc, conf=conf, iou=iou, device=cuda_dev, cap=None)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=fused.in_hw, src_hw=src_hw, seed=0, conf_thres=conf)
    fd, hd = frames.to(cuda_dev), head.to(cuda_dev)
    r1, r2 = fused(fd, hd), general(fd, hd)
    torch.cuda.synchronize()
    # two independent chains (fused sparse kernel / class filter + select-sort + decode + windowed NMS) agree bit for bit
    assert torch.equal(r1.det.count, r2.det.count) and torch.equal(r1.net_in, r2.net_in)
    counts = _nms_properties(r1.det, conf, 300)
    for b, k in enumerate(counts):
        assert torch.equal(r1.det.rows[b, :k], r2.det.rows[b, :k]) and torch.equal(r1.det.anchor[b, :k], r2.det.anchor[b, :k])
    assert sum(counts) > 1000
    rows = r1.det.rows.cpu()
    for b, k in enumerate(counts):                                                  # scale_boxes + clip: inside the frame
        bx = rows[b, :k, :4]
        assert bool((bx[:, [0, 2]] >= 0).all() and (bx[:, [0, 2]] <= src_hw[1]).all())
        assert bool((bx[:, [1, 3]] >= 0).all() and (bx[:, [1, 3]] <= src_hw[0]).all())
    # letterbox: the padding rows are exactly 114/255, the constant frame stays constant inside, values in [0,1]
    g = fused.geom
    pad = torch.tensor(114, dtype=torch.float32).div(255).item()
    assert bool((r1.net_in[:, :, :g["top"], :] == pad).all() and (r1.net_in[:, :, g["top"] + g["new_h"]:, :] == pad).all())
    inner = r1.net_in[1, :, g["top"]:g["top"] + g["new_h"], g["left"]:g["left"] + g["new_w"]]
    assert bool((inner == torch.tensor(77, dtype=torch.float32).div(255).item()).all())
    assert float(r1.net_in.min()) >= 0.0 and float(r1.net_in.max()) <= 1.0
    # ROI stage: as many ROIs as rank-class detections, image-major; the constant frame gives constant crops
    n = r1.n_rois()
    want = sum(int(c) in m.pipeline.RANK_CLASS_IDS for b, k in enumerate(counts) for c in rows[b, :k, 5].tolist())
    assert int(r1.roi_count) == want and n == min(want, fused.roi_cap) and n > 100      # roi_count is not clamped
    rb = r1.roi_batch[:n].cpu()
    assert bool((rb[:-1] <= rb[1:]).all()) and set(r1.roi_valid[:n].cpu().tolist()) <= {1, 2}
    for i in (rb == 1).nonzero().view(-1).tolist():
        assert bool((r1.rois[i] == torch.tensor(77, dtype=torch.float32).div(255).item()).all())
    assert torch.equal(r1.rois[:n], r2.rois[:n])
    # idempotence on the letterboxed (unscaled) detections
    cands = m.decode_and_filter(hd, conf_thres=conf, level_hw=fused.level_hw)
    det = m.nms_candidates(cands, iou)
    _nms_idempotent(det, iou, cuda_dev)
    # host-fed form at full size == device-resident form
    runner = m.HostRunner(fused, stage="rows", dfl_zero_copy=True)
    h_rows, h_count, h_roi = runner.submit(frames.pin_memory(), head.pin_memory())
    runner.wait()
    torch.cuda.synchronize()
    assert torch.equal(h_count, r2.det.count.cpu()) and int(h_roi) == n
    for b, k in enumerate(counts):
        assert torch.equal(h_rows[b, :k], rows[b, :k])


def test_config3_full_batch_properties(cuda_dev):
    """configs[2]: 256 images x 8400 candidates, nc=80, conf 0.001, max_det 300 (NMS-heavy evaluation regime)."""
    B, nc, conf, iou = 256, 80, 0.001, 0.7
    lv = geometry.level_shapes(640, 640)
    head = torch.cat([synth.synth_head_dense(64, nc, seed=s) for s in range(B // 64)]).to(cuda_dev)
    stage = m.nms_candidates(m.decode_and_filter(head, conf_thres=conf, level_hw=lv), iou, max_det=300)
    s_rows, s_anchor, s_count = stage.rows.clone(), stage.anchor.clone(), stage.count.clone()
    cands = m.decode_and_filter(head, conf_thres=conf, level_hw=lv, defer_boxes=True)
    assert int(cands.count.min()) == 8400
    ws = m.Workspace(B, cands.cap, 300, cuda_dev)
    det = m.postprocess_dense(cands, ws, head, level_hw=lv, iou_thres=iou, max_det=300)
    counts = _nms_properties(det, conf, 300)
    assert counts == [300] * B and torch.equal(det.count, s_count)
    assert torch.equal(det.anchor, s_anchor) and torch.equal(det.rows, s_rows)       # dense chain == stage-wise kernels
    _nms_idempotent(det, iou, cuda_dev)
    # heavy max_det: no early exit anywhere, still sorted / unique / idempotent
    ws2 = m.Workspace(B, cands.cap, 1000, cuda_dev)
    cands.count.zero_()
    cands = m.decode_and_filter(head, conf_thres=conf, level_hw=lv, defer_boxes=True, out=cands)
    det2 = m.postprocess_dense(cands, ws2, head, level_hw=lv, iou_thres=0.45, max_det=1000)
    c2 = _nms_properties(det2, conf, 1000)
    assert min(c2) > 300
    _nms_idempotent(det2, 0.45, cuda_dev)


def test_config4_full_batch_properties(cuda_dev):
    """configs[3]: 4096 ROIs from 64 frames of 1920x1200 -> (4096,3,64,64)."""
    B, N = 64, 4096
    frames = synth.synth_frames(B, 1200, 1920, seed=0)
    frames[5] = 200
    boxes, bidx = synth.synth_rois(N, B, seed=0)
    out, valid = m.crop_resize_rois(frames.to(cuda_dev), boxes.to(cuda_dev), bidx.to(cuda_dev), pad=6)
    torch.cuda.synchronize()
    v = valid.cpu()
    assert set(v.tolist()) <= {1, 2} and int((v == 1).sum()) > 3900
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    const = torch.tensor(200, dtype=torch.float32).div(255).item()
    on5 = (bidx == 5).nonzero().view(-1).tolist()
    assert len(on5) > 30 and all(bool((out[i] == const).all()) for i in on5)        # resampling a constant is exact
    # mean preservation: a bilinear resample of noise keeps the crop mean to within a few percent
    sel = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:48].tolist()
    for i in sel:                                                                   # oracle spot check, bit-exact
        crop = oboxes.safe_crop_ref(frames[bidx[i]].numpy(), *[int(t) for t in boxes[i]], pad=6)
        assert torch.equal(out[i].cpu(), oroi.classify_preprocess_ref(crop)), i
    # the same boxes read zero-copy from pinned host frames give the same bytes
    out2, valid2 = m.crop_resize_rois(frames.pin_memory(), boxes.to(cuda_dev), bidx.to(cuda_dev), pad=6)
    assert torch.equal(out2, out) and torch.equal(valid2, valid)


def _oracle_vs_det(det, head_cpu, lv, conf, iou, images, in_hw=None, src_hw=None, max_det=300):
    """Kept anchor indices + class ids bit-exact, boxes/scores <= 1e-4 against the CPU oracle for the given images."""
    from oracle import head as ohead
    from oracle import nms as onms
    pred = ohead.detect_inference_ref(head_cpu[images], lv)
    out, idx = onms.non_max_suppression_ref(pred, conf, iou, max_det=max_det, return_idxs=True)
    worst = 0.0
    for j, b in enumerate(images):
        k = int(det.count[b])
        assert k == out[j].shape[0], (b, k, out[j].shape)
        assert torch.equal(det.anchor[b, :k].cpu().long(), idx[j]), b                # kept sets, in order
        got, exp = det.rows[b, :k].cpu(), out[j].clone()
        if src_hw is not None:
            exp[:, :4] = oboxes.scale_boxes_ref(in_hw, exp[:, :4], src_hw)
        assert torch.equal(got[:, 5], exp[:, 5]), b                                  # class ids
        if k:
            worst = max(worst, float((got[:, :5] - exp[:, :5]).abs().max()))
    assert worst <= 1e-4, worst
    return worst


def test_config2_bench_batch_equals_oracle_all_64_frames(cuda_dev):
    """The EXACT batch bench.py times (seed 0, 64 frames of 1920x1200, label-derived head, conf 0.25, iou 0.45) against the
    CPU oracle for every frame: letterbox bit-exact, kept sets / class ids bit-exact, boxes + scores <= 1e-4, every ROI
    <= 1/255 (observed 0)."""
    from oracle import letterbox as olb
    B, nc, src_hw, conf, iou = 64, 64, (1200, 1920), 0.25, 0.45
    frames = synth.synth_frames(B, *src_hw, seed=0)
    pipe = m.Pipeline(B, src_hw, nc, conf=conf, iou=iou, device=cuda_dev, cap=1024)
    head, _ = synth.synth_head_from_labels(B, nc, in_hw=pipe.in_hw, src_hw=src_hw, seed=0, conf_thres=conf)
    res = pipe(frames.to(cuda_dev), head.to(cuda_dev))
    torch.cuda.synchronize()
    assert pipe.check_overflow() <= pipe.cap
    for b0 in range(0, B, 16):
        ref = olb.preprocess_ref(list(frames[b0:b0 + 16].numpy()), (640, 640))
        assert torch.equal(res.net_in[b0:b0 + 16].cpu(), ref)
    _oracle_vs_det(res.det, head, pipe.level_hw, conf, iou, list(range(B)), in_hw=pipe.in_hw, src_hw=src_hw)
    n = res.n_rois()
    rb, rd = res.roi_batch[:n].cpu().tolist(), res.roi_det[:n].cpu().tolist()
    rows = res.det.rows.cpu()
    want = [(b, i) for b in range(B) for i in range(int(res.det.count[b])) if int(rows[b, i, 5]) in m.pipeline.RANK_CLASS_IDS]
    assert list(zip(rb, rd)) == want[:pipe.roi_cap] and int(res.roi_count) == len(want) and n == min(len(want), pipe.roi_cap) and n > 100
    rois = res.rois[:n].cpu()
    for g, (b, i) in enumerate(zip(rb, rd)):
        crop = oboxes.safe_crop_ref(frames[b].numpy(), *[int(v) for v in rows[b, i, :4]], pad=6)
        assert crop is not None and torch.equal(rois[g], oroi.classify_preprocess_ref(crop)), (g, b, i)


@pytest.mark.parametrize("iou", [0.7, 0.45])
def test_config3_full_batch_oracle_subsample_of_16(cuda_dev, iou):
    """configs[2] at its full size (256 x 144 x 8400, conf 0.001): 16 images drawn from the whole batch (every seed
    block, incl. the adversarial columns) equal the CPU oracle -- kept anchors + class ids bit-exact, boxes/scores 1e-4."""
    B, nc, conf = 256, 80, 0.001
    lv = geometry.level_shapes(640, 640)
    head_cpu = torch.cat([synth.synth_head_dense(64, nc, seed=s) for s in range(B // 64)])
    head = head_cpu.to(cuda_dev)
    cands = m.decode_and_filter(head, conf_thres=conf, level_hw=lv, defer_boxes=True)
    ws = m.Workspace(B, cands.cap, 300, cuda_dev)
    det = m.postprocess_dense(cands, ws, head, level_hw=lv, iou_thres=iou, max_det=300)
    torch.cuda.synchronize()
    images = [0, 17, 34, 63, 64, 81, 100, 127, 128, 150, 171, 191, 192, 213, 234, 255]
    _oracle_vs_det(det, head_cpu, lv, conf, iou, images)
