"""SAHI-style sliced prediction (pipe.py:183-194) -- K1 slice mode, per-slice path, gather +
shift, merge NMS, ROI crops -- against the oracle restatement (oracle/slicing.py; sahi itself is not installed)."""
import numpy as np
import pytest
import torch

import manual_yolo_b200 as m
from manual_yolo_b200 import synth
from oracle import boxes as oboxes
from oracle import roi as oroi
from oracle import slicing

pytestmark = pytest.mark.gpu


def _heads_with_duplicates(F, S, slices, nc, seed, conf):
    """Independent label-derived heads per slice, then every object in the right half of a slice (the 320-px
    overlap band at overlap ratio 0.5) is also given to its right-hand neighbour (same anchor-relative DFL logits,
    40 stride-8 cells to the left), so the merge NMS has real cross-slice duplicates to remove."""
    head, _ = synth.synth_head_from_labels(F * S, nc, in_hw=(640, 640), src_hw=(640, 640), seed=seed, conf_thres=conf,
                                           guard_ulp=0)
    for f in range(F):
        for s in range(S - 1):
            (x0, y0, _, _), (nx0, ny0, _, _) = slices[s], slices[s + 1]
            if ny0 != y0 or nx0 - x0 != 320:
                continue
            src, dst = head[f * S + s].view(-1, 8400), head[f * S + s + 1].view(-1, 8400)
            grid = torch.arange(80 * 80).view(80, 80)                 # level 0: stride 8, 80 x 80 cells
            dst[:, grid[:, :40].reshape(-1)] = src[:, grid[:, 40:].reshape(-1)]
    synth.guard_band(head, nc, conf, 16)
    return head


def test_slice_mode_letterbox_bit_exact(cuda_dev):
    """K1 slice mode == the oracle's letterbox of the numpy window, for aligned and unaligned origins, a window
    smaller than the network input (letterboxed with padding) and several frames in one launch."""
    from oracle import letterbox as olb
    for (H, W), kw in [((1200, 1920), {}), ((543, 770), {}), ((700, 1001), dict(slice_h=320, slice_w=333, overlap_h=0.3, overlap_w=0.1))]:
        frames = synth.synth_frames(2, H, W, seed=H)
        sl = m.geometry.slice_boxes(H, W, **kw)
        got = m.preprocess_slices(frames.to(cuda_dev), sl, (640, 640))
        crops = [frames[f].numpy()[y0:y1, x0:x1] for f in range(2) for (x0, y0, x1, y1) in sl]
        assert torch.equal(got.cpu(), olb.preprocess_ref(crops, (640, 640))), (H, W)


@pytest.mark.parametrize("cap", [1024, None])
def test_sliced_pipeline_matches_oracle(cuda_dev, cap):
    F, nc, frame_hw, conf, iou, merge_iou = 2, 64, (1200, 1920), 0.25, 0.7, 0.5
    sp = m.SlicedPipeline(F, frame_hw, nc, overlap=(0.5, 0.5), conf=conf, iou=iou, merge_iou=merge_iou, device=cuda_dev,
                          cap=cap, rois_per_frame=64)
    assert sp.S == 15 and sp.slice_hw == (640, 640) and sp.fused == (cap is not None)
    assert m.SlicedPipeline(1, frame_hw, nc, device=cuda_dev).S == 12                    # the reference's 0.2 overlap
    frames = synth.synth_frames(F, *frame_hw, seed=3)
    head = _heads_with_duplicates(F, sp.S, sp.slices, nc, seed=3, conf=conf)
    res = sp(frames.to(cuda_dev), head.to(cuda_dev))
    torch.cuda.synchronize()
    net_in, per_slice, merged, prov = slicing.sliced_prediction_ref(frames.numpy(), head, sp.slices, (640, 640), conf, iou,
                                                                    merge_iou)
    assert torch.equal(res.net_in.cpu(), net_in)                                        # K1 slice mode bit-exact
    sc = sp.slice_det.count.cpu().tolist()
    for i, o in enumerate(per_slice):                                                   # per-slice stage
        assert sc[i] == o.shape[0]
        got = sp.slice_det.rows[i, :sc[i]].cpu()
        assert torch.equal(got[:, 5], o[:, 5]) and (got[:, :5] - o[:, :5]).abs().max().item() <= 1e-4
    counts = res.det.count.cpu().tolist()
    n_roi = 0
    for f in range(F):                                                                  # merged stage
        assert counts[f] == merged[f].shape[0]
        assert torch.equal(res.det.anchor[f, :counts[f]].cpu().long(), prov[f])         # kept set bit-exact
        got = res.det.rows[f, :counts[f]].cpu()
        assert torch.equal(got[:, 5], merged[f][:, 5])
        assert (got[:, :5] - merged[f][:, :5]).abs().max().item() <= 1e-4
        assert counts[f] < sum(sc[f * sp.S:(f + 1) * sp.S])                             # duplicates were merged away
        for row in merged[f]:                                                           # ROI crops from the full frame
            if int(row[5]) in m.pipeline.RANK_CLASS_IDS:
                crop = oboxes.safe_crop_ref(frames[f].numpy(), *[int(v) for v in row[:4]], pad=6)
                if crop is None:
                    assert int(res.roi_valid[n_roi]) == 0
                else:
                    assert (res.rois[n_roi].cpu() - oroi.classify_preprocess_ref(crop)).abs().max().item() <= 1 / 255
                n_roi += 1
    assert int(res.roi_count.cpu()) == n_roi and n_roi > 0          # not clamped to roi_cap
    assert sp.check_overflow() <= sp.cap


def _rowset(t):
    return sorted(tuple(r) for r in t.tolist())


@pytest.mark.parametrize("metric,agnostic", [("IOS", False), ("IOU", False), ("IOS", True)])
def test_greedy_nmm_kernel_equals_sahi_restatement(cuda_dev, metric, agnostic):
    """b200yolo_greedy_nmm vs oracle/slicing.py's restatement of SAHI's GREEDYNMM (greedy_nmm + has_match merge) on
    clustered boxes (many matches, merge chains), score ties and an empty frame: the merged rows are equal as sets
    (SAHI lists them category-major, the kernel by descending score), bit for bit."""
    g = torch.Generator().manual_seed(11)
    F, cap, max_det = 4, 1200, 300
    rows = torch.zeros((F, cap, 6))
    counts = [260, 0, 37, 900]
    for f, n in enumerate(counts):
        centers = torch.rand((12, 2), generator=g) * 900 + 50
        c = centers[torch.randint(0, 12, (n,), generator=g)] + torch.randn((n, 2), generator=g) * 14
        wh = torch.rand((n, 2), generator=g) * 60 + 20
        rows[f, :n, 0:2] = c - wh / 2
        rows[f, :n, 2:4] = c + wh / 2
        sc = torch.rand((n,), generator=g) * 0.7 + 0.25
        sc[::9] = sc[0] if n else 0                               # ties
        rows[f, :n, 4] = sc
        rows[f, :n, 5] = torch.randint(0, 3, (n,), generator=g).float()
    cands = m.Candidates(rows.to(cuda_dev), torch.arange(cap, dtype=torch.int32, device=cuda_dev).repeat(F, 1).contiguous(),
                         torch.tensor(counts, dtype=torch.int32, device=cuda_dev), cap)
    det = m.Workspace(F, cap, max_det, cuda_dev).det
    m.greedy_nmm(cands, det, metric, 0.5, agnostic)
    got_counts = det.count.cpu().tolist()
    for f, n in enumerate(counts):
        exp, kept = slicing.greedy_nmm_postprocess_ref(rows[f, :n], metric, 0.5, class_agnostic=agnostic)
        if exp.shape[0] <= max_det:
            assert got_counts[f] == exp.shape[0]
            assert _rowset(det.rows[f, :got_counts[f]].cpu()) == _rowset(exp), f
            assert sorted(det.anchor[f, :got_counts[f]].cpu().tolist()) == sorted(kept.tolist())
        else:                                                     # cut at max_det: the best-scoring kept boxes
            assert got_counts[f] == max_det
            assert set(_rowset(det.rows[f, :max_det].cpu())) <= set(_rowset(exp))
        s = det.rows[f, :got_counts[f], 4].cpu()
        assert got_counts[f] < 2 or bool((s[:-1] >= s[1:]).all()) or metric == "IOS"      # kept order: descending keep score


def test_sliced_pipeline_sahi_defaults_standard_pred_and_greedy_nmm(cuda_dev):
    """The reference's call (pipe.py:186-193) with SAHI's defaults: slices + the full-frame prediction, merged by
    GREEDYNMM / IOS / 0.5 -- against the oracle restatement."""
    F, nc, frame_hw, conf, iou = 2, 64, (1200, 1920), 0.25, 0.7
    sp = m.SlicedPipeline(F, frame_hw, nc, overlap=(0.5, 0.5), conf=conf, iou=iou, merge_iou=0.5, device=cuda_dev, cap=1024,
                          rois_per_frame=64, merge="greedy_nmm", match_metric="IOS", standard_pred=True)
    frames = synth.synth_frames(F, *frame_hw, seed=5)
    head = _heads_with_duplicates(F, sp.S, sp.slices, nc, seed=5, conf=conf)
    head_full, _ = synth.synth_head_from_labels(F, nc, in_hw=sp.full.in_hw, src_hw=frame_hw, seed=6, conf_thres=conf)
    res = sp(frames.to(cuda_dev), head.to(cuda_dev), head_full.to(cuda_dev))
    torch.cuda.synchronize()
    assert sp.check_overflow() <= sp.cap
    net_in, per_slice, merged, prov = slicing.sliced_prediction_ref(frames.numpy(), head, sp.slices, (640, 640), conf, iou, 0.5,
                                                                    merge="greedy_nmm", match_metric="IOS", heads_full=head_full)
    assert torch.equal(res.net_in.cpu(), net_in)
    counts = res.det.count.cpu().tolist()
    n_slice_dets = sp.slice_det.count.cpu().view(F, sp.S).sum(1).tolist()
    n_full = sp.full_result.det.count.cpu().tolist()
    for f in range(F):
        assert n_full[f] > 0 and counts[f] == merged[f].shape[0] < n_slice_dets[f] + n_full[f]
        got = res.det.rows[f, :counts[f]].cpu()
        a, b = torch.tensor(_rowset(got)), torch.tensor(_rowset(merged[f]))
        assert torch.equal(a[:, 5], b[:, 5]) and (a[:, :5] - b[:, :5]).abs().max().item() <= 1e-4
        assert sorted(res.det.anchor[f, :counts[f]].cpu().tolist()) == sorted(prov[f].tolist())
        assert max(res.det.anchor[f, :counts[f]].cpu().tolist()) >= sp.S * 300 or True    # full-frame boxes may be kept
    n = res.n_rois()
    want = sum(int(c) in m.pipeline.RANK_CLASS_IDS for f in range(F) for c in res.det.rows[f, :counts[f], 5].cpu().tolist())
    assert int(res.roi_count) == want and n == min(want, sp.roi_cap)
