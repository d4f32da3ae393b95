"""K3/K4 parity (rows a8-a10): kept anchor indices, class ids and rows BIT-EXACT against the oracle
(real torchvision.ops.nms) on identical score bits."""
import os

import numpy as np
import pytest
import torch
import torchvision

import manual_yolo_b200 as m
from manual_yolo_b200 import geometry, synth
from oracle import boxes as oboxes
from oracle import head as ohead
from oracle import nms as onms

pytestmark = pytest.mark.gpu


def _check(pred, cuda_dev, **kw):
    ref_out, ref_idx = onms.non_max_suppression_ref(pred, return_idxs=True, **kw)
    out, idx = m.non_max_suppression(pred.to(cuda_dev), return_idxs=True, **kw)
    assert len(out) == len(ref_out)
    kept = 0
    for o, i, ro, ri in zip(out, idx, ref_out, ref_idx):
        assert torch.equal(i.cpu(), ri), "kept anchor indices differ"
        assert torch.equal(o.cpu(), ro), "kept rows differ"
        kept += len(ri)
    return kept


@pytest.mark.parametrize("iou", [0.45, 0.5, 0.6, 0.7])
def test_dense_eval_regime_bit_exact(cuda_dev, iou):
    """n ~ 8400 candidates/image, nc=80, adversarial ties / degenerate boxes / class-79 offsets (config 3)."""
    head = synth.synth_head_dense(3, 80, seed=int(iou * 100))
    pred = ohead.detect_inference_ref(head, geometry.level_shapes(640, 640))
    kept = _check(pred, cuda_dev, conf_thres=0.001, iou_thres=iou, max_det=300)
    assert kept == 3 * 300


@pytest.mark.parametrize("conf,iou", [(0.25, 0.45), (0.25, 0.7), (0.5, 0.7), (0.35, 0.7)])
def test_label_derived_bit_exact(cuda_dev, conf, iou):
    head, _ = synth.synth_head_from_labels(6, 64, seed=1, conf_thres=conf)
    pred = ohead.detect_inference_ref(head, geometry.level_shapes(640, 640))
    assert _check(pred, cuda_dev, conf_thres=conf, iou_thres=iou) > 50


def test_options_max_det_agnostic_classes_max_nms(cuda_dev):
    head = synth.synth_head_dense(2, 80, seed=9)
    pred = ohead.detect_inference_ref(head, geometry.level_shapes(640, 640))
    _check(pred, cuda_dev, conf_thres=0.01, iou_thres=0.7, max_det=17)
    _check(pred, cuda_dev, conf_thres=0.01, iou_thres=0.5, agnostic=True, max_det=1000)
    _check(pred, cuda_dev, conf_thres=0.001, iou_thres=0.7, classes=[0, 5, 79])
    _check(pred, cuda_dev, conf_thres=0.05, iou_thres=0.7, max_det=3000)      # no early exit: full sweep
    # n > max_nms keeps the top max_nms scores (ties at the cut are unspecified upstream: none here)
    _check(pred, cuda_dev, conf_thres=0.001, iou_thres=0.7, max_nms=1000)


def test_medium_sizes_cover_enumeration_and_radix_paths(cuda_dev):
    g = torch.Generator().manual_seed(5)
    for n_obj in (1, 40, 511, 512, 513, 700, 1500, 4000):
        A = 8400
        pred = torch.zeros((1, 4 + 16, A))
        pred[0, 0] = torch.rand(A, generator=g) * 600
        pred[0, 1] = torch.rand(A, generator=g) * 600
        pred[0, 2] = torch.rand(A, generator=g) * 90 + 5
        pred[0, 3] = torch.rand(A, generator=g) * 90 + 5
        sel = torch.randperm(A, generator=g)[:n_obj]
        pred[0, 4 + torch.randint(0, 16, (n_obj,), generator=g), sel] = torch.rand(n_obj, generator=g) * 0.7 + 0.3
        _check(pred, cuda_dev, conf_thres=0.25, iou_thres=0.45, max_det=300)


def test_ties_duplicates_degenerates(cuda_dev):
    A = 64
    pred = torch.zeros((2, 4 + 3, A))
    # image 0: identical boxes with identical scores (stable tie -> lower anchor kept), duplicates in
    # other classes survive, zero-area twins both survive (0/0 = NaN never suppresses)
    pred[0, :4, 5] = torch.tensor([100., 100., 50., 50.]); pred[0, 4, 5] = 0.9
    pred[0, :4, 9] = torch.tensor([100., 100., 50., 50.]); pred[0, 4, 9] = 0.9
    pred[0, :4, 2] = torch.tensor([100., 100., 50., 50.]); pred[0, 5, 2] = 0.9
    pred[0, :4, 20] = torch.tensor([300., 300., 0., 0.]); pred[0, 6, 20] = 0.8
    pred[0, :4, 21] = torch.tensor([300., 300., 0., 0.]); pred[0, 6, 21] = 0.8
    # IoU exactly float32(0.6): suppressed at iou_thres=0.6 (double compare), kept at float32(0.6)
    pred[1, :4, 0] = torch.tensor([2.5, 0.5, 5., 1.]); pred[1, 4, 0] = 0.9
    pred[1, :4, 1] = torch.tensor([1.5, 0.5, 3., 1.]); pred[1, 4, 1] = 0.8
    out = m.non_max_suppression(pred.to(cuda_dev), 0.25, 0.45, return_idxs=True)
    assert out[1][0].cpu().tolist() == [2, 5, 20, 21]
    _check(pred, cuda_dev, conf_thres=0.25, iou_thres=0.45)
    _check(pred, cuda_dev, conf_thres=0.25, iou_thres=0.6)
    _check(pred, cuda_dev, conf_thres=0.25, iou_thres=float(np.float32(0.6)))
    a = m.non_max_suppression(pred.to(cuda_dev), 0.25, 0.6)[1]
    b = m.non_max_suppression(pred.to(cuda_dev), 0.25, float(np.float32(0.6)))[1]
    assert a.shape[0] == 1 and b.shape[0] == 2


def test_empty_images_and_ragged_batch(cuda_dev):
    pred = torch.zeros((3, 4 + 5, 100))
    pred[1, :4, 7] = torch.tensor([50., 50., 20., 20.]); pred[1, 6, 7] = 0.6
    out = m.non_max_suppression(pred.to(cuda_dev), 0.25, 0.45)
    assert [o.shape for o in out] == [(0, 6), (1, 6), (0, 6)]
    _check(pred, cuda_dev, conf_thres=0.25, iou_thres=0.45)


def test_torchvision_golden(cuda_dev, golden_dir):
    """Kept indices recorded from torchvision.ops.nms in the dev container (tests/golden/nms_golden.npz)."""
    z = np.load(os.path.join(golden_dir, "nms_golden.npz"))
    for t in range(4):
        b, s, c, thr = z[f"boxes{t}"], z[f"scores{t}"], z[f"cls{t}"], float(z[f"thr{t}"])
        n, nc = len(s), int(c.max()) + 1
        rows = torch.zeros((1, n, 6))
        rows[0, :, :4] = torch.from_numpy(b); rows[0, :, 4] = torch.from_numpy(s); rows[0, :, 5] = torch.from_numpy(c)
        cands = m.Candidates(rows.to(cuda_dev), torch.arange(n, dtype=torch.int32, device=cuda_dev)[None].contiguous(),
                             torch.tensor([n], dtype=torch.int32, device=cuda_dev), n)
        det = m.nms_candidates(cands, thr, max_det=4096)
        k = int(det.count[0])
        assert np.array_equal(det.anchor[0, :k].cpu().numpy(), z[f"keep{t}"][:4096])


def test_scale_boxes_bit_exact(cuda_dev):
    g = torch.Generator().manual_seed(3)
    for img1, img0 in [((640, 640), (1200, 1920)), ((384, 640), (900, 1600)), ((640, 544), (1130, 930)),
                       ((928, 1280), (543, 770))]:
        b = torch.rand((500, 6), generator=g) * 700 - 30
        ref = oboxes.scale_boxes_ref(img1, b[:, :4], img0)
        got = m.scale_boxes(img1, b.clone().to(cuda_dev), img0)
        assert torch.equal(got[:, :4].cpu(), ref) and torch.equal(got[:, 4:].cpu(), b[:, 4:])
    # fused form inside the NMS epilogue
    head, _ = synth.synth_head_from_labels(2, 64, seed=8)
    pred = ohead.detect_inference_ref(head, geometry.level_shapes(640, 640))
    ref = onms.non_max_suppression_ref(pred, 0.25, 0.7)
    cands = m.filter_decoded(pred.to(cuda_dev), 0.25)
    scale = m.scale_params_tensor((640, 640), [(1200, 1920)] * 2, cuda_dev)
    det = m.nms_candidates(cands, 0.7, scale=scale)
    for bi in range(2):
        exp = ref[bi].clone()
        exp[:, :4] = oboxes.scale_boxes_ref((640, 640), exp[:, :4], (1200, 1920))
        assert torch.equal(det.rows[bi, :int(det.count[bi])].cpu(), exp)


def test_oversize_images_use_workspace_paths(cuda_dev):
    """pipe.py runs imgsz=1280 (928x1280 rect -> 24 360 anchors): with a low threshold an image has more
    candidates than fit the shared-memory sort/NMS paths (12 288 keys / 10 240 boxes), so the L2-resident
    workspace paths run.  Same bit-exact bar."""
    g = torch.Generator().manual_seed(11)
    A, nc = 24360, 8
    pred = torch.zeros((2, 4 + nc, A))
    pred[:, 0] = torch.rand((2, A), generator=g) * 1200
    pred[:, 1] = torch.rand((2, A), generator=g) * 900
    pred[:, 2] = torch.rand((2, A), generator=g) * 60 + 4
    pred[:, 3] = torch.rand((2, A), generator=g) * 60 + 4
    pred[:, 4:] = torch.rand((2, nc, A), generator=g) * 0.5
    pred[1, 4:, 15000:] = 0.0                                   # image 1: 15 000 candidates, image 0: all 24 360
    kept = _check(pred, cuda_dev, conf_thres=0.05, iou_thres=0.6, max_det=500)
    assert kept == 1000
    _check(pred, cuda_dev, conf_thres=0.05, iou_thres=0.45, max_det=300, max_nms=13000)
    # heavy overlap at this size: few keeps -> the NMS exhausts the ordered prefix -> exact fallback, whose full sort
    # runs in the workspace (keys do not fit shared memory)
    centres = torch.rand((30, 2), generator=g) * 800 + 50
    which = torch.randint(0, 30, (A,), generator=g)
    pred[0, 0] = centres[which, 0] + torch.randn(A, generator=g)
    pred[0, 1] = centres[which, 1] + torch.randn(A, generator=g)
    pred[0, 2:4] = 50.0
    kept = _check(pred[:1], cuda_dev, conf_thres=0.05, iou_thres=0.5, max_det=300)
    assert kept < 300


def test_dense_sort_prefix_and_fallback(cuda_dev):
    """cap > 2048: K3 orders only the best 2048 entries (radix select + bitonic sort).  (a) max_det reached inside the
    prefix: no fallback; (b) heavy overlap -- 8400 candidates in ~40 clusters per class, far fewer keeps than max_det:
    the NMS runs out of ordered entries, flags the image, and the exact full-sort fallback produces the result;
    (c) a batch mixing both kinds, plus an image with fewer than 2048 candidates."""
    g = torch.Generator().manual_seed(17)
    A, nc = 8400, 4

    def clustered(n_clusters, n_live):
        pred = torch.zeros((4 + nc, A))
        centres = torch.rand((n_clusters, 2), generator=g) * 560 + 40
        which = torch.randint(0, n_clusters, (A,), generator=g)
        pred[0] = centres[which, 0] + torch.randn(A, generator=g) * 1.5
        pred[1] = centres[which, 1] + torch.randn(A, generator=g) * 1.5
        pred[2] = 60 + torch.rand(A, generator=g) * 4
        pred[3] = 60 + torch.rand(A, generator=g) * 4
        live = torch.randperm(A, generator=g)[:n_live]
        sc = torch.rand(n_live, generator=g) * 0.98 + 0.01
        sc[::7] = sc[0]                                                      # ties: anchor order decides
        pred[4 + (which[live] % nc), live] = sc
        return pred

    spread = ohead.detect_inference_ref(synth.synth_head_dense(1, nc, seed=3), geometry.level_shapes(640, 640))[0]
    batch = torch.stack([clustered(40, A), spread, clustered(25, 5000), clustered(30, 1500)])
    kept = _check(batch, cuda_dev, conf_thres=0.001, iou_thres=0.5, max_det=300)
    ref = onms.non_max_suppression_ref(batch, 0.001, 0.5, max_det=300)
    assert ref[0].shape[0] < 300 and ref[1].shape[0] == 300 and ref[2].shape[0] < 300      # fallback, prefix, fallback
    assert kept == sum(r.shape[0] for r in ref)
    _check(batch, cuda_dev, conf_thres=0.001, iou_thres=0.5, max_det=300, agnostic=True)
    _check(batch[:1], cuda_dev, conf_thres=0.001, iou_thres=0.9, max_det=300)             # many keeps, deep into the list


def test_postprocess_dense_equals_stagewise_chain(cuda_dev):
    """b200yolo_postprocess_dense (class filter -> select-sort -> decode of the ordered prefix -> windowed NMS, with
    the exact fallback) must be bit-identical to decode_and_filter + sort_candidates + nms_sorted and to the oracle,
    for a spread dense head (max_det reached inside the prefix) and for a low max_det / high max_det mix."""
    lv = geometry.level_shapes(640, 640)
    head = synth.synth_head_dense(3, 80, seed=5)
    hd = head.to(cuda_dev)
    pred = ohead.detect_inference_ref(head, lv)
    for conf, iou, max_det in ((0.001, 0.7, 300), (0.001, 0.45, 3000), (0.05, 0.6, 50)):
        ref_out, ref_idx = onms.non_max_suppression_ref(pred, conf, iou, max_det=max_det, return_idxs=True)
        stage = m.nms_candidates(m.decode_and_filter(hd, conf_thres=conf, level_hw=lv), iou, max_det=max_det)
        s_rows, s_anchor, s_count = stage.rows.clone(), stage.anchor.clone(), stage.count.clone()
        cands = m.decode_and_filter(hd, conf_thres=conf, level_hw=lv, defer_boxes=True)
        ws = m.Workspace(3, cands.cap, max_det, cuda_dev)
        det = m.postprocess_dense(cands, ws, hd, level_hw=lv, iou_thres=iou, max_det=max_det)
        assert torch.equal(det.count, s_count)
        for b in range(3):
            k = int(s_count[b])
            assert k == ref_out[b].shape[0]
            assert torch.equal(det.anchor[b, :k], s_anchor[b, :k]) and torch.equal(det.rows[b, :k], s_rows[b, :k])
            assert torch.equal(det.anchor[b, :k].cpu().long(), ref_idx[b])
            assert torch.equal(det.rows[b, :k].cpu()[:, 5], ref_out[b][:, 5])
            assert (det.rows[b, :k].cpu()[:, :5] - ref_out[b][:, :5]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("splits,pipelined", [(2, True), (5, True), (3, False)])
def test_dense_chain_sub_batches_equal_single_stream(cuda_dev, splits, pipelined):
    """api.DenseChain (sub-batches of the images on their own streams; pipelined = filters back to back, each
    sub-batch's tail forked under the next filter) == the single-stream chain, eagerly and as a CUDA graph."""
    lv = geometry.level_shapes(640, 640)
    B = 7
    hd = synth.synth_head_dense(B, 80, seed=11).to(cuda_dev)
    cands = m.decode_and_filter(hd, conf_thres=0.001, level_hw=lv, defer_boxes=True)
    ws = m.Workspace(B, cands.cap, 300, cuda_dev)
    ref = m.postprocess_dense(cands, ws, hd, level_hw=lv, iou_thres=0.7, max_det=300)
    r_rows, r_anchor, r_count = ref.rows.clone(), ref.anchor.clone(), ref.count.clone()
    dc = m.DenseChain(B, cands.cap, 300, cuda_dev, splits=splits, pipelined=pipelined)
    for _ in range(2):
        det = dc(hd, conf_thres=0.001, iou_thres=0.7, level_hw=lv)
        torch.cuda.synchronize()
        assert torch.equal(det.count, r_count) and torch.equal(dc.cand_count, cands.count)
        for b in range(B):
            k = int(r_count[b])
            assert torch.equal(det.anchor[b, :k], r_anchor[b, :k]) and torch.equal(det.rows[b, :k], r_rows[b, :k])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        dc(hd, conf_thres=0.001, iou_thres=0.7, level_hw=lv)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        dc(hd, conf_thres=0.001, iou_thres=0.7, level_hw=lv)
    dc.det.rows.zero_(); dc.det.count.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(dc.det.count, r_count)
    for b in range(B):
        k = int(r_count[b])
        assert torch.equal(dc.det.rows[b, :k], r_rows[b, :k])
