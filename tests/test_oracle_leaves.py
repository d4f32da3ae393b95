"""The numpy restatements that specify the CUDA kernels must equal the real leaf libraries bit for bit."""
import hashlib
import json
import os

import cv2
import numpy as np
import pytest
import torch
import torchvision

from oracle import boxes, head, letterbox, nms, roi


# ---- a1: letterbox geometry table -------------------------------------
@pytest.mark.parametrize("hw,new,auto,exp", [
    ((1200, 1920), 640, False, dict(new_w=640, new_h=400, top=120, bottom=120, left=0, right=0)),
    ((1200, 1920), 640, True, dict(new_w=640, new_h=400, top=8, bottom=8, left=0, right=0)),
    ((900, 1600), 640, False, dict(new_w=640, new_h=360, top=140, bottom=140, left=0, right=0)),
    ((900, 1600), 640, True, dict(new_w=640, new_h=360, top=12, bottom=12, left=0, right=0)),
    ((1130, 930), 640, True, dict(new_w=527, new_h=640, top=0, bottom=0, left=8, right=9)),
    ((543, 770), 1280, True, dict(new_w=1280, new_h=903, top=12, bottom=13, left=0, right=0)),
])
def test_letterbox_geometry(hw, new, auto, exp):
    g = letterbox.letterbox_geometry(hw, (new, new), auto=auto)
    for k, v in exp.items():
        assert g[k] == v, (k, g)


@pytest.mark.parametrize("src,dst", [((900, 1600), (640, 360)), ((1200, 1920), (640, 400)),
                                     ((1130, 930), (527, 640)), ((543, 770), (640, 451)),
                                     ((1194, 1919), (640, 398)), ((1034, 1700), (640, 389)),
                                     ((543, 770), (1280, 903)), ((100, 200), (256, 128)), ((5, 7), (640, 457))])
def test_cv2_resize_restated_bit_exact(src, dst):
    rng = np.random.default_rng(src[0] + dst[0])
    img = rng.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
    ref = cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR)
    got = letterbox.cv2_resize_linear_restated(img, dst)
    assert np.array_equal(ref, got)


def test_letterbox_golden_frames(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "letterbox_golden.json")))
    for key, g in gold.items():
        name, auto = key.split("|auto=")
        im = cv2.imread(os.path.join(golden_dir, "frames", name))
        lb = letterbox.letterbox_ref(im, (640, 640), auto=bool(int(auto)))
        assert list(lb.shape) == g["shape"]
        assert hashlib.sha256(lb.tobytes()).hexdigest() == g["sha256"], key
        assert np.array_equal(lb, letterbox.letterbox_restated(im, (640, 640), auto=bool(int(auto))))


def test_preprocess_ref_layout():
    rng = np.random.default_rng(0)
    f = rng.integers(0, 256, (2, 90, 160, 3), dtype=np.uint8)
    t = letterbox.preprocess_ref(list(f), (64, 64))
    assert t.shape == (2, 3, 64, 64) and t.dtype == torch.float32
    lb = letterbox.letterbox_ref(f[0], (64, 64))
    assert torch.equal(t[0, 0], torch.from_numpy(lb[..., 2].astype(np.float32)) / 255)  # R plane first


# ---- a9: NMS restatement vs the real torchvision kernel --------------------------------------
def test_nms_restated_matches_torchvision_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "nms_golden.npz"))
    for t in range(4):
        b, s, c, thr = z[f"boxes{t}"], z[f"scores{t}"], z[f"cls{t}"], float(z[f"thr{t}"])
        off = (b + (c[:, None] * np.float32(7680)).astype(np.float32)).astype(np.float32)
        keep_tv = torchvision.ops.nms(torch.from_numpy(off), torch.from_numpy(s), thr).numpy()
        assert np.array_equal(keep_tv, z[f"keep{t}"])            # this box's torchvision == golden
        assert np.array_equal(nms.nms_numpy_restated(off, s, thr), keep_tv)


def test_nms_double_threshold_semantics():
    # IoU == float32(0.6) > 0.6 (double) suppresses, but not vs float32(0.6)
    b = torch.tensor([[0., 0., 5., 1.], [0., 0., 3., 1.]])
    s = torch.tensor([0.9, 0.8])
    assert torchvision.ops.nms(b, s, 0.6).tolist() == [0]
    assert torchvision.ops.nms(b, s, float(np.float32(0.6))).tolist() == [0, 1]
    assert nms.nms_numpy_restated(b.numpy(), s.numpy(), 0.6).tolist() == [0]
    assert nms.nms_numpy_restated(b.numpy(), s.numpy(), float(np.float32(0.6))).tolist() == [0, 1]
    # degenerate zero-area duplicates: 0/0 = NaN never suppresses
    z = torch.tensor([[3., 3., 3., 3.], [3., 3., 3., 3.]])
    assert torchvision.ops.nms(z, s, 0.5).tolist() == [0, 1]
    assert nms.nms_numpy_restated(z.numpy(), s.numpy(), 0.5).tolist() == [0, 1]


def test_non_max_suppression_ref_shapes_and_order():
    g = torch.Generator().manual_seed(0)
    pred = torch.rand((2, 4 + 8, 500), generator=g)
    pred[:, :2] *= 600
    pred[:, 2:4] = pred[:, 2:4] * 80 + 4
    out, idx = nms.non_max_suppression_ref(pred, 0.5, 0.45, return_idxs=True, max_det=50)
    for o, i in zip(out, idx):
        assert o.shape[1] == 6 and o.shape[0] == i.shape[0] <= 50
        assert (o[:, 4] > 0.5).all() and (o[:-1, 4] >= o[1:, 4]).all()
    # empty image -> zeros((0,6))
    out = nms.non_max_suppression_ref(torch.zeros((1, 12, 10)), 0.25, 0.45)
    assert out[0].shape == (0, 6)
    # classes filter
    out = nms.non_max_suppression_ref(pred, 0.5, 0.45, classes=[1, 3])
    assert all(set(o[:, 5].tolist()) <= {1.0, 3.0} for o in out)


# ---- a3-a6: decode ----------------------------------------------------------------------------
def test_decode_roundtrip_to_boxes():
    lv = head.level_shapes(640, 640)
    assert sum(h * w for h, w in lv) == 8400
    anchors, strides = head.make_anchors_ref(lv)
    assert anchors.shape == (8400, 2) and anchors[0].tolist() == [0.5, 0.5] and strides[-1].item() == 32
    # a one-hot DFL at bin k decodes to distance k exactly
    x = torch.full((1, 64 + 3, 8400), -100.0)
    for s_, k in enumerate((2, 3, 4, 5)):
        x[0, s_ * 16 + k] = 100.0
    x[0, 64 + 1] = 2.0
    y = head.detect_inference_ref(x, lv)
    assert y.shape == (1, 7, 8400)
    # anchor 0 (stride 8, centre 4,4): x1=4-16=-12, x2=4+32=36, y1=4-24=-20, y2=4+40=44
    assert y[0, :4, 0].tolist() == [12.0, 12.0, 48.0, 64.0]
    assert torch.allclose(y[0, 5, 0], torch.tensor(2.0).sigmoid())


# ---- a10, a12 ---------------------------------------------------------------------------------
def test_scale_boxes_and_safe_crop():
    b = torch.tensor([[100., 150., 300., 400.], [-5., 0., 700., 640.]])
    out = boxes.scale_boxes_ref((640, 640), b, (1200, 1920))
    gain = np.float32(640 / 1920)
    exp0 = (np.float32([100, 150 - 120, 300, 400 - 120]) / gain)
    assert np.array_equal(out[0].numpy(), exp0)
    assert out[1].tolist() == [0.0, 0.0, 1920.0, 1200.0]
    assert boxes.safe_crop_box_ref((1200, 1920), 3, 4, 50, 60, pad=6) == (0, 0, 56, 66)
    assert boxes.safe_crop_box_ref((1200, 1920), 1915, 1190, 1925, 1205) == (1909, 1184, 1920, 1200)
    assert boxes.safe_crop_box_ref((100, 100), 50, 50, 30, 60) is None


# ---- a13: PIL restatement on up- and down-scales ----------------------------------------------
@pytest.mark.parametrize("hw", [(26, 30), (93, 103), (48, 58), (35, 83), (64, 64), (64, 100), (130, 70),
                                (200, 65), (31, 97), (160, 160), (12, 300)])
def test_pil_restated_bit_exact(hw):
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    crop = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    assert torch.equal(roi.classify_preprocess_restated(crop), roi.classify_preprocess_ref(crop))
