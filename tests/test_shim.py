"""ultralytics_shim: every seam behind the reference's `model(frame)` (detect.py:541, yolo.py:361, pipe.py:179) and
`rank_model(crop)` (detect.py:121) is routed to this package.  Ultralytics cannot be installed here, so the tests inject
a stand-in `ultralytics` package with the upstream module layout and signatures (8.3.176): the CPU tests check the
routing itself (which api function receives which arguments, what raises, what is delegated), the GPU tests run the
routed calls for real and compare them with the oracle."""
import sys
import types
import warnings

import numpy as np
import pytest
import torch

from manual_yolo_b200 import api
from manual_yolo_b200 import ultralytics_shim as shim


def _fake_ultralytics():
    """Stand-in package: same module paths, class and function names, signatures and call order as upstream."""
    calls = []
    mods = {}

    def mod(name):
        m = types.ModuleType(name)
        mods[name] = m
        return m
    ul = mod("ultralytics")
    utils = mod("ultralytics.utils")
    ops = mod("ultralytics.utils.ops")
    nms = mod("ultralytics.utils.nms")
    data = mod("ultralytics.data")
    augment = mod("ultralytics.data.augment")
    engine = mod("ultralytics.engine")
    predictor = mod("ultralytics.engine.predictor")
    nn = mod("ultralytics.nn")
    nnm = mod("ultralytics.nn.modules")
    head = mod("ultralytics.nn.modules.head")
    models = mod("ultralytics.models")
    yolo = mod("ultralytics.models.yolo")
    classify = mod("ultralytics.models.yolo.classify")
    clspred = mod("ultralytics.models.yolo.classify.predict")

    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                            labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, in_place=True,
                            rotated=False, end2end=False, return_idxs=False):
        calls.append(("orig_nms", rotated, multi_label))
        return "ORIG_NMS"

    def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
        calls.append(("orig_scale", xywh, padding))
        return "ORIG_SCALE"
    ops.non_max_suppression = non_max_suppression
    ops.scale_boxes = scale_boxes
    nms.non_max_suppression = non_max_suppression

    class LetterBox:
        def __init__(self, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True, center=True, stride=32,
                     padding_value=114):
            self.new_shape, self.auto, self.scale_fill, self.scaleup = new_shape, auto, scale_fill, scaleup
            self.center, self.stride, self.padding_value = center, stride, padding_value

        def __call__(self, labels=None, image=None):
            calls.append(("orig_letterbox",))
            return "ORIG_LB"
    augment.LetterBox = LetterBox

    class BasePredictor:
        def preprocess(self, im):
            calls.append(("orig_preprocess",))
            return "ORIG_PRE"
    predictor.BasePredictor = BasePredictor

    class Detect:
        def __init__(self, nc=64):
            self.nc, self.reg_max, self.stride, self.export, self.end2end = nc, 16, torch.tensor([8.0, 16.0, 32.0]), False, False

        def _inference(self, x):
            calls.append(("orig_inference",))
            return "ORIG_INF"

        def forward(self, x):                     # upstream predict-mode tail: y = self._inference(x); return (y, x)
            y = self._inference(x)
            return y if self.export else (y, x)

    class Segment(Detect):
        pass
    head.Detect, head.Segment = Detect, Segment

    class ClassificationPredictor(BasePredictor):
        def preprocess(self, img):
            calls.append(("orig_cls_preprocess",))
            return "ORIG_CLS"
    clspred.ClassificationPredictor = ClassificationPredictor
    ul.utils, utils.ops, utils.nms, ul.data, data.augment = utils, ops, nms, data, augment
    return mods, calls


@pytest.fixture
def fake_ul(monkeypatch):
    mods, calls = _fake_ultralytics()
    for name, m in mods.items():
        monkeypatch.setitem(sys.modules, name, m)
    shim._WARNED.clear()
    originals = shim.install()
    yield mods, calls, originals
    shim.uninstall(originals)


class _Cuda:
    """Minimal stand-in for a CUDA tensor on a box without a GPU (the routing only inspects these attributes)."""
    is_cuda, dtype = True, torch.float32

    def __init__(self, shape=(4, 6)):
        self.shape = shape

    def dim(self):
        return len(self.shape)

    def stride(self, i):
        return 1


def test_install_patches_all_six_seams_and_uninstall_restores(fake_ul):
    mods, calls, originals = fake_ul
    assert set(originals) == {
        "ultralytics.utils.ops:non_max_suppression", "ultralytics.utils.ops:scale_boxes",
        "ultralytics.utils.nms:non_max_suppression", "ultralytics.data.augment:LetterBox.__call__",
        "ultralytics.engine.predictor:BasePredictor.preprocess", "ultralytics.nn.modules.head:Detect._inference",
        "ultralytics.models.yolo.classify.predict:ClassificationPredictor.preprocess"}
    ops = mods["ultralytics.utils.ops"]
    patched = ops.non_max_suppression
    shim.uninstall(originals)
    assert ops.non_max_suppression is not patched and ops.non_max_suppression(None) == "ORIG_NMS"
    assert mods["ultralytics.data.augment"].LetterBox()(image=np.zeros((4, 4, 3), np.uint8)) == "ORIG_LB"


def test_nms_and_scale_boxes_route_to_api_with_the_references_call_pattern(fake_ul, monkeypatch):
    mods, calls, _ = fake_ul
    seen = []
    monkeypatch.setattr(api, "non_max_suppression", lambda *a: seen.append(("nms", a)) or "OURS")
    monkeypatch.setattr(api, "scale_boxes", lambda *a: seen.append(("scale", a)) or "OURS_SCALE")
    ops = mods["ultralytics.utils.ops"]
    pred = _Cuda((1, 68, 8400))
    # DetectionPredictor.postprocess: ops.non_max_suppression(preds, conf, iou, classes, agnostic, max_det=..., nc=...)
    assert ops.non_max_suppression((pred, ["levels"]), 0.25, 0.7, None, False, max_det=300, nc=64) == "OURS"
    assert seen[0][0] == "nms" and seen[0][1][0] is pred and seen[0][1][1:3] == (0.25, 0.7) and seen[0][1][7] == 300
    assert mods["ultralytics.utils.nms"].non_max_suppression(pred, 0.35) == "OURS"      # the later 8.3.x location
    # construct_result: ops.scale_boxes(img.shape[2:], pred[:, :4], orig_img.shape)
    assert ops.scale_boxes((384, 640), _Cuda((5, 4)), (900, 1600, 3)) == "OURS_SCALE"
    assert seen[-1][1][0] == (384, 640) and seen[-1][1][2] == (900, 1600, 3)
    assert not any(c[0].startswith("orig") for c in calls)


def test_cpu_tensors_raise_and_foreign_modes_are_delegated_loudly(fake_ul, monkeypatch):
    mods, calls, _ = fake_ul
    ops = mods["ultralytics.utils.ops"]
    with pytest.raises(shim.ShimError):                        # a model running on the CPU: no silent dispatch
        ops.non_max_suppression(torch.zeros((1, 68, 100)), 0.25)
    with pytest.raises(shim.ShimError):
        ops.scale_boxes((640, 640), torch.zeros((3, 4)), (1200, 1920))

    def raising(*a):
        raise NotImplementedError("rotated / end2end / labels are outside the reference's call pattern")
    monkeypatch.setattr(api, "non_max_suppression", raising)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert ops.non_max_suppression(_Cuda((1, 20, 100)), 0.25, rotated=True) == "ORIG_NMS"   # OBB task: upstream's job
        assert ops.non_max_suppression(_Cuda((1, 20, 100)), 0.25, rotated=True) == "ORIG_NMS"
    assert len([x for x in w if "outside the detect path" in str(x.message)]) == 1          # loud, once
    assert calls.count(("orig_nms", True, False)) == 2
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert ops.scale_boxes((640, 640), _Cuda((3, 4)), (1200, 1920), xywh=True) == "ORIG_SCALE"
    assert len(w) == 1


def test_detect_inference_returns_raw_head_consumed_by_nms(fake_ul, monkeypatch):
    mods, calls, _ = fake_ul
    head = mods["ultralytics.nn.modules.head"]
    det = head.Detect(nc=64)
    levels = [_Cuda((2, 128, 80, 80)), _Cuda((2, 128, 40, 40)), _Cuda((2, 128, 20, 20))]
    for lv in levels:
        lv.float = lambda self=lv: self
        lv.device = "cuda:0"
    y, x = det.forward(levels)                                 # predict mode: (y, x)
    assert isinstance(y, shim.RawHead) and y.shape == (2, 68, 8400) and y.nc == 64 and list(y.strides) == [8.0, 16.0, 32.0]
    assert ("orig_inference",) not in calls                    # the dense decode never ran
    got = {}
    monkeypatch.setattr(api, "decode_and_filter", lambda lv, st, conf, classes: got.update(lv=lv, st=st, conf=conf) or "CANDS")

    class Det:
        def to_list(self, return_idxs):
            return [torch.zeros((0, 6))]
    monkeypatch.setattr(api, "nms_candidates", lambda c, iou, ag, md, mn, mw: got.update(c=c, iou=iou, md=md) or Det())
    out = mods["ultralytics.utils.ops"].non_max_suppression((y, x), 0.25, 0.7, max_det=300)
    assert got["lv"] == levels and got["st"] == (8.0, 16.0, 32.0) and got["conf"] == 0.25 and got["iou"] == 0.7
    assert got["c"] == "CANDS" and len(out) == 1 and tuple(out[0].shape) == (0, 6)
    # other heads (segment / pose / OBB) and export mode keep the upstream decode
    assert head.Segment(nc=3)._inference(levels) == "ORIG_INF"
    det.export = True
    assert det._inference(levels) == "ORIG_INF"


def test_preprocess_letterbox_and_classifier_routes(fake_ul, monkeypatch):
    mods, calls, _ = fake_ul
    seen = {}
    monkeypatch.setattr(shim, "_upload", lambda t, dev: ("DEV", tuple(t.shape), str(dev)))
    monkeypatch.setattr(api, "preprocess", lambda d, **kw: seen.update(pre=(d, kw)) or "NET_IN")
    monkeypatch.setattr(api, "classify_preprocess", lambda crops, device: seen.update(cls=(len(crops), str(device))) or torch.zeros((len(crops), 3, 64, 64)))

    class Model:
        stride, pt, fp16 = 32, True, False
    pred = mods["ultralytics.engine.predictor"].BasePredictor()
    pred.device, pred.model, pred.imgsz = torch.device("cuda:0"), Model(), (640, 640)
    pred.args = types.SimpleNamespace(rect=True)
    frame = np.zeros((900, 1600, 3), np.uint8)                 # yolo.py:360-361: one cv2.imread frame
    assert pred.preprocess([frame]) == "NET_IN"
    d, kw = seen["pre"]
    assert d == ("DEV", (1, 900, 1600, 3), "cuda:0") and kw == dict(new_shape=(640, 640), auto=True, stride=32, half=False)
    with pytest.raises(shim.ShimError):                        # mixed shapes cannot be one K1 batch
        pred.preprocess([frame, np.zeros((10, 10, 3), np.uint8)])
    pred.device = torch.device("cpu")
    with pytest.raises(shim.ShimError):
        pred.preprocess([frame])
    assert pred.preprocess(torch.zeros((1, 3, 8, 8))) == "ORIG_PRE"      # tensors: upstream only casts them
    # rank_model(crop) (detect.py:121): ClassificationPredictor.preprocess([crop])
    cp = mods["ultralytics.models.yolo.classify.predict"].ClassificationPredictor()
    cp.device, cp.model, cp.imgsz = torch.device("cuda:0"), Model(), 64
    out = cp.preprocess([np.zeros((60, 45, 3), np.uint8)])
    assert tuple(out.shape) == (1, 3, 64, 64) and seen["cls"] == (1, "cuda:0")
    cp.imgsz = 224
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert cp.preprocess([np.zeros((60, 45, 3), np.uint8)]) == "ORIG_CLS"
    assert len(w) == 1
    # LetterBox()(image=frame): uint8 in, uint8 out, through K1
    monkeypatch.setattr(api, "letterbox", lambda t, **kw: seen.update(lb=(t, kw)) or torch.zeros((384, 640, 3), dtype=torch.uint8))
    lb = mods["ultralytics.data.augment"].LetterBox((640, 640), auto=True, stride=32)
    res = lb(image=frame)
    assert isinstance(res, np.ndarray) and res.shape == (384, 640, 3)
    assert seen["lb"][1] == dict(new_shape=(640, 640), auto=True, scale_fill=False, scaleup=True, center=True, stride=32,
                                 padding_value=114)
    assert lb(labels={"img": frame}) == "ORIG_LB"              # training-time use stays upstream


@pytest.mark.gpu
def test_shim_routes_run_for_real_on_the_gpu(cuda_dev, monkeypatch):
    """The same routes with the real kernels: the patched Ultralytics entry points return the oracle's results."""
    import cv2
    from manual_yolo_b200 import synth
    from oracle import head as ohead
    from oracle import letterbox as olb
    from oracle import nms as onms
    from oracle import roi as oroi
    mods, calls = _fake_ultralytics()
    for name, m_ in mods.items():
        monkeypatch.setitem(sys.modules, name, m_)
    originals = shim.install()
    try:
        ops = mods["ultralytics.utils.ops"]
        # model(frame): preprocess -> (backbone: synthetic head) -> Detect -> postprocess
        frames = synth.synth_frames(1, 900, 1600, seed=3).numpy()

        class Model:
            stride, pt, fp16 = 32, True, False
        pred = mods["ultralytics.engine.predictor"].BasePredictor()
        pred.device, pred.model, pred.imgsz, pred.args = cuda_dev, Model(), (640, 640), types.SimpleNamespace(rect=True)
        net_in = pred.preprocess([frames[0]])
        assert torch.equal(net_in.cpu(), olb.preprocess_ref([frames[0]], (640, 640), auto=True))
        assert np.array_equal(mods["ultralytics.data.augment"].LetterBox((640, 640), auto=True)(image=frames[0]),
                              olb.letterbox_ref(frames[0], (640, 640), auto=True))
        in_hw = tuple(net_in.shape[2:])
        lv = synth.level_shapes(*in_hw)
        head, _ = synth.synth_head_from_labels(1, 64, in_hw=in_hw, src_hw=(900, 1600), seed=3)
        levels, off = [], 0
        for h, w in lv:
            levels.append(head[:, :, off:off + h * w].reshape(1, 128, h, w).contiguous().to(cuda_dev))
            off += h * w
        y, x = mods["ultralytics.nn.modules.head"].Detect(nc=64).forward(levels)
        out, idx = ops.non_max_suppression((y, x), 0.25, 0.7, max_det=300, return_idxs=True)
        ref, ridx = onms.non_max_suppression_ref(ohead.detect_inference_ref(head, lv), 0.25, 0.7, return_idxs=True)
        assert torch.equal(idx[0].cpu(), ridx[0]) and torch.equal(out[0][:, 5].cpu(), ref[0][:, 5])
        assert (out[0][:, :5].cpu() - ref[0][:, :5]).abs().max().item() <= 1e-4
        # the dense-tensor form of the same call, and scale_boxes on its (possibly empty) result
        dense = ohead.detect_inference_ref(head, lv).to(cuda_dev)
        out2 = ops.non_max_suppression(dense, 0.25, 0.7, max_det=300)
        assert torch.equal(out2[0], out[0])
        boxes = out2[0][:, :4].clone()
        from oracle import boxes as oboxes
        assert torch.equal(ops.scale_boxes(in_hw, boxes, (900, 1600, 3)).cpu(), oboxes.scale_boxes_ref(in_hw, ref[0][:, :4], (900, 1600)))
        assert ops.scale_boxes(in_hw, out2[0][:0, :4], (900, 1600, 3)).shape == (0, 4)
        # rank_model(crop): ClassificationPredictor.preprocess([crop]) == the real PIL/torchvision transform
        cp = mods["ultralytics.models.yolo.classify.predict"].ClassificationPredictor()
        cp.device, cp.model, cp.imgsz = cuda_dev, Model(), 64
        rng = np.random.default_rng(0)
        crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(57, 41), (105, 115), (30, 90), (64, 64), (200, 150)]]
        got = cp.preprocess(crops)
        for i, c in enumerate(crops):
            assert torch.equal(got[i].cpu(), oroi.classify_preprocess_ref(c)), i
        assert torch.equal(cp.preprocess([crops[0]])[0], got[0])          # the reference's one-crop-per-call form
    finally:
        shim.uninstall(originals)
