"""Drop-in for an environment that HAS Ultralytics installed: ``install()`` routes the six seams behind the
reference's two calls -- ``model(frame)`` (``/root/reference/detect.py:541``, ``yolo.py:361``, ``pipe.py:179``) and
``rank_model(crop)`` (``detect.py:121``) -- through libb200yolo.so by replacing the upstream functions in place:

    import manual_yolo_b200.ultralytics_shim as shim
    shim.install()          # before YOLO(...) is constructed (detect.py:20)

| upstream symbol (ultralytics==8.3.176, requirements.txt:95)          | routed to                                   |
|----------------------------------------------------------------------|---------------------------------------------|
| ``data.augment.LetterBox.__call__(labels=None, image=None)``         | ``api.letterbox`` (K1, uint8 form)          |
| ``engine.predictor.BasePredictor.preprocess(im)``                    | ``api.preprocess`` (K1 fused: letterbox + RGB + CHW + /255) |
| ``nn.modules.head.Detect._inference(x)``                             | returns a ``RawHead`` (no dense decode): the per-level tensors go straight to K2 |
| ``utils.ops.non_max_suppression`` (``utils.nms`` in later 8.3.x)     | ``api.non_max_suppression`` / ``nms_from_head`` (K2-K4) |
| ``utils.ops.scale_boxes``                                            | ``api.scale_boxes``                         |
| ``models.yolo.classify.predict.ClassificationPredictor.preprocess``  | ``api.classify_preprocess`` (K5)            |

No CPU path and no silent dispatch: a CPU prediction tensor (a model running on the CPU) RAISES -- uninstall the shim
or move the model to the GPU.  Modes outside the reference's call pattern (``rotated``, ``end2end``, ``multi_label``,
``labels``, mask channels: other Ultralytics tasks share the function) are handed to the original with ONE loud
warning per mode.  Ultralytics is not installable in the build container: ``tests/test_shim.py`` exercises every route
against a stand-in package with the upstream signatures.
"""

from __future__ import annotations

import functools
import importlib
import warnings
from dataclasses import dataclass
from typing import List, Sequence

_WARNED = set()


class ShimError(RuntimeError):
    pass


@dataclass
class RawHead:
    """What the patched ``Detect._inference`` returns instead of the dense (B, 4+nc, A) tensor: the per-level
    (B, 64+nc, Hi, Wi) tensors as the Detect convolutions produced them.  The patched ``non_max_suppression`` consumes
    it directly (class filter -> DFL decode of the survivors -> sort -> NMS), so the decoded tensor never exists."""
    levels: List["object"]
    strides: Sequence[float]
    nc: int

    @property
    def shape(self):                      # (B, 4+nc, A), as upstream code that only inspects the shape expects
        b = self.levels[0].shape[0]
        return (b, 4 + self.nc, sum(int(x.shape[2]) * int(x.shape[3]) for x in self.levels))

    @property
    def device(self):
        return self.levels[0].device


def _warn_once(key, msg):
    if key not in _WARNED:
        _WARNED.add(key)
        warnings.warn(f"manual_yolo_b200.ultralytics_shim: {msg}", RuntimeWarning, stacklevel=3)


def _upload(host_tensor, device):
    """Pinned host tensor -> device (the `.to(device)` of BasePredictor.preprocess); a seam of its own so that the
    routing can be exercised without a GPU."""
    return host_tensor.pin_memory().to(device, non_blocking=True)


def _require_cuda(t, what):
    if not getattr(t, "is_cuda", False):
        raise ShimError(f"{what} is not on a CUDA device: manual_yolo_b200 has no CPU path (move the model to the GPU or "
                        "call ultralytics_shim.uninstall())")


def nms_from_head(raw: RawHead, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, max_det=300,
                  max_nms=30000, max_wh=7680, return_idxs=False):
    """``ops.non_max_suppression`` on a ``RawHead``: K2 (class filter + DFL decode of the survivors) -> K3 -> K4.
    Returns the same ragged ``list[Tensor(n_i, 6)]`` (letterboxed pixels)."""
    from . import api
    cands = api.decode_and_filter(list(raw.levels), tuple(float(s) for s in raw.strides), conf_thres, classes)
    det = api.nms_candidates(cands, iou_thres, agnostic, max_det, max_nms, max_wh)
    out = det.to_list(return_idxs)
    if return_idxs:
        return [o.clone() for o in out[0]], out[1]
    return [o.clone() for o in out]


def install(patch_scale_boxes: bool = True) -> dict:
    """Patch Ultralytics in place; returns {qualified name: original} for ``uninstall``."""
    try:
        import ultralytics  # noqa: F401
    except ImportError as e:
        raise ImportError("manual_yolo_b200.ultralytics_shim needs `ultralytics` to be installed") from e
    import numpy as np
    import torch

    from . import api

    originals = {}

    def _patch(modname, attr, make, owner=None):
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            return
        target = getattr(mod, owner) if owner else mod
        if target is None or not hasattr(target, attr):
            return
        orig = getattr(target, attr)
        originals[f"{modname}:{owner + '.' if owner else ''}{attr}"] = orig
        setattr(target, attr, make(orig))

    # ---- K2-K4: ops.non_max_suppression ------------------------------------------------------------------------
    def make_nms(orig):
        @functools.wraps(orig)
        def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                                multi_label=False, labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000,
                                max_wh=7680, in_place=True, rotated=False, end2end=False, return_idxs=False, **kw):
            p = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
            if isinstance(p, RawHead):
                if rotated or end2end or labels or multi_label or kw:
                    raise ShimError("the RawHead route covers the detect task's default call pattern only")
                _require_cuda(p.levels[0], "the Detect head output")
                return nms_from_head(p, conf_thres, iou_thres, classes, agnostic, max_det, max_nms, max_wh, return_idxs)
            _require_cuda(p, "prediction")
            try:
                return api.non_max_suppression(p, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nc,
                                               max_time_img, max_nms, max_wh, in_place, rotated, end2end, return_idxs)
            except NotImplementedError as e:     # another task's mode (OBB / segment / multi-label validation)
                _warn_once(("nms", str(e)), f"non_max_suppression mode outside the detect path ({e}); using the original "
                                            "Ultralytics implementation for such calls")
                return orig(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nc,
                            max_time_img, max_nms, max_wh, in_place, rotated, end2end, return_idxs, **kw)
        return non_max_suppression

    # ---- a10: ops.scale_boxes ----------------------------------------------------------------------------------
    def make_scale(orig):
        @functools.wraps(orig)
        def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
            _require_cuda(boxes, "boxes")
            if xywh or not padding or boxes.dim() != 2 or boxes.dtype != torch.float32:
                _warn_once(("scale", xywh, padding), "scale_boxes variant outside the detect path; using the original")
                return orig(img1_shape, boxes, img0_shape, ratio_pad, padding, xywh)
            if boxes.stride(-1) != 1:
                boxes = boxes.contiguous()
            return api.scale_boxes(img1_shape, boxes, img0_shape, ratio_pad, padding, xywh)
        return scale_boxes

    for modname in ("ultralytics.utils.ops", "ultralytics.utils.nms"):
        _patch(modname, "non_max_suppression", make_nms)
        if patch_scale_boxes:
            _patch(modname, "scale_boxes", make_scale)

    # ---- K1 (uint8 form): LetterBox.__call__(labels=None, image=None) --------------------------------------------
    def make_letterbox(orig):
        @functools.wraps(orig)
        def __call__(self, labels=None, image=None):
            if labels:                                   # training-time use (labels are transformed too): not this path
                return orig(self, labels, image)
            img = image
            if img is None:
                raise ShimError("LetterBox.__call__ needs image= (inference use)")
            kw = dict(new_shape=self.new_shape, auto=self.auto, scale_fill=getattr(self, "scale_fill", getattr(self, "scaleFill", False)),
                      scaleup=self.scaleup, center=getattr(self, "center", True), stride=self.stride,
                      padding_value=getattr(self, "padding_value", 114))
            if isinstance(img, np.ndarray):              # the reference hands host frames (detect.py:541): up, K1, down
                t = _upload(torch.from_numpy(np.ascontiguousarray(img)), "cuda")
                return api.letterbox(t, **kw).cpu().numpy()
            _require_cuda(img, "image")
            return api.letterbox(img, **kw)
        return __call__

    _patch("ultralytics.data.augment", "__call__", make_letterbox, owner="LetterBox")

    # ---- K1 (fused form): BasePredictor.preprocess(im) ----------------------------------------------------------
    def make_preprocess(orig):
        @functools.wraps(orig)
        def preprocess(self, im):
            if isinstance(im, torch.Tensor):             # already a tensor: upstream only moves / casts / scales it
                return orig(self, im)
            frames = list(im)
            shapes = {f.shape for f in frames}
            if len(shapes) != 1 or frames[0].ndim != 3 or frames[0].shape[2] != 3 or frames[0].dtype != np.uint8:
                raise ShimError("BasePredictor.preprocess: the K1 route needs same-shape HxWx3 uint8 frames")
            dev = getattr(self, "device", None)
            if dev is None or torch.device(dev).type != "cuda":
                raise ShimError("BasePredictor.preprocess: the predictor's device is not CUDA (no CPU path)")
            d = _upload(torch.from_numpy(np.ascontiguousarray(np.stack(frames))), dev)
            model = self.model
            rect = bool(getattr(self.args, "rect", True))
            auto = rect and bool(getattr(model, "pt", True))      # pre_transform: same_shapes and args.rect and model.pt
            stride = int(getattr(model, "stride", 32))
            return api.preprocess(d, new_shape=tuple(self.imgsz) if hasattr(self.imgsz, "__len__") else (self.imgsz, self.imgsz),
                                  auto=auto, stride=stride, half=bool(getattr(model, "fp16", False)))
        return preprocess

    _patch("ultralytics.engine.predictor", "preprocess", make_preprocess, owner="BasePredictor")

    # ---- K2 entry: Detect._inference(x) returns the raw levels ---------------------------------------------------
    def make_inference(orig):
        @functools.wraps(orig)
        def _inference(self, x):
            if getattr(self, "export", False) or getattr(self, "end2end", False) or int(getattr(self, "reg_max", 16)) != 16:
                return orig(self, x)
            _require_cuda(x[0], "the Detect head input")
            if type(self).__name__ != "Detect":          # Segment / Pose / OBB heads extend _inference's output
                return orig(self, x)
            return RawHead(levels=[xi.float() if xi.dtype != torch.float32 else xi for xi in x],
                           strides=[float(s) for s in self.stride], nc=int(self.nc))
        return _inference

    _patch("ultralytics.nn.modules.head", "_inference", make_inference, owner="Detect")

    # ---- K5: ClassificationPredictor.preprocess(img) -----------------------------------------------------------
    def make_cls_preprocess(orig):
        @functools.wraps(orig)
        def preprocess(self, img):
            if isinstance(img, torch.Tensor):
                return orig(self, img)
            size = self.imgsz[0] if hasattr(self.imgsz, "__len__") else self.imgsz
            if int(size) != 64:
                _warn_once(("cls", size), f"classifier imgsz {size} != 64: K5 produces 64x64 only; using the original preprocess")
                return orig(self, img)
            dev = getattr(self, "device", None)
            if dev is None or torch.device(dev).type != "cuda":
                raise ShimError("ClassificationPredictor.preprocess: the predictor's device is not CUDA (no CPU path)")
            out = api.classify_preprocess(list(img), device=dev)
            return out.half() if bool(getattr(self.model, "fp16", False)) else out
        return preprocess

    _patch("ultralytics.models.yolo.classify.predict", "preprocess", make_cls_preprocess, owner="ClassificationPredictor")
    return originals


def uninstall(originals: dict) -> None:
    for qual, fn in originals.items():
        modname, attr = qual.split(":")
        mod = importlib.import_module(modname)
        if "." in attr:
            owner, name = attr.split(".")
            setattr(getattr(mod, owner), name, fn)
        else:
            setattr(mod, attr, fn)
