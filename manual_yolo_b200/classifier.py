"""Rank classifier (YOLOv8n-cls) loader + batched forward for the K5 output -- the network the reference loads at
``/root/reference/detect.py:21`` (``rank_model = YOLO("rank_classifier.pt")``) and calls once per crop at
``detect.py:121``.  The CNN itself is not a kernel of this package ("stays torch"): this
module only makes the hand-off runnable without Ultralytics installed --

* ``load_rank_classifier(path)`` reads the Ultralytics checkpoint with a restricted unpickler: every
  ``ultralytics.*`` class is replaced by an inert ``nn.Module`` shell (no Ultralytics code runs); beyond that only
  an explicit allow-list resolves (tensor rebuild helpers, ``torch.nn`` layer classes, torchvision transform classes,
  plain containers) -- any other global, in particular any function, is refused;
* ``RankClassifier.forward_logits`` restates the four upstream module types on plain torch ops (``Conv`` =
  SiLU(BN(Conv2d)), ``Bottleneck``, ``C2f``, ``Classify``; ultralytics==8.3.176 ``nn/modules``), on whatever device the
  ROI batch lives on, fp32 like the reference (``runs/rank_classifier/args.yaml:42`` ``half: false``);
* ``RankClassifier.predict`` returns (top1, top1conf) as ``Results.probs.top1 / top1conf`` (``detect.py:122-124``).

Known answer: 63/67 top-1 on ``rank_classifier/valid`` (``runs/rank_classifier/results.csv:21``), reproduced through
the CUDA K5 output in ``tests/test_gpu_roi.py``.
"""

from __future__ import annotations

import pickle
import sys
import types
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

_STRIDE2 = ("model.0", "model.1", "model.3", "model.5", "model.7")      # ckpt yaml: the five stride-2 Convs
# Globals a YOLOv8 classification checkpoint may reference: tensor / storage rebuild helpers, torch.nn layer classes,
# the torchvision transform objects stored with the model, plain containers.  Everything else is refused -- in
# particular every callable that could run code or touch files (torch.load / hub / cpp_extension, os, subprocess ...).
_ALLOWED_MODULE_PREFIXES = ("torch.nn.modules.", "torchvision.transforms.")
_ALLOWED_GLOBALS = {
    "collections": {"OrderedDict", "defaultdict"},
    "builtins": {"set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "str", "bytes", "complex", "slice", "range"},
    "__builtin__": {"set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "str", "bytes", "complex", "slice", "range"},
    "torch": {"Size", "device", "dtype", "Tensor", "FloatStorage", "HalfStorage", "BFloat16Storage", "DoubleStorage",
              "LongStorage", "IntStorage", "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage", "float16", "float32",
              "float64", "bfloat16", "int64", "int32", "int16", "int8", "uint8", "bool"},
    "torch._utils": {"_rebuild_tensor_v2", "_rebuild_parameter", "_rebuild_parameter_with_state", "_rebuild_tensor"},
    "torch.storage": {"TypedStorage", "UntypedStorage", "_load_from_bytes"},
    "torch.nn.parameter": {"Parameter", "Buffer"},
    "numpy": {"dtype", "ndarray"},
    "numpy.core.multiarray": {"_reconstruct", "scalar"},
    "numpy._core.multiarray": {"_reconstruct", "scalar"},
    "_codecs": {"encode"},
    "pathlib": {"PosixPath", "PurePosixPath", "WindowsPath", "PureWindowsPath", "Path"},
    "datetime": {"datetime", "date", "timedelta"},
    "copyreg": {"_reconstructor"},
}


class _Shell(nn.Module):
    """Inert stand-in for an ultralytics module class: holds the pickled attributes / parameters, runs no code."""

    def __init__(self, *a, **k):
        super().__init__()


def _shell_class(modname: str, clsname: str):
    mod = sys.modules.get(modname)
    if mod is None:
        mod = types.ModuleType(modname)
        mod.__dict__["__b200_shell__"] = True
        sys.modules[modname] = mod
    if not hasattr(mod, clsname):
        setattr(mod, clsname, type(clsname, (_Shell,), {"__module__": modname}))
    return getattr(mod, clsname)


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "ultralytics":
            return _shell_class(module, name)
        if name in _ALLOWED_GLOBALS.get(module, ()) and module != "torch.storage" or \
                (module == "torch.storage" and name in ("TypedStorage", "UntypedStorage")):
            return super().find_class(module, name)
        if module.startswith(_ALLOWED_MODULE_PREFIXES) and "." not in name and not name.startswith("_"):
            obj = super().find_class(module, name)
            if isinstance(obj, type):                      # layer / transform / enum CLASSES only, never functions
                return obj
        raise pickle.UnpicklingError(f"rank classifier checkpoint references {module}.{name}: refused")


class _PickleModule:
    __name__ = "manual_yolo_b200_restricted_pickle"
    Unpickler = _RestrictedUnpickler

    @staticmethod
    def load(f, **kw):
        return _RestrictedUnpickler(f, **kw).load()


class RankClassifier:
    """Flat fp32 state dict of the YOLOv8n-cls checkpoint + functional forward."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], names: Dict[int, str], bn_eps: float = 1e-5, imgsz: int = 64):
        self.sd = {k: v.float() for k, v in state_dict.items()}
        self.names = dict(names)
        self.bn_eps = float(bn_eps)
        self.imgsz = int(imgsz)

    def to(self, device) -> "RankClassifier":
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    # ---- upstream module types on plain torch ops ----
    def _conv(self, p, x):
        w = self.sd[p + ".conv.weight"]
        x = F.conv2d(x, w, None, stride=2 if p in _STRIDE2 else 1, padding=w.shape[-1] // 2)
        x = F.batch_norm(x, self.sd[p + ".bn.running_mean"], self.sd[p + ".bn.running_var"], self.sd[p + ".bn.weight"],
                         self.sd[p + ".bn.bias"], False, 0.0, self.bn_eps)
        return F.silu(x)

    def _c2f(self, p, x):
        y = list(self._conv(p + ".cv1", x).chunk(2, 1))
        n = 0
        while f"{p}.m.{n}.cv1.conv.weight" in self.sd:          # Bottleneck(shortcut=True): x + cv2(cv1(x))
            y.append(y[-1] + self._conv(f"{p}.m.{n}.cv2", self._conv(f"{p}.m.{n}.cv1", y[-1])))
            n += 1
        return self._conv(p + ".cv2", torch.cat(y, 1))

    @torch.no_grad()
    def forward_logits(self, x: torch.Tensor) -> torch.Tensor:
        """(N,3,64,64) fp32 RGB in [0,1] (the K5 batch) -> (N, n_classes) logits."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an (N,3,S,S) ROI batch")
        if next(iter(self.sd.values())).device != x.device:
            self.to(x.device)
        x = x.float()
        x = self._conv("model.0", x)
        x = self._conv("model.1", x)
        x = self._c2f("model.2", x)
        x = self._conv("model.3", x)
        x = self._c2f("model.4", x)
        x = self._conv("model.5", x)
        x = self._c2f("model.6", x)
        x = self._conv("model.7", x)
        x = self._c2f("model.8", x)
        x = self._conv("model.9.conv", x)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return F.linear(x, self.sd["model.9.linear.weight"], self.sd["model.9.linear.bias"])

    __call__ = forward_logits

    @torch.no_grad()
    def predict(self, rois: torch.Tensor):
        """(top1 (N,) int64, top1conf (N,) fp32): ``Results.probs.top1`` / ``top1conf`` of ``detect.py:122-124``."""
        probs = self.forward_logits(rois).softmax(1)
        conf, top1 = probs.max(1)
        return top1, conf


def load_rank_classifier(path: str, device: Optional[str] = None) -> RankClassifier:
    """Load ``rank_classifier.pt`` (``detect.py:21``) without Ultralytics.  The checkpoint stores the model object in
    fp16; weights are converted to fp32 (the reference predicts with ``half=False``)."""
    ck = torch.load(path, map_location="cpu", pickle_module=_PickleModule, weights_only=False)
    model = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    if not isinstance(model, nn.Module):
        raise ValueError(f"{path}: no model object in the checkpoint")
    sd = {k: v.float() for k, v in model.state_dict().items() if v.is_floating_point()}
    eps = {m.eps for m in model.modules() if isinstance(m, nn.BatchNorm2d)} or {1e-5}
    if len(eps) != 1:
        raise ValueError("mixed BatchNorm eps values in the checkpoint")
    names = dict(getattr(model, "names", {}) or {})
    if not names:
        raise ValueError(f"{path}: the checkpoint carries no class names")
    args = getattr(model, "args", None)
    imgsz = (args.get("imgsz") if isinstance(args, dict) else getattr(args, "imgsz", None)) or 64
    clf = RankClassifier(sd, names, eps.pop(), int(imgsz))
    return clf.to(device) if device is not None else clf


def rank_classifier_from_arrays(arrays: Dict[str, "object"], names: Dict[int, str], bn_eps: float = 1e-5,
                                device: Optional[str] = None) -> RankClassifier:
    """Build from a plain {parameter name: array} mapping (e.g. an .npz export of the weights)."""
    sd = {k: torch.as_tensor(v).float() for k, v in arrays.items()}
    clf = RankClassifier(sd, names, bn_eps)
    return clf.to(device) if device is not None else clf
