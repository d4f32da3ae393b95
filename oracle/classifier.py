"""Oracle: YOLOv8n-cls rank classifier forward.  TEST INFRASTRUCTURE ONLY.

The network is NOT a kernel of this build (it "stays torch"); it exists here so the one known
answer the reference holds for the path -- ``/root/reference/runs/rank_classifier/results.csv:21``
(top-1 0.9403 = 63/67, top-5 0.98507, val loss 0.2352; also ``train_metrics`` inside
``/root/reference/rank_classifier.pt``) -- can pin the ROI leg of the oracle.

``load_checkpoint`` unpickles ``rank_classifier.pt`` with stub classes (ultralytics is not
installed); ``forward`` restates the upstream modules (SURVEY.md Appendix A.12:
``Conv = SiLU(BN(Conv2d))``, ``Bottleneck``, ``C2f``, ``Classify``) functionally from the flat
state dict, so the same code also runs from ``tests/golden/rank_classifier_kat.npz``.
"""

from __future__ import annotations

import pickle
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

NAMES = {0: "10", 1: "2", 2: "3", 3: "4", 4: "5", 5: "6", 6: "7", 7: "8", 8: "9", 9: "A",
         10: "J", 11: "K", 12: "Q"}
BN_EPS = 1e-5  # as pickled in rank_classifier.pt (checked in tests/test_oracle_kat.py)
_STRIDE2 = {"model.0", "model.1", "model.3", "model.5", "model.7"}


class _Stub(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


def _stub_class(modname, clsname):
    mod = sys.modules.get(modname)
    if mod is None:
        mod = types.ModuleType(modname)
        sys.modules[modname] = mod
    if not hasattr(mod, clsname):
        setattr(mod, clsname, type(clsname, (_Stub,), {}))
    return getattr(mod, clsname)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("ultralytics"):
            return _stub_class(module, name)
        return super().find_class(module, name)


class _PickleShim:
    __name__ = "oracle_stub_pickle"
    Unpickler = _Unpickler

    @staticmethod
    def load(f, **kw):
        return _Unpickler(f, **kw).load()


def load_checkpoint(path):
    """Returns (state_dict fp32, names, stored transforms Compose, train_metrics, bn_eps)."""
    ck = torch.load(path, map_location="cpu", pickle_module=_PickleShim, weights_only=False)
    m = ck["model"]
    sd = {k: v.float() for k, v in m.state_dict().items() if v.is_floating_point()}
    eps = {mod.eps for mod in m.modules() if isinstance(mod, nn.BatchNorm2d)}
    assert len(eps) == 1
    return sd, dict(m.names), m.transforms, ck.get("train_metrics"), eps.pop()


def _conv(sd, p, x):
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    x = F.conv2d(x, w, None, stride=2 if p in _STRIDE2 else 1, padding=k // 2)
    x = F.batch_norm(x, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                     sd[p + ".bn.bias"], False, 0.0, BN_EPS)
    return F.silu(x)


def _c2f(sd, p, x):
    y = list(_conv(sd, p + ".cv1", x).chunk(2, 1))
    n = 0
    while f"{p}.m.{n}.cv1.conv.weight" in sd:
        y.append(y[-1] + _conv(sd, f"{p}.m.{n}.cv2", _conv(sd, f"{p}.m.{n}.cv1", y[-1])))
        n += 1
    return _conv(sd, p + ".cv2", torch.cat(y, 1))


@torch.no_grad()
def forward_logits(sd, x: torch.Tensor) -> torch.Tensor:
    """(N,3,64,64) fp32 in [0,1] -> (N,13) logits."""
    x = _conv(sd, "model.0", x)
    x = _conv(sd, "model.1", x)
    x = _c2f(sd, "model.2", x)
    x = _conv(sd, "model.3", x)
    x = _c2f(sd, "model.4", x)
    x = _conv(sd, "model.5", x)
    x = _c2f(sd, "model.6", x)
    x = _conv(sd, "model.7", x)
    x = _c2f(sd, "model.8", x)
    x = _conv(sd, "model.9.conv", x)
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    return F.linear(x, sd["model.9.linear.weight"], sd["model.9.linear.bias"])


def forward(sd, x):
    """Classify head at inference: softmax probabilities."""
    return forward_logits(sd, x).softmax(1)


def state_dict_to_npz_dict(sd):
    return {"w:" + k: v.half().numpy() for k, v in sd.items()}


def state_dict_from_npz(npz, device="cpu"):
    return {k[2:]: torch.from_numpy(np.asarray(npz[k])).float().to(device) for k in npz.files
            if k.startswith("w:")}
