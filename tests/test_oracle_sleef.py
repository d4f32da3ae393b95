"""The oracle's C restatement of torch's CPU expf (Sleef expf_u10, FMA) is bit-identical to what
torch.sigmoid / torch.softmax produce on this host -- the CUDA expf_torch() follows the same sequence."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libexpf_torch.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "c")], check=True)
    return ctypes.CDLL(SO)


def _call(fn, x):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    fn(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
    return y


def test_sigmoid_bitwise(lib):
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.randn(1_000_000, generator=g) * 5, torch.linspace(-30, 30, 100_001),
                   torch.tensor([0.0, -0.0, 88.0, -88.0, 16.6, 17.0, -103.0])]).float()
    got, ref = _call(lib.oracle_sigmoid_torch, x.numpy()), x.sigmoid().numpy()
    # torch evaluates the last few elements of every per-thread chunk with scalar libm expf instead of
    # the vectorised Sleef kernel, so a handful of positions (not values) may differ by one ulp.
    diff = np.abs(got.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
    assert diff.max() <= 2
    assert (diff != 0).sum() <= 4 * torch.get_num_threads() + 4


def test_dfl_softmax_bitwise(lib):
    g = torch.Generator().manual_seed(1)
    B, A = 2, 8400
    box = torch.randn((B, 64, A), generator=g) * 2
    v = box.view(B, 4, 16, A).transpose(2, 1)
    ref = v.softmax(1).numpy()
    xn = v.numpy()
    e = _call(lib.oracle_expf_torch, (xn - xn.max(1, keepdims=True)).astype(np.float32)).reshape(xn.shape)
    s = np.zeros((B, 1, 4, A), np.float32)
    for i in range(16):
        s = (s + e[:, i:i + 1]).astype(np.float32)
    assert np.array_equal((e / s).astype(np.float32), ref)
