"""The C-ABI shared library loads and exports every symbol include/b200yolo.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from manual_yolo_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200yolo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200yolo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200yolo.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_strerror():
    lib = _lib.load()
    assert lib.b200yolo_version() == 100
    assert lib.b200yolo_strerror(0) == b"ok"
    assert b"NULL" in lib.b200yolo_strerror(-1)
    assert b"workspace" in lib.b200yolo_strerror(-5)


def test_argument_errors_before_any_launch():
    """Host-side validation returns negative codes without touching a device."""
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    assert lib.b200yolo_letterbox_u8_to_f32(null, 1, 8, 8, 24, 192, null, 8, 8, 8, 8, 0, 0, 114, 1, null) == -1
    one = ctypes.c_void_p(16)
    assert lib.b200yolo_letterbox_u8_to_f32(one, 1, 8, 8, 24, 192, one, 8, 8, 9, 8, 0, 0, 114, 1, null) == -2
    assert lib.b200yolo_letterbox_u8_to_f32(one, 1, 8, 8, 24, 192, one, 8, 8, 8, 8, 0, 0, 300, 1, null) == -6
    assert lib.b200yolo_filter_decoded(one, 1, 10, 6, 100, 1.5, null, one, one, one, 100, null) == -6
    assert lib.b200yolo_filter_decoded(one, 1, 8, 6, 100, 0.5, null, one, one, one, 100, null) == -2
    assert lib.b200yolo_nms(one, one, one, one, 1, 100, 30000, 2.0, 7680.0, 0, 300, null, one, one, one, null, 0, null, null, 0, null) == -6
    assert lib.b200yolo_nms(one, one, one, one, 1, 100, 30000, 0.5, 7680.0, 0, 300, null, one, one, one, null, 0, one, null, 0, null) == -1
    assert lib.b200yolo_sort_topk(one, one, one, 1, 70000, 30000, one, null, 0, null) == -4
    assert lib.b200yolo_sort_topk(one, one, one, 1, 20000, 30000, one, null, 0, null) == -1   # needs workspace
    assert lib.b200yolo_roi_crop_resize(one, 1, 8, 8, 24, 192, one, one, null, 4, 6, 128, one, one, null) == -4
    assert lib.b200yolo_postprocess_small(None, 0, one, one, one, 1, 2048, 30000, 0.5, 7680.0, 0, 300, null, one, one, one,
                                          null, 0, null, null, null) == -4
    hdr = 112                                                    # 7 ints per image (sorted count, fallback flag, 2 thresholds, decoded count), 16-B padded
    assert lib.b200yolo_workspace_bytes(4, 8400) == hdr + 16
    assert lib.b200yolo_workspace_bytes(4, 20000) == hdr + 4 * (20000 + 1250) * 16
    with pytest.raises(ValueError):
        _lib.check(-2, "x")


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.B200YoloError):
        _lib.load()


def test_cpu_tensors_rejected():
    import torch
    import manual_yolo_b200 as m
    with pytest.raises(ValueError):
        m.preprocess(torch.zeros((1, 8, 8, 3), dtype=torch.uint8))
    with pytest.raises(ValueError):
        m.non_max_suppression(torch.zeros((1, 10, 20)))
    with pytest.raises(NotImplementedError):
        m.non_max_suppression(torch.zeros((1, 10, 20)), rotated=True)


def test_header_is_valid_c_and_matches_bindings(tmp_path):
    """include/b200yolo.h must compile as plain C (it is the drop-in boundary for non-Python hosts) and every function it
    declares must be bound in _lib.SIGNATURES with the same number of parameters."""
    import re
    import subprocess
    hdr = os.path.join(ROOT, "include", "b200yolo.h")
    src = tmp_path / "use.c"
    src.write_text('#include "b200yolo.h"\nint (*entry)(void) = b200yolo_version;\nint main(void) { return entry == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.dirname(hdr), str(src)],
                   check=True)
    text = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)
    decls = re.findall(r"\b(b200yolo_\w+)\s*\(([^;{}]*?)\)\s*;", text)
    names = {n for n, _ in decls}
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    for name, params in decls:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))
