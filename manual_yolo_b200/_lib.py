"""ctypes binding of libb200yolo.so (the C ABI in include/b200yolo.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``manual_yolo_b200/csrc/build.sh``.
There is no CPU fallback: a missing library is a hard error.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200YOLO_LIB") or os.path.join(_HERE, "libb200yolo.so")   # env override: tuning experiments only


class Level(Structure):
    """struct b200yolo_level"""
    _fields_ = [("ptr", c_void_p), ("batch_stride", c_int64), ("chan_stride", c_int64),
                ("h", c_int), ("w", c_int), ("stride", c_float)]


# name -> (restype, argtypes); must list every symbol include/b200yolo.h declares
SIGNATURES = {
    "b200yolo_version": (c_int, []),
    "b200yolo_strerror": (c_char_p, [c_int]),
    "b200yolo_letterbox_u8_to_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_int,
                                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200yolo_letterbox_u8_to_f16": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_int,
                                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200yolo_letterbox_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200yolo_decode_filter": (c_int, [POINTER(Level), c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_void_p]),
    "b200yolo_class_filter": (c_int, [POINTER(Level), c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_void_p]),
    "b200yolo_postprocess_small": (c_int, [POINTER(Level), c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                           c_double, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200yolo_postprocess_dense": (c_int, [POINTER(Level), c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                           c_double, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200yolo_filter_decoded": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_int, c_void_p]),
    "b200yolo_sort_topk": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "b200yolo_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_float,
                             c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    "b200yolo_roi_from_detections": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p,
                                             c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200yolo_scale_boxes": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_float,
                                     c_void_p]),
    "b200yolo_roi_crop_resize": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p,
                                         c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200yolo_select_rois": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p]),
    "b200yolo_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b200yolo_stage_rows_h2d": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, c_int, c_int, c_int,
                                        c_void_p, c_int64, c_int64, c_void_p]),
    "b200yolo_letterbox_slices_u8_to_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int64, POINTER(c_int), c_int,
                                                    c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                                    c_int, c_int, c_void_p]),
    "b200yolo_gather_slice_detections": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(c_int), c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b200yolo_greedy_nmm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b200yolo_iou_cost_matrix": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                         c_void_p, c_void_p]),
    "b200yolo_kalman_predict": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b200yolo_kalman_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "b200yolo_kalman_initiate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "b200yolo_selftest_math": (c_int, [c_int, ctypes.c_uint64, c_void_p, c_void_p]),
    "b200yolo_copy2d_h2d": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
}

_lib = None


class B200YoloError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200YoloError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(manual_yolo_b200/csrc/build.sh). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str):
    """C ABI status -> exception (negative: argument error -> ValueError; positive: CUDA error)."""
    if rc == 0:
        return
    msg = load().b200yolo_strerror(rc).decode()
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise B200YoloError(f"{what}: CUDA error {rc}: {msg}")
