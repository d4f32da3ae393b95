"""Optional drop-in for an environment that HAS Ultralytics installed (this container does not, so this
module is import-guarded and not exercised by the tests here -- the functions it installs are, against the
oracle).  ``install()`` routes the reference's ``model(frame)`` post-processing (``/root/reference/detect.py:541``,
``yolo.py:361``, ``pipe.py:179``) through libb200yolo.so by replacing the upstream functions in place:

    import manual_yolo_b200.ultralytics_shim as shim
    shim.install()          # before YOLO(...) is constructed (detect.py:20)

``ops.non_max_suppression`` lives in ``ultralytics/utils/ops.py`` in 8.3.176 (the release the reference pins,
``requirements.txt:95``) and moved to ``ultralytics/utils/nms.py`` in later 8.3.x releases: both are patched
when present.  CPU tensors keep going to the original functions (this package has no CPU path).
"""

from __future__ import annotations

import functools


def install(patch_scale_boxes: bool = True) -> dict:
    """Patch Ultralytics in place; returns {qualified name: original function} for ``uninstall``."""
    try:
        import ultralytics  # noqa: F401
    except ImportError as e:  # pragma: no cover - ultralytics is absent in this image
        raise ImportError("manual_yolo_b200.ultralytics_shim needs `ultralytics` to be installed") from e
    import importlib

    from . import api

    originals = {}

    def _route(orig, ours, tensor_arg):
        @functools.wraps(orig)
        def wrapper(*args, **kwargs):
            t = args[tensor_arg] if len(args) > tensor_arg else None
            if isinstance(t, (list, tuple)):
                t = t[0]
            if getattr(t, "is_cuda", False):
                try:
                    return ours(*args, **kwargs)
                except NotImplementedError:          # e.g. rotated / multi_label: upstream handles it
                    return orig(*args, **kwargs)
            return orig(*args, **kwargs)
        return wrapper

    for modname in ("ultralytics.utils.ops", "ultralytics.utils.nms"):
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        if hasattr(mod, "non_max_suppression"):
            originals[f"{modname}.non_max_suppression"] = mod.non_max_suppression
            mod.non_max_suppression = _route(mod.non_max_suppression, api.non_max_suppression, 0)
        if patch_scale_boxes and hasattr(mod, "scale_boxes"):
            originals[f"{modname}.scale_boxes"] = mod.scale_boxes
            mod.scale_boxes = _route(mod.scale_boxes, api.scale_boxes, 1)
    return originals


def uninstall(originals: dict) -> None:
    import importlib
    for qual, fn in originals.items():
        modname, attr = qual.rsplit(".", 1)
        setattr(importlib.import_module(modname), attr, fn)
