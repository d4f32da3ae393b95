#!/usr/bin/env python
"""Time K5 (b200yolo_roi_crop_resize: fast launch + general-body launch) on N synthetic rank-card ROIs, cold L2.
B200YOLO_LIB selects the library variant.  Prints one JSON line per N."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import manual_yolo_b200 as m  # noqa: E402
from manual_yolo_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
B = 64
frames = synth.synth_frames(B, 1200, 1920, seed=0).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for N in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096").split(",")]:
    boxes, bidx = synth.synth_rois(N, B, seed=0)
    boxes, bidx = boxes.to(dev), bidx.to(dev)
    dst = torch.empty((N, 3, 64, 64), dtype=torch.float32, device=dev)
    valid = torch.empty((N,), dtype=torch.int32, device=dev)
    ts = []
    for it in range(16):
        flush.fill_(it & 1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.crop_resize_rois(frames, boxes, bidx, pad=6, out=dst, valid=valid)
        b.record()
        torch.cuda.synchronize()
        if it >= 4:
            ts.append(a.elapsed_time(b) * 1e3)
    print(json.dumps({"lib": os.environ.get("B200YOLO_LIB", "default"), "N": N, "us_median": round(statistics.median(ts), 2),
                      "us_min": round(min(ts), 2), "ns_per_roi": round(1e3 * statistics.median(ts) / N, 2),
                      "checksum": float(dst.double().sum())}))
    del dst, valid, boxes, bidx
