#!/usr/bin/env python
"""Time the class filter (K2, defer_boxes) alone with a cold L2: config 2 (64 x (64+64) x 8400, conf 0.25) and config 3
(256 x (64+80) x 8400, conf 0.001).  B200YOLO_LIB selects the library variant.  Prints one JSON line."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import manual_yolo_b200 as m  # noqa: E402
from manual_yolo_b200 import geometry, synth  # noqa: E402

dev = torch.device("cuda", 0)
lv = geometry.level_shapes(640, 640)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = {"lib": os.environ.get("B200YOLO_LIB", "default")}
for name, B, nc, conf, dense in (("config2", 64, 64, 0.25, False), ("config3", 256, 80, 0.001, True)):
    if dense:
        head = torch.cat([synth.synth_head_dense(64, nc, seed=s) for s in range(B // 64)]).to(dev)
    else:
        head = synth.synth_head_from_labels(B, nc, in_hw=(640, 640), src_hw=(1200, 1920), seed=0, conf_thres=conf)[0].to(dev)
    cands = m.decode_and_filter(head, conf_thres=conf, level_hw=lv, defer_boxes=True, cap=None if dense else 1024)
    ts = []
    for it in range(25):
        cands.count.zero_()
        flush.fill_(it & 1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.decode_and_filter(head, conf_thres=conf, level_hw=lv, defer_boxes=True, out=cands, zero=False, cap=cands.cap)
        b.record()
        torch.cuda.synchronize()
        if it >= 5:
            ts.append(a.elapsed_time(b) * 1e3)
    req = B * nc * 8400 * 4
    out[name] = {"us_median": round(statistics.median(ts), 2), "us_min": round(min(ts), 2),
                 "TBps_required": round(req / statistics.median(ts) / 1e6, 3), "count_sum": int(cands.count.sum())}
    del head, cands
print(json.dumps(out))
