#!/usr/bin/env python
"""Time the dense-regime select-sort (K3, 256 images x 8400 candidates) alone.  B200YOLO_LIB selects the variant."""
import json, os, statistics, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import manual_yolo_b200 as m
from manual_yolo_b200 import geometry, synth
dev = torch.device("cuda", 0)
lv = geometry.level_shapes(640, 640)
B = 256
head = torch.cat([synth.synth_head_dense(64, 80, seed=s) for s in range(B // 64)]).to(dev)
cands = m.decode_and_filter(head, conf_thres=0.001, level_hw=lv)
ws = m.Workspace(B, cands.cap, 300, dev)
ts = []
for it in range(30):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); m.sort_candidates(cands, 30000, ws); b.record()
    torch.cuda.synchronize()
    if it >= 5: ts.append(a.elapsed_time(b) * 1e3)
print(json.dumps({"lib": os.environ.get("B200YOLO_LIB", "default"), "sort_us_median": round(statistics.median(ts), 2), "min": round(min(ts), 2),
                  "order_checksum": int(ws.order[:, :2048].long().sum()) if hasattr(ws, "order") else None}))
