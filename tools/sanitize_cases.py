#!/usr/bin/env python
"""Small-shape invocations of every kernel with hand-rolled synchronisation, for compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_cases.py

Covered: K1's TMA bulk-copy ring + its cooperative fallback (unaligned rows), K5's per-warp bulk-copy rings, its
guarded path (crops at the first / last bytes of the buffer), vertical tiling and the general split body, the fused
per-image post-processing kernel, the select-sort + windowed NMS of the dense chain (incl. on-demand decode of a second
window) and the slice gather.  Shapes are tiny: the sanitizer slows kernels by 10-100x.  One tool per gpurun call
(see tools/sanitize.sh); logs are committed under profiles/.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import manual_yolo_b200 as m  # noqa: E402
from manual_yolo_b200 import synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    # K1: bulk-copy ring (pitch % 16 == 0), cooperative fallback (odd width), slice mode
    for hw in [(300, 480), (211, 333)]:
        f = synth.synth_frames(2, *hw, seed=1).to(dev)
        m.preprocess(f, (160, 160))
        m.preprocess(f, (160, 160), auto=True, half=True)
        m.letterbox(f, (160, 160))
    f = synth.synth_frames(1, 256, 384, seed=2).to(dev)
    m.preprocess_slices(f, m.geometry.slice_boxes(256, 384, 128, 128, 0.2, 0.2), (128, 128))
    # K5 list form: interior crops, crops on every border and at the first / last bytes, up- and down-scales,
    # tall crops (vertical tiles), one beyond the fast envelope (general split body), invalid ones
    H, W = 420, 640
    frames = synth.synth_frames(2, H, W, seed=3).to(dev)
    boxes = [[10.3, 12.2, 60.9, 70.1], [0., 0., 40., 30.], [W - 30., H - 25., W + 5., H + 5.], [100., 5., 130., 400.],
             [5., 100., 630., 140.], [200., 200., 330., 390.], [0., 0., W, H], [50., 50., 40., 90.], [300.5, 10.5, 364.5, 74.5],
             [W - 50., 0., W - 1., 45.], [0., H - 40., 35., H - 1.]]
    for _ in range(12):
        x1, y1 = float(torch.rand(1, generator=g)) * (W - 120), float(torch.rand(1, generator=g)) * (H - 120)
        boxes.append([x1, y1, x1 + 20 + float(torch.rand(1, generator=g)) * 100, y1 + 20 + float(torch.rand(1, generator=g)) * 100])
    bx = torch.tensor(boxes, dtype=torch.float32, device=dev)
    bidx = (torch.arange(len(boxes), dtype=torch.int32) % 2).to(dev)
    out, valid = m.crop_resize_rois(frames, bx, bidx, pad=6)
    assert set(valid.cpu().tolist()) <= {0, 1, 2}
    odd = synth.synth_frames(1, 97, 211, seed=4).to(dev)             # pitch 633: per-row 16-byte phases
    m.crop_resize_rois(odd, torch.tensor([[3., 3., 80., 90.], [0., 0., 211., 97.]], device=dev), torch.zeros(2, dtype=torch.int32, device=dev))
    # fused sparse-regime pipeline (class filter -> postprocess_small -> roi_det) eager and as a graph
    pipe = m.Pipeline(2, (300, 480), 64, imgsz=160, conf=0.25, iou=0.45, device=dev, cap=256)
    head, _ = synth.synth_head_from_labels(2, 64, in_hw=pipe.in_hw, src_hw=(300, 480), seed=5)
    fr = synth.synth_frames(2, 300, 480, seed=5).to(dev)
    pipe(fr, head.to(dev))
    pipe.check_overflow()
    # dense chain: select-sort + pre-decode + windowed NMS; max_det large enough to force a second (on-demand) window
    lv = m.geometry.level_shapes(320, 320)
    dense = synth.synth_head_dense(2, 16, in_hw=(320, 320), seed=6).to(dev)
    cands = m.decode_and_filter(dense, conf_thres=0.001, level_hw=lv, defer_boxes=True)
    ws = m.Workspace(2, cands.cap, 700, dev)
    m.postprocess_dense(cands, ws, dense, level_hw=lv, iou_thres=0.7, max_det=700)
    m.DenseChain(2, cands.cap, 300, dev, splits=2)(dense, conf_thres=0.001, iou_thres=0.45, level_hw=lv)
    # stage-wise kernels + sliced prediction (gather)
    full = m.decode_and_filter(dense, conf_thres=0.001, level_hw=lv)
    m.nms_candidates(full, 0.7)
    sp = m.SlicedPipeline(1, (256, 384), 64, slice_hw=(128, 128), imgsz=128, conf=0.25, device=dev, cap=256)
    h2, _ = synth.synth_head_from_labels(sp.S, 64, in_hw=sp.in_hw, src_hw=sp.slice_hw, seed=7)
    sp(f, h2.to(dev))
    torch.cuda.synchronize()
    print("sanitize cases ok")


if __name__ == "__main__":
    main()
