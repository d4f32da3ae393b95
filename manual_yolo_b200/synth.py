"""Seeded synthetic inputs for the detection path.

The detector weights (``poker_model.pt``, ``yolov8m.pt``) are absent from the reference
(``/root/reference/.MISSING_LARGE_BLOBS:3-4``), so head tensors are synthesised from the dataset's
label geometry (``/root/reference/roadmap1.v3i.yolov8/*/labels/*.txt``, committed as the small
fixture ``tests/golden/labels.npz`` by ``tests/golden/make_golden.py``): every ground-truth box is
inverse-decoded into peaked DFL logits on the anchors that fall inside it, plus clutter.

Everything is generated on the CPU with a seeded ``torch.Generator`` so the oracle and the CUDA
path see identical bits.
"""

from __future__ import annotations

import os

import numpy as np
import torch

REG_MAX = 16
RANK_CLASS_IDS = (6, 11, 16, 21, 26, 37, 43)  # *_rank ids, roadmap1.v3i.yolov8/data.yaml:6
_DEFAULT_LABELS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                               "tests", "golden", "labels.npz")


def load_labels(path=_DEFAULT_LABELS):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def level_shapes(in_h, in_w, strides=(8, 16, 32)):
    return [(in_h // s, in_w // s) for s in strides]


def num_anchors(in_h, in_w, strides=(8, 16, 32)):
    return sum(h * w for h, w in level_shapes(in_h, in_w, strides))


def synth_frames(B, H=1200, W=1920, seed=0, device="cpu"):
    """(B,H,W,3) uint8 BGR noise frames (config 2/5: ``torch.randint`` seed 0)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g, device=device)


def _letterbox_params(src_hw, in_hw):
    h, w = src_hw
    r = min(in_hw[0] / h, in_hw[1] / w)
    nw, nh = int(round(w * r)), int(round(h * r))
    left = int(round((in_hw[1] - nw) / 2 - 0.1))
    top = int(round((in_hw[0] - nh) / 2 - 0.1))
    return r, left, top


def _f32_bits(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous().view(torch.int32).to(torch.int64)


def guard_band(head, nc, conf_thres, ulp=16, max_iter=20):
    """Nudge class logits so that no best-class sigmoid score lies within ``ulp`` ulps of
    ``conf_thres`` and no two same-image candidate scores lie within ``ulp`` ulps of each other
    unless their logits are bit-identical.  In place; returns #nudges."""
    B = head.shape[0]
    cls = head[:, 4 * REG_MAX:, :]
    thr_bits = int(_f32_bits(torch.tensor([conf_thres], dtype=torch.float32))[0])
    nudges = 0
    for b in range(B):
        for _ in range(max_iter):
            logit, j = cls[b].max(0)
            score = logit.sigmoid()
            bits = _f32_bits(score)
            near_thr = (bits - thr_bits).abs() <= ulp
            cand = (score > conf_thres) | near_thr
            idx = cand.nonzero().view(-1)
            bad = near_thr.clone()
            if idx.numel() > 1:
                sb, order = bits[idx].sort()
                lg = logit[idx][order]
                close = ((sb[1:] - sb[:-1]) <= ulp) & (lg[1:] != lg[:-1])
                bad[idx[order[1:][close]]] = True
            bad_idx = bad.nonzero().view(-1)
            if bad_idx.numel() == 0:
                break
            nudges += int(bad_idx.numel())
            # move the offending anchor's best logit by a relative 2^-12 step (>> 16 ulp of score)
            cls[b, j[bad_idx], bad_idx] *= 1.0 + 2.0 ** -12 * (1 + torch.arange(bad_idx.numel()) % 7)
        else:
            raise RuntimeError("guard band did not converge")
    return nudges


def synth_head_from_labels(B, nc=64, in_hw=(640, 640), src_hw=(1200, 1920), seed=0, conf_thres=0.25,
                           strides=(8, 16, 32), labels=None, guard_ulp=16, max_anchors_per_box=9):
    """Config 1/2 head: (B, 64+nc, A) fp32 + list of per-image ground-truth boxes (letterboxed xyxy, cls)."""
    labels = labels if labels is not None else load_labels()
    g = torch.Generator()
    g.manual_seed(seed)
    lv = level_shapes(in_hw[0], in_hw[1], strides)
    A = sum(h * w for h, w in lv)
    no = 4 * REG_MAX + nc
    head = torch.randn((B, no, A), generator=g, dtype=torch.float32)
    head[:, 4 * REG_MAX:, :] = head[:, 4 * REG_MAX:, :] - 6.0          # class clutter ~ N(-6,1)
    # images of the requested source size
    sizes = labels["img_hw"]
    pool = np.nonzero((sizes[:, 0] == src_hw[0]) & (sizes[:, 1] == src_hw[1]))[0]
    if pool.size == 0:
        pool = np.arange(sizes.shape[0])
    pick = torch.randint(0, len(pool), (B,), generator=g).numpy()
    r, left, top = _letterbox_params(src_hw, in_hw)
    offs = np.cumsum([0] + [h * w for h, w in lv])
    bins = torch.arange(REG_MAX, dtype=torch.float32)
    gts = []
    for b in range(B):
        img = pool[pick[b]]
        rows = labels["boxes"][labels["box_img"] == img]
        gt = []
        taken = set()          # one object per anchor, as a trained head produces (no class/box mix-ups)
        for cls_id, cx, cy, bw, bh in rows:
            cls_id = int(cls_id) % nc
            x1 = (cx - bw / 2) * src_hw[1] * r + left
            x2 = (cx + bw / 2) * src_hw[1] * r + left
            y1 = (cy - bh / 2) * src_hw[0] * r + top
            y2 = (cy + bh / 2) * src_hw[0] * r + top
            gt.append((x1, y1, x2, y2, cls_id))
            # smallest stride whose DFL range (15 bins) covers the box from any interior anchor
            for li, s in enumerate(strides):
                if max(x2 - x1, y2 - y1) / s <= 13.0 or li == len(strides) - 1:
                    break
            h, w = lv[li]
            ax0, ax1 = int(np.ceil(x1 / s - 0.5)), int(np.floor(x2 / s - 0.5))
            ay0, ay1 = int(np.ceil(y1 / s - 0.5)), int(np.floor(y2 / s - 0.5))
            ax0, ax1, ay0, ay1 = max(ax0, 0), min(ax1, w - 1), max(ay0, 0), min(ay1, h - 1)
            if ax1 < ax0 or ay1 < ay0:      # box smaller than a cell: nearest anchor
                ax0 = ax1 = min(max(int(round((x1 + x2) / 2 / s - 0.5)), 0), w - 1)
                ay0 = ay1 = min(max(int(round((y1 + y2) / 2 / s - 0.5)), 0), h - 1)
            cells = [(ax, ay) for ay in range(ay0, ay1 + 1) for ax in range(ax0, ax1 + 1)]
            mx, my = (x1 + x2) / 2 / s - 0.5, (y1 + y2) / 2 / s - 0.5
            cells.sort(key=lambda c: (c[0] - mx) ** 2 + (c[1] - my) ** 2)
            k = 1 + int(torch.randint(0, max_anchors_per_box, (1,), generator=g))
            cells = [cc for cc in cells if (li, cc[0], cc[1]) not in taken]
            for ax, ay in cells[:k]:
                taken.add((li, ax, ay))
                a = int(offs[li]) + ay * w + ax
                px, py = (ax + 0.5) * s, (ay + 0.5) * s
                dist = torch.tensor([px - x1, py - y1, x2 - px, y2 - py], dtype=torch.float32) / s
                dist = dist.clamp(0.0, REG_MAX - 1.01)
                logits = -((bins[None, :] - dist[:, None]) ** 2) / (2 * 0.7 ** 2)
                head[b, :4 * REG_MAX, a] = (logits * 1.0 + 0.05 * torch.randn((4, REG_MAX), generator=g)).reshape(-1)
                head[b, 4 * REG_MAX + cls_id, a] = 0.5 + 3.5 * float(torch.rand((1,), generator=g))
        gts.append(np.asarray(gt, np.float32).reshape(-1, 5))
        # a trained head does not fire on background: keep the N(-6,1) class clutter of unassigned anchors
        # below the threshold (logit(conf) - 1), so every candidate belongs to a labelled object and the
        # ROI sizes follow the dataset's rank-box statistics instead of random clutter boxes
        bg = torch.ones(A, dtype=torch.bool)
        if taken:
            bg[torch.tensor([int(offs[l]) + y * lv[l][1] + x for (l, x, y) in taken])] = False
        ceil_logit = float(np.log(conf_thres / (1.0 - conf_thres))) - 1.0 if 0.0 < conf_thres < 1.0 else -2.0
        cl = head[b, 4 * REG_MAX:, :]
        cl[:, bg] = cl[:, bg].clamp(max=ceil_logit)
    if guard_ulp:
        guard_band(head, nc, conf_thres, guard_ulp)
    return head, gts


def synth_head_dense(B, nc=80, in_hw=(640, 640), seed=0, objects=40, anchors_per_obj=6,
                     bg_mu=-3.5, bg_sigma=1.2, strides=(8, 16, 32), adversarial=True):
    """Config 3 (NMS-heavy eval regime): ~every anchor passes conf=0.001.  (B, 64+nc, A) fp32."""
    g = torch.Generator()
    g.manual_seed(seed)
    lv = level_shapes(in_hw[0], in_hw[1], strides)
    A = sum(h * w for h, w in lv)
    no = 4 * REG_MAX + nc
    head = torch.randn((B, no, A), generator=g, dtype=torch.float32)
    head[:, 4 * REG_MAX:, :] = head[:, 4 * REG_MAX:, :] * bg_sigma + bg_mu
    # sharpen DFL a bit so boxes have plausible extents
    head[:, :4 * REG_MAX, :] *= 2.0
    offs = np.cumsum([0] + [h * w for h, w in lv])
    bins = torch.arange(REG_MAX, dtype=torch.float32)
    for b in range(B):
        for _ in range(objects):
            li = int(torch.randint(0, len(strides), (1,), generator=g))
            s = strides[li]
            h, w = lv[li]
            ax = int(torch.randint(1, w - 2, (1,), generator=g))
            ay = int(torch.randint(1, h - 2, (1,), generator=g))
            cls_id = int(torch.randint(0, nc, (1,), generator=g))
            half = 1.0 + 9.0 * torch.rand((4,), generator=g)
            for k in range(anchors_per_obj):
                cx, cy = ax + (k % 3) - 1, ay + (k // 3)
                if not (0 <= cx < w and 0 <= cy < h):
                    continue
                a = int(offs[li]) + cy * w + cx
                dist = (half + torch.tensor([(k % 3) - 1.0, float(k // 3), 1.0 - (k % 3), -float(k // 3)])).clamp(0.2, 14.5)
                logits = -((bins[None, :] - dist[:, None]) ** 2) / (2 * 0.6 ** 2)
                head[b, :4 * REG_MAX, a] = logits.reshape(-1)
                head[b, 4 * REG_MAX + cls_id, a] = 1.0 + 3.0 * float(torch.rand((1,), generator=g))
        if adversarial and A >= 64:
            # exact score ties on different anchors (bit-identical columns)
            head[b, :, 11] = head[b, :, 7]
            head[b, :, 4000 % A] = head[b, :, 7]
            # identical box, different class, same score
            head[b, :, 13] = head[b, :, 12]
            col = head[b, 4 * REG_MAX:, 13].clone()
            j = int(col.argmax())
            col[(j + 1) % nc], col[j] = col[j].clone(), col[(j + 1) % nc].clone()
            head[b, 4 * REG_MAX:, 13] = col
            # highest class id carries the largest class offset (ulp 0.0625 at 79*7680)
            head[b, 4 * REG_MAX + nc - 1, 21] = 5.0
            head[b, 4 * REG_MAX + nc - 1, 22] = 4.5
            # degenerate zero-area boxes: huge mass on bin 0 for all four sides
            for a in (30, 31):
                head[b, :4 * REG_MAX, a] = -60.0
                head[b, 0:4 * REG_MAX:REG_MAX, a] = 60.0
    return head


def synth_rois(N, B, frame_hw=(1200, 1920), seed=0, down_frac=0.25, border_frac=0.05):
    """Config 4: N float xyxy boxes in source pixels + frame index, from the empirical rank-box
    size distribution (w 30-103, h 26-93).  ``down_frac`` of them are
    forced to short side > 64 (antialias down-scaling); ``border_frac`` touch the image border."""
    g = torch.Generator()
    g.manual_seed(seed)
    H, W = frame_hw
    w = 30 + 73 * torch.rand((N,), generator=g)
    h = 26 + 67 * torch.rand((N,), generator=g)
    big = torch.rand((N,), generator=g) < down_frac
    w = torch.where(big, 66 + 120 * torch.rand((N,), generator=g), w)
    h = torch.where(big, 66 + 120 * torch.rand((N,), generator=g), h)
    x1 = torch.rand((N,), generator=g) * (W - w - 1)
    y1 = torch.rand((N,), generator=g) * (H - h - 1)
    border = torch.rand((N,), generator=g) < border_frac
    side = torch.randint(0, 4, (N,), generator=g)
    x1 = torch.where(border & (side == 0), torch.zeros(()), x1)
    y1 = torch.where(border & (side == 1), torch.zeros(()), y1)
    x1 = torch.where(border & (side == 2), W - w + 2.0, x1)
    y1 = torch.where(border & (side == 3), H - h + 2.0, y1)
    boxes = torch.stack((x1, y1, x1 + w, y1 + h), 1).float()
    bidx = torch.randint(0, B, (N,), generator=g, dtype=torch.int32)
    return boxes, bidx
