"""Generate the committed golden fixtures from the reference tree (run in the dev container only).

    python tests/golden/make_golden.py            # needs /root/reference

Outputs (all small, committed):
  labels.npz                 -- the 200 YOLO label files of roadmap1.v3i.yolov8 (4 259 boxes) + image sizes
  frames/*.jpg               -- three real dataset frames (1920x1200, 1600x900, 1919x1194) as test inputs
  letterbox_golden.json      -- sha256 of the real-cv2 letterbox of those frames (square + rect)
  rank_classifier_kat.npz    -- rank_classifier.pt weights (fp16), the 67 valid crops, the real
                                PIL/torchvision 64x64 outputs, reference logits, labels
  nms_golden.npz             -- torchvision.ops.nms kept indices on seeded class-offset boxes

/root/reference does not exist on the GPU box, so GPU tests read only these files.
"""

import glob
import hashlib
import json
import os
import shutil
import sys

import cv2
import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import classifier, letterbox, roi  # noqa: E402


def make_labels():
    ds = os.path.join(REF, "roadmap1.v3i.yolov8")
    boxes, box_img, img_hw, names = [], [], [], []
    for split in ("train", "valid", "test"):
        for lab in sorted(glob.glob(os.path.join(ds, split, "labels", "*.txt"))):
            img = os.path.join(ds, split, "images", os.path.basename(lab)[:-4] + ".jpg")
            im = cv2.imread(img)
            idx = len(img_hw)
            img_hw.append(im.shape[:2])
            names.append(f"{split}/{os.path.basename(img)}")
            with open(lab) as f:
                for line in f:
                    p = line.split()
                    if len(p) == 5:
                        boxes.append([float(v) for v in p])
                        box_img.append(idx)
    np.savez_compressed(os.path.join(HERE, "labels.npz"), boxes=np.asarray(boxes, np.float64),
                        box_img=np.asarray(box_img, np.int32), img_hw=np.asarray(img_hw, np.int32))
    print("labels:", len(boxes), "boxes in", len(img_hw), "images")
    return names, np.asarray(img_hw)


def make_frames(names, img_hw):
    os.makedirs(os.path.join(HERE, "frames"), exist_ok=True)
    want = {(1200, 1920): "frame_1920x1200.jpg", (900, 1600): "frame_1600x900.jpg",
            (1194, 1919): "frame_1919x1194.jpg"}
    gold = {}
    for hw, out in want.items():
        i = int(np.nonzero((img_hw[:, 0] == hw[0]) & (img_hw[:, 1] == hw[1]))[0][0])
        src = os.path.join(REF, "roadmap1.v3i.yolov8", names[i].split("/")[0], "images", names[i].split("/")[1])
        dst = os.path.join(HERE, "frames", out)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        im = cv2.imread(dst)
        for auto in (False, True):
            lb = letterbox.letterbox_ref(im, (640, 640), auto=auto)
            gold[f"{out}|auto={int(auto)}"] = dict(shape=list(lb.shape),
                                                   sha256=hashlib.sha256(lb.tobytes()).hexdigest())
    json.dump(gold, open(os.path.join(HERE, "letterbox_golden.json"), "w"), indent=1, sort_keys=True)
    print("frames:", list(gold))


def make_classifier():
    sd, names, tf, metrics, eps = classifier.load_checkpoint(os.path.join(REF, "rank_classifier.pt"))
    assert eps == classifier.BN_EPS and names == classifier.NAMES
    name2id = {v: k for k, v in names.items()}
    crops, shapes, labels, outs = [], [], [], []
    for d in sorted(os.listdir(os.path.join(REF, "rank_classifier", "valid"))):
        for f in sorted(glob.glob(os.path.join(REF, "rank_classifier", "valid", d, "*.jpg"))):
            im = cv2.imread(f)
            crops.append(im.reshape(-1))
            shapes.append(im.shape[:2])
            labels.append(name2id[d])
            t = roi.classify_preprocess_ref(im)                      # real PIL/torchvision leaves
            outs.append((t * 255).round().to(torch.uint8).numpy())   # exact: values are k/255
    x = torch.stack([torch.from_numpy(o).float().div(255) for o in outs])
    logits = classifier.forward_logits(sd, x)
    labels = np.asarray(labels)
    top1 = int((logits.argmax(1).numpy() == labels).sum())
    print("classifier KAT: top1", top1, "/", len(labels), "logged", metrics)
    z = classifier.state_dict_to_npz_dict(sd)
    z.update(crops=np.concatenate(crops), crop_hw=np.asarray(shapes, np.int32), labels=labels.astype(np.int32),
             roi_u8=np.stack(outs), logits=logits.numpy().astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "rank_classifier_kat.npz"), **z)


def make_nms():
    g = torch.Generator().manual_seed(1234)
    out = {}
    for t, (n, nc, thr) in enumerate([(300, 64, 0.45), (3000, 80, 0.7), (1000, 64, 0.6), (500, 1, 0.5)]):
        xy = torch.rand((n, 2), generator=g) * 600
        wh = torch.rand((n, 2), generator=g) * 120 + 2
        boxes = torch.cat((xy, xy + wh), 1)
        scores = torch.rand((n,), generator=g)
        scores[::7] = scores[3]                 # ties
        cls = torch.randint(0, nc, (n,), generator=g).float()
        keep = torchvision.ops.nms(boxes + cls[:, None] * 7680, scores, thr)
        out[f"boxes{t}"], out[f"scores{t}"], out[f"cls{t}"] = boxes.numpy(), scores.numpy(), cls.numpy()
        out[f"thr{t}"], out[f"keep{t}"] = np.float64(thr), keep.numpy()
    np.savez_compressed(os.path.join(HERE, "nms_golden.npz"), **out)
    print("nms golden:", [len(out[f'keep{t}']) for t in range(4)])


if __name__ == "__main__":
    names, img_hw = make_labels()
    make_frames(names, img_hw)
    make_classifier()
    make_nms()
