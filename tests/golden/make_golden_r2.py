"""Round-2 golden fixtures generated from the reference tree (run in the dev container only):

    python tests/golden/make_golden_r2.py            # needs /root/reference

Outputs (committed):
  frames/test2.png                 -- BASELINE configs[0]'s literal input (/root/reference/test2.png, 1600x900 BGRA)
  frames_r2/*.jpg                  -- 10 dataset frames per production size (1920x1200, 1600x900) + the 1700x1034 one
  letterbox_golden_r2.json         -- per frame and letterbox mode: shape + sha256 of the REAL cv2 letterbox (uint8) and of
                                      the fp32 network input (BasePredictor.preprocess restated on the real cv2 leaves)
  rank_crops_all.npz               -- all 579 crops of rank_classifier/{train,valid} as their original JPEG bytes, labels,
                                      split flag, sha256 of the real PIL/torchvision 64x64 output of each crop, and the
                                      oracle's top-1 class through rank_classifier.pt (63/67 on valid: results.csv:21)

/root/reference does not exist on the GPU box, so the GPU tests read only these files.
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

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import classifier, letterbox, roi  # noqa: E402


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_frames():
    ds = os.path.join(REF, "roadmap1.v3i.yolov8")
    os.makedirs(os.path.join(HERE, "frames_r2"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "test2.png"), os.path.join(HERE, "frames", "test2.png"))
    os.chmod(os.path.join(HERE, "frames", "test2.png"), 0o644)
    picked = {(1200, 1920): [], (900, 1600): [], (1034, 1700): []}
    for split in ("test", "valid", "train"):
        for f in sorted(glob.glob(os.path.join(ds, split, "images", "*.jpg"))):
            hw = cv2.imread(f).shape[:2]
            if hw in picked and len(picked[hw]) < (1 if hw == (1034, 1700) else 10):
                picked[hw].append(f)
    gold = {}
    files = [("frames/test2.png", os.path.join(HERE, "frames", "test2.png"))]
    for hw, fs in picked.items():
        for i, f in enumerate(fs):
            name = f"frames_r2/f{hw[1]}x{hw[0]}_{i:02d}.jpg"
            shutil.copyfile(f, os.path.join(HERE, name))
            os.chmod(os.path.join(HERE, name), 0o644)
            files.append((name, os.path.join(HERE, name)))
    for name, path in files:
        im = cv2.imread(path)                       # yolo.py:360 reads with cv2.imread (3-channel BGR)
        for auto in (False, True):
            lb = letterbox.letterbox_ref(im, (640, 640), auto=auto)
            net = letterbox.preprocess_ref([im], (640, 640), auto=auto)
            gold[f"{name}|auto={int(auto)}"] = dict(src_hw=list(im.shape[:2]), shape=list(lb.shape), sha256=_sha(lb),
                                                    net_shape=list(net.shape), net_sha256=_sha(net.numpy()))
    json.dump(gold, open(os.path.join(HERE, "letterbox_golden_r2.json"), "w"), indent=1, sort_keys=True)
    print("frames:", len(files), "entries:", len(gold))


def make_crops():
    sd, names, tf, metrics, eps = classifier.load_checkpoint(os.path.join(REF, "rank_classifier.pt"))
    name2id = {v: k for k, v in names.items()}
    blobs, offs, labels, split_flag, shas, outs = [], [0], [], [], [], []
    for si, split in enumerate(("train", "valid")):
        for d in sorted(os.listdir(os.path.join(REF, "rank_classifier", split))):
            for f in sorted(glob.glob(os.path.join(REF, "rank_classifier", split, d, "*.jpg"))):
                raw = np.fromfile(f, np.uint8)
                blobs.append(raw)
                offs.append(offs[-1] + raw.size)
                labels.append(name2id[d])
                split_flag.append(si)
                im = cv2.imdecode(raw, cv2.IMREAD_COLOR)
                assert np.array_equal(im, cv2.imread(f))
                t = roi.classify_preprocess_ref(im)                      # real PIL/torchvision leaves
                u8 = (t * 255).round().to(torch.uint8)
                assert torch.equal(u8.float().div(255), t)
                shas.append(_sha(u8.numpy()))
                outs.append(t)
    logits = classifier.forward_logits(sd, torch.stack(outs))
    top1 = logits.argmax(1).numpy().astype(np.int32)
    labels = np.asarray(labels, np.int32)
    split_flag = np.asarray(split_flag, np.int8)
    v = split_flag == 1
    print("crops:", len(labels), "valid top-1", int((top1[v] == labels[v]).sum()), "/", int(v.sum()),
          "train top-1", int((top1[~v] == labels[~v]).sum()), "/", int((~v).sum()))
    assert int((top1[v] == labels[v]).sum()) == 63 and int(v.sum()) == 67       # runs/rank_classifier/results.csv:21
    np.savez_compressed(os.path.join(HERE, "rank_crops_all.npz"), jpeg=np.concatenate(blobs),
                        offs=np.asarray(offs, np.int64), labels=labels, split=split_flag,
                        roi_sha256=np.asarray(shas), oracle_top1=top1)


if __name__ == "__main__":
    make_frames()
    make_crops()
