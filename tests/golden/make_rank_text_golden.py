"""Generate tests/golden/rank_text_golden.json by EXECUTING the reference's own helper functions (dev container
only; /root/reference does not exist on the GPU box).

    python tests/golden/make_rank_text_golden.py

detect.py cannot be imported (it imports mss / easyocr / supervision / ultralytics at module top), so the two
pure helpers on the classifier hand-off are cut out of its AST and executed in isolation, unmodified:
``normalize_rank_text`` (detect.py:60-98) with its tables ``VALID_CARD_RANKS`` / ``MAPPING_CORRECTION``
(detect.py:36-37), and ``safe_crop`` (detect.py:100-113).  Their outputs on a fixed input set are the golden
vectors that pin ``manual_yolo_b200.handoff.normalize_rank_text`` / ``rank_text_from_top1`` and the crop geometry
of K5 (``oracle.boxes.safe_crop_box_ref``).
"""
import ast
import itertools
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/detect.py"


def load_reference_helpers():
    tree = ast.parse(open(REF, encoding="utf-8").read())
    keep = []
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(getattr(t, "id", "") in ("VALID_CARD_RANKS", "MAPPING_CORRECTION") for t in node.targets):
            keep.append(node)
        if isinstance(node, ast.FunctionDef) and node.name in ("normalize_rank_text", "safe_crop"):
            keep.append(node)
    ns = {}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns


def main():
    ns = load_reference_helpers()
    norm, safe_crop = ns["normalize_rank_text"], ns["safe_crop"]
    singles = list("AKQJT0123456789OISZBakqjt|xyz ") + ["10", "1O", " 10 ", "l0", "IO", "T ", " a", "Q ", "11", "12", "01", "00", "",
                                                        "K|", "|", "||", "1|", "9 ", " 2", "A A", "JO", "S5", "B", "b", "z", "o", "i"]
    pairs = ["".join(p) for p in itertools.product("AKQJT01259O|I ", repeat=2)]
    texts = sorted(set(singles + pairs))
    text_vectors = [[t, norm(t)] for t in texts]
    rng = np.random.default_rng(0)
    frame = np.zeros((1200, 1920, 3), np.uint8)
    crop_vectors = []
    for _ in range(400):
        x1, y1 = rng.uniform(-40, 1930), rng.uniform(-40, 1210)
        w, h = rng.uniform(-10, 140), rng.uniform(-10, 140)
        x1i, y1i, x2i, y2i = int(x1), int(y1), int(x1 + w), int(y1 + h)
        pad = int(rng.integers(0, 9))
        c = safe_crop(frame, x1i, y1i, x2i, y2i, pad=pad)
        crop_vectors.append([x1i, y1i, x2i, y2i, pad, None if c is None else [int(c.shape[0]), int(c.shape[1])]])
    out = {"source": "executed from /root/reference/detect.py (normalize_rank_text :60-98, safe_crop :100-113), unmodified",
           "valid_card_ranks": sorted(ns["VALID_CARD_RANKS"]), "frame_hw": [1200, 1920],
           "normalize_rank_text": text_vectors, "safe_crop_shapes": crop_vectors}
    with open(os.path.join(HERE, "rank_text_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(text_vectors), "text vectors,", len(crop_vectors), "crop vectors")


if __name__ == "__main__":
    main()
