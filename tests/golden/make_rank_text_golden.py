"""Generate tests/golden/rank_text_golden.json by EXECUTING the reference's own helper functions (dev container
only; /root/reference does not exist on the GPU box).

    python tests/golden/make_rank_text_golden.py

detect.py cannot be imported (it imports mss / easyocr / supervision / ultralytics at module top), so the two
helpers on the classifier hand-off are cut out of its AST and executed in isolation, unmodified:
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
        if isinstance(node, ast.FunctionDef) and node.name in ("normalize_rank_text", "safe_crop", "classify_card_rank"):
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
    # classify_card_rank (detect.py:115-139), executed unmodified against a stand-in for the global `rank_model`
    # (the YOLO classifier object): a callable returning [results] with .probs.top1 / .probs.top1conf, and .names --
    # exactly the attributes the function touches.  Pins the confidence thresholds and the text clean-up decision.
    class _Probs:
        def __init__(self, top1, conf):
            self.top1, self.top1conf = top1, conf

    class _Results:
        def __init__(self, top1, conf):
            self.probs = _Probs(top1, conf)

    class _RankModel:
        names = {0: "10", 1: "2", 2: "3", 3: "4", 4: "5", 5: "6", 6: "7", 7: "8", 8: "9", 9: "A", 10: "J", 11: "K", 12: "Q",
                 13: "joker", 14: "1O", 15: "a"}          # 0-12: the checkpoint's names; 13-15: clean-up / fallback cases

        def __call__(self, crop):
            return [_Results(self.next_top1, self.next_conf)]
    rm = _RankModel()
    ns["rank_model"] = rm
    classify = ns["classify_card_rank"]
    crop = np.zeros((8, 8, 3), np.uint8)
    classify_vectors = []
    for top1 in list(range(16)) + [99]:                    # 99: an id the names table does not hold
        for conf in (0.0, 0.19, 0.1999999, 0.2, 0.2000001, 0.39, 0.3999999, 0.4, 0.4000001, 0.7, 1.0):
            for cname in ("card1_rank", "flop2_rank", "turn_rank", "TURN_rank", "river_rank", "River", ""):
                rm.next_top1, rm.next_conf = top1, np.float32(conf)
                classify_vectors.append([rm.names.get(top1, ""), float(np.float32(conf)), cname, classify(crop, cname)])
    empty = [classify(None, "card1_rank"), classify(np.zeros((0, 4, 3), np.uint8), "turn_rank")]
    out = {"source": "executed from /root/reference/detect.py (normalize_rank_text :60-98, safe_crop :100-113, "
                     "classify_card_rank :115-139 with a stand-in rank_model), unmodified",
           "classify_card_rank": classify_vectors, "classify_card_rank_empty": empty,
           "valid_card_ranks": sorted(ns["VALID_CARD_RANKS"]), "frame_hw": [1200, 1920],
           "normalize_rank_text": text_vectors, "safe_crop_shapes": crop_vectors}
    with open(os.path.join(HERE, "rank_text_golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(text_vectors), "text vectors,", len(crop_vectors), "crop vectors,", len(classify_vectors), "classify vectors")


if __name__ == "__main__":
    main()
