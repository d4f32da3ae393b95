"""Generate tests/golden/clean_detections_golden.json by EXECUTING the reference's own ``create_clean_detections``
(/root/reference/detect.py:253-310), cut out of the AST unmodified, with a stand-in for ``sv.Detections`` that just
records its keyword arguments (supervision is not installed).  Dev container only.

    python tests/golden/make_clean_detections_golden.py

Pins the cleaning rules of the tracker hand-off (``manual_yolo_b200.handoff.to_tracker_arrays``).
"""
import ast
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/detect.py"


class _Detections:
    def __init__(self, **kw):
        self.kw = kw

    @staticmethod
    def empty():
        return _Detections(empty=True)


class _SV:
    Detections = _Detections


def main():
    tree = ast.parse(open(REF, encoding="utf-8").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "create_clean_detections"]
    ns = {"np": np, "sv": _SV}
    exec(compile(ast.Module(body=fn, type_ignores=[]), REF, "exec"), ns)
    clean = ns["create_clean_detections"]
    nan = float("nan")
    cases = [
        dict(xyxy=[[1.5, 2.5, 30.25, 40.75], [5, 6, 7, 8], [0, 0, 1, 1]], class_id=[6.0, nan, 11.0], confidence=[0.9, nan, 0.25],
             tracker_id=[3, None, nan]),
        dict(xyxy=[[10, 20, 30, 40]], class_id=[None], confidence=[None], tracker_id=None),
        dict(xyxy=[[10, 20, 30, 40], [1, 1, 2, 2]], class_id=None, confidence=None, tracker_id=[7, 8]),
        dict(xyxy=[], class_id=[], confidence=[], tracker_id=[]),
    ]
    out = []
    for c in cases:
        d = clean(c["xyxy"], c["class_id"], c["confidence"], c["tracker_id"]).kw
        rec = {"in": json.loads(json.dumps(c).replace("NaN", '"nan"')), "empty": bool(d.get("empty", False))}
        if not rec["empty"]:
            rec["xyxy"] = np.asarray(d["xyxy"]).tolist()
            rec["xyxy_dtype"] = str(np.asarray(d["xyxy"]).dtype)
            rec["class_id"] = np.asarray(d["class_id"]).tolist()
            rec["class_id_dtype"] = str(np.asarray(d["class_id"]).dtype)
            rec["confidence"] = np.asarray(d["confidence"]).tolist()
            rec["confidence_dtype"] = str(np.asarray(d["confidence"]).dtype)
            rec["tracker_id"] = None if d["tracker_id"] is None else np.asarray(d["tracker_id"]).tolist()
        out.append(rec)
    with open(os.path.join(HERE, "clean_detections_golden.json"), "w") as f:
        json.dump({"source": "executed from /root/reference/detect.py create_clean_detections (:253-310), unmodified; "
                             "sv.Detections replaced by a kwargs recorder", "cases": out}, f)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
