"""Generate tests/golden/pipe_records_golden.json by EXECUTING the reference's own ``parse_ultralytics_results``
(/root/reference/pipe.py:100-134), cut out of the file's AST unmodified (pipe.py itself cannot be imported: it
imports mss / easyocr / deep_sort_realtime at module top).  Dev container only.

    python tests/golden/make_pipe_records_golden.py

The function is the consumer of SURVEY row a11 (``Results.boxes`` -> per-detection dicts with ``int()``-truncated,
image-clamped coordinates).  It is fed minimal stand-ins for ``Results`` / ``Boxes`` built from torch tensors; its
outputs pin ``manual_yolo_b200.handoff.to_pipe_records``.
"""
import ast
import json
import os
from typing import Dict, List  # noqa: F401  (names used by the extracted function's annotations)

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/pipe.py"


class _Box:
    def __init__(self, row):
        self.xyxy = row[None, :4].clone()
        self.conf = row[4:5].clone()
        self.cls = row[5:6].clone()


class _Res:
    def __init__(self, rows, names):
        self.boxes = [_Box(r) for r in rows]
        self.names = names


def load_parse():
    tree = ast.parse(open(REF, encoding="utf-8").read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "parse_ultralytics_results"]
    ns = {"List": List, "Dict": Dict, "np": np}
    exec(compile(ast.Module(body=fn, type_ignores=[]), REF, "exec"), ns)
    return ns["parse_ultralytics_results"]


def main():
    parse = load_parse()
    g = torch.Generator().manual_seed(0)
    names = {i: f"name{i}" for i in range(0, 64, 2)}             # odd ids fall back to "class<id>"
    cases = []
    for (h, w) in [(543, 770), (1200, 1920)]:
        n = 40
        rows = torch.zeros((n, 6))
        rows[:, 0] = torch.rand(n, generator=g) * (w + 40) - 20
        rows[:, 1] = torch.rand(n, generator=g) * (h + 40) - 20
        rows[:, 2] = rows[:, 0] + torch.rand(n, generator=g) * 200
        rows[:, 3] = rows[:, 1] + torch.rand(n, generator=g) * 200
        rows[:, 4] = torch.rand(n, generator=g)
        rows[:, 5] = torch.randint(0, 64, (n,), generator=g).float()
        out = parse([_Res(rows, names)], (h, w, 3))
        cases.append({"image_shape": [h, w, 3], "rows": rows.tolist(), "records": out})
    cases.append({"image_shape": [100, 100, 3], "rows": [], "records": parse([_Res(torch.zeros((0, 6)), names)], (100, 100, 3))})
    with open(os.path.join(HERE, "pipe_records_golden.json"), "w") as f:
        json.dump({"source": "executed from /root/reference/pipe.py parse_ultralytics_results (:100-134), unmodified",
                   "names": {str(k): v for k, v in names.items()}, "cases": cases}, f)
    print(sum(len(c["records"]) for c in cases), "records")


if __name__ == "__main__":
    main()
