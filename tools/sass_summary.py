#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libb200yolo.so (cuobjdump -sass): the evidence for which Blackwell paths the
kernels use -- UBLKCP (TMA 1-D bulk copies), SYNCS (mbarrier), LDGSTS (cp.async), 128-bit global accesses, packed fp32
(FADD2 / FMUL2 / FFMA2), REDUX/CREDUX, and the absence of tensor-core instructions (the path has no contraction).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "manual_yolo_b200", "libb200yolo.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.EF.128", "STG.E.64", "LDS.128",
         "LDS.U8", "STS.U8", "IMAD", "PRMT", "FFMA2", "FADD2", "FMUL2", "DFMA", "CREDUX", "REDUX", "ELECT", "MATCH", "VOTE",
         "SHFL", "ATOMS", "ATOMG", "RED", "BAR", "HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                head, _, width = w.partition(".E.")
                if width and head in ("LDG", "STG"):       # LDG.E.NA.128.CONSTANT etc.: opcode + access width anywhere
                    parts = op.split(".")
                    if parts[0] == head and width.split(".")[-1] in parts and (("EF" in parts) == ("EF" in width.split("."))):
                        cur[w] += 1
                elif op == w or op.startswith(w + "."):
                    cur[w] += 1
    print(f"# SASS summary of {os.path.relpath(SO, ROOT)} (sm_100a; cuobjdump -sass; instruction counts are static)")
    print("# UBLKCP = cp.async.bulk (TMA 1-D), SYNCS = mbarrier ops, LDGSTS = cp.async, FFMA2/FADD2/FMUL2 = packed fp32 (Blackwell)")
    print("# no HMMA / IMMA / UTC*MMA / LDTM / STTM anywhere: the path has no dense contraction (SURVEY.md section 7.1)\n")
    for name, c in kernels.items():
        hits = ", ".join(f"{w} {c[w]}" for w in WATCH if c[w])
        print(f"{name}\n    {c['_total']} instructions: {hits}")
    tensor = sum(c[w] for c in kernels.values() for w in ("HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"))
    print(f"\n# tensor-core / TMEM instructions in the library: {tensor}")


if __name__ == "__main__":
    sys.exit(main())
