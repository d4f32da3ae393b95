#!/usr/bin/env bash
# The stand-in for compute-sanitizer memcheck on a pool where the tool is closed: build the library with -DB200_CHECKS
# (device-side assertions on every shared-memory ring / strip / key-array index and global store index the kernels
# compute; a failed check prints its location and traps) and run the whole GPU test suite against it.
#   tools/checked.sh build     (here, no GPU)        tools/checked.sh run   (on a GPU box; log -> gpurun_out/)
set -uo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
cd "$ROOT"
LIB="$ROOT/manual_yolo_b200/libb200yolo_checked.so"
case "${1:-build}" in
  build)
    B200YOLO_OUT="$LIB" B200YOLO_OBJ="$ROOT/manual_yolo_b200/csrc/_obj_checked" bash manual_yolo_b200/csrc/build.sh -DB200_CHECKS
    ;;
  run)
    [ -f "$LIB" ] || { echo "build first"; exit 1; }
    mkdir -p gpurun_out
    n=$(grep -c "B200_CHECK(" manual_yolo_b200/csrc/*.cu | awk -F: '{s+=$2} END {print s}')
    { echo "checked build: $n B200_CHECK sites in csrc/*.cu; $(strings "$LIB" | grep -c 'B200_CHECK failed') format string(s) and $(cuobjdump -sass "$LIB" 2>/dev/null | grep -c 'BPT.TRAP') trap instructions in $(basename "$LIB") (product library: $(cuobjdump -sass manual_yolo_b200/libb200yolo.so 2>/dev/null | grep -c 'BPT.TRAP'))";
      B200YOLO_LIB="$LIB" python -c "from manual_yolo_b200 import _lib; _lib.load(); print('library under test:', _lib.LIB_PATH)" 2>&1 | tail -1;
      B200YOLO_LIB="$LIB" timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6;
      B200YOLO_LIB="$LIB" timeout 600 python tools/sanitize_cases.py 2>&1 | tail -2; } | tee gpurun_out/checked_r02.log
    ;;
esac
