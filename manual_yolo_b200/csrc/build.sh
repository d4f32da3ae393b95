#!/usr/bin/env bash
# Build libb200yolo.so in-tree for sm_100a.  Usage: manual_yolo_b200/csrc/build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I$ROOT/include $ARCH"
OUT="${B200YOLO_OUT:-$ROOT/manual_yolo_b200/libb200yolo.so}"     # (tuning experiments build variants side by side)
OBJ="${B200YOLO_OBJ:-$HERE/_obj}"
mkdir -p "$OBJ"
# parity-critical fp32 kernels: no FMA contraction (every add/mul rounds like the torch CPU ops)
for f in decode_filter nms postprocess_small assoc slices track; do
  "$NVCC" $COMMON -fmad=false "$@" -c "$HERE/$f.cu" -o "$OBJ/$f.o" &
done
for f in abi letterbox sort_topk roi; do
  "$NVCC" $COMMON "$@" -c "$HERE/$f.cu" -o "$OBJ/$f.o" &
done
wait
"$NVCC" -shared $ARCH -o "$OUT" "$OBJ"/{abi,letterbox,decode_filter,sort_topk,nms,roi,postprocess_small,slices,assoc,track}.o
echo "built $OUT"
