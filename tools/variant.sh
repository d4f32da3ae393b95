#!/usr/bin/env bash
# Build a tuning variant of libb200yolo.so: tools/variant.sh <name> <file-stem> [nvcc -D flags...]
# Only <file-stem>.cu is recompiled (with the flags); every other object comes from the regular build.
# Output: scratch/variants/lib_<name>.so (git-ignored; travels to the GPU box).  Use with B200YOLO_LIB=...
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
NAME="$1"; STEM="$2"; shift 2
CS="$ROOT/manual_yolo_b200/csrc"; OBJ="$CS/_obj"; V="$ROOT/scratch/variants"
mkdir -p "$V/obj_$NAME"
ARCH="-gencode arch=compute_100a,code=sm_100a"
FMAD=""
case "$STEM" in decode_filter|nms|postprocess_small|assoc|slices|track) FMAD="-fmad=false";; esac
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I"$ROOT/include" $ARCH $FMAD "$@" -c "$CS/$STEM.cu" -o "$V/obj_$NAME/$STEM.o"
OBJS=()
for f in abi letterbox decode_filter sort_topk nms roi postprocess_small slices assoc track; do
  if [ "$f" = "$STEM" ]; then OBJS+=("$V/obj_$NAME/$f.o"); else OBJS+=("$OBJ/$f.o"); fi
done
/usr/local/cuda/bin/nvcc -shared $ARCH -o "$V/lib_$NAME.so" "${OBJS[@]}"
echo "built $V/lib_$NAME.so"
