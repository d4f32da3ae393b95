// Helpers shared by nms.cu, sort_topk.cu and postprocess_small.cu (the general and the fused path must
// apply the SAME arithmetic and tie rules; tests assert their outputs are bit-identical).
#pragma once

#include "common.cuh"

// sort_topk.cu: pass 0 = the regular per-image sort (dense regime: only the best 2048 entries are ordered, the count
// goes to the workspace header), pass 1 = full sort of the images whose fallback flag the NMS raised.
int b200_sort_launch(const float* cand, const int* cand_anchor, const int* cand_count, int B, int cap, int max_nms,
                     int* order, void* workspace, size_t workspace_bytes, int pass, cudaStream_t s);

// decode_filter.cu: DFL box decode of the ordered candidates only (pass 0) / of everything for flagged images (pass 1)
int b200_box_decode_sorted_launch(const b200yolo_level* levels, int n_levels, float* cand, const int* cand_anchor,
                                  const int* cand_count, const int* order, int B, int cap, int max_nms, const int* hdr,
                                  int pass, cudaStream_t s);

namespace b200 {

// Sort key: ascending key order == score descending, anchor ascending on ties (the reference's stable
// descending sort over anchor-ordered rows); the low 16 bits carry the candidate slot.
__device__ __forceinline__ uint64_t make_key(float score, int anchor, int slot) {
  const uint32_t sb = ~__float_as_uint(score);
  return ((uint64_t)sb << 32) | ((uint64_t)(uint32_t)(anchor & 0xffff) << 16) | (uint32_t)(slot & 0xffff);
}

// torchvision nms_kernel (CPU): inter = max(0,xx2-xx1) * max(0,yy2-yy1); ovr = inter / (ai + aj - inter);
// suppress iff ovr > thr with the fp32 ovr promoted to double.  0/0 = NaN never suppresses.
__device__ __forceinline__ bool iou_suppresses(const float4 a, const float area_a, const float4 b, const double thr) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  if (!(inter > 0.f)) return false;  // ovr is 0, -0 or NaN: never > thr (thr >= 0)
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return (double)ovr > thr;
}

// Exact shortcut of the class-offset trick: when both raw boxes lie inside [-max_wh/2, max_wh/2) on every
// coordinate, boxes of different classes occupy disjoint offset ranges [(c-1/2)*max_wh, (c+1/2)*max_wh]
// (end points exact in fp32, rounding is monotone), so inter == 0 and the pair can never suppress.
// meta = class id for such boxes, -1 otherwise (agnostic mode, out-of-range or non-integral class, NaN):
// pairs involving a -1 always take the full IoU test.
__device__ __forceinline__ int box_meta(const float* row, float max_wh, int agnostic) {
  if (agnostic) return -1;
  const float c = row[5];
  const float hw = 0.5f * max_wh;
  const bool inb = row[0] >= -hw && row[0] < hw && row[1] >= -hw && row[1] < hw && row[2] >= -hw && row[2] < hw &&
                   row[3] >= -hw && row[3] < hw;
  // (c +- 1/2)*max_wh must be exact in fp32: integral class, even integral max_wh, products < 2^24
  if (!inb || !(c >= 0.f) || c != floorf(c) || hw != floorf(hw) || (c + 1.f) * max_wh >= 16777216.f) return -1;
  return (int)c;
}
__device__ __forceinline__ bool may_overlap(int mi, int mj) { return mi == mj || (mi | mj) < 0; }

// ops.scale_boxes + clip_boxes on one coordinate: (x - pad) / gain, clamp to [0, limit]
__device__ __forceinline__ float scale_clip(float x, float pad, float gain, float limit) {
  return fminf(fmaxf(__fdiv_rn(__fsub_rn(x, pad), gain), 0.f), limit);
}

}  // namespace b200
