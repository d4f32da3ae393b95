// Shared by decode_filter.cu and postprocess_small.cu: Detect-head level description and the DFL box decode.
#pragma once

#include "common.cuh"

namespace b200 {

constexpr int kReg = B200YOLO_REG_MAX;

struct Levels {  // by-value kernel parameter; index with static indices only
  const float* ptr[B200YOLO_MAX_LEVELS];
  long long bstride[B200YOLO_MAX_LEVELS];
  long long cstride[B200YOLO_MAX_LEVELS];
  int w[B200YOLO_MAX_LEVELS];
  int off[B200YOLO_MAX_LEVELS + 1];  // anchor offsets, off[n_levels] = A
  float stride[B200YOLO_MAX_LEVELS];
  int n;
};

// Host: validate the C-ABI level array and flatten it.  Returns a B200YOLO_* status.
static inline int build_levels(const b200yolo_level* levels, int n_levels, Levels& L) {
  B200_REQUIRE(levels, B200YOLO_ERR_NULL);
  B200_REQUIRE(n_levels >= 1 && n_levels <= B200YOLO_MAX_LEVELS, B200YOLO_ERR_SHAPE);
  L.n = n_levels;
  long long off = 0;
  for (int l = 0; l < B200YOLO_MAX_LEVELS; ++l) {
    if (l < n_levels) {
      B200_REQUIRE(levels[l].ptr, B200YOLO_ERR_NULL);
      B200_REQUIRE(levels[l].h > 0 && levels[l].w > 0 && levels[l].stride > 0.f, B200YOLO_ERR_SHAPE);
      B200_REQUIRE((reinterpret_cast<uintptr_t>(levels[l].ptr) & 3) == 0, B200YOLO_ERR_ALIGN);
      L.ptr[l] = levels[l].ptr; L.bstride[l] = levels[l].batch_stride; L.cstride[l] = levels[l].chan_stride;
      L.w[l] = levels[l].w; L.stride[l] = levels[l].stride; L.off[l] = (int)off;
      off += (long long)levels[l].h * levels[l].w;
    } else {
      L.ptr[l] = nullptr; L.bstride[l] = 0; L.cstride[l] = 0; L.w[l] = 1; L.stride[l] = 1.f; L.off[l] = (int)off;
    }
  }
  // the sort key carries the anchor index in 16 bits (nms_common.cuh make_key): more anchors than that would break
  // the anchor-order tie rule and the uniqueness of keys the select-sort relies on
  B200_REQUIRE(off <= B200YOLO_MAX_ANCHORS, B200YOLO_ERR_UNSUPPORTED);
  for (int l = n_levels; l <= B200YOLO_MAX_LEVELS; ++l) L.off[l] = (int)off;
  return B200YOLO_OK;
}

// Per-anchor view of the head: channel-0 element pointer, channel stride, grid width, stride, local index.
struct AnchorRef {
  const float* p;
  long long cs;
  int lw, i;
  float st;
};
__device__ __forceinline__ AnchorRef anchor_ref(const Levels& L, int b, int a) {
  int off = 0, lw = L.w[0];
  long long cs = L.cstride[0], bs = L.bstride[0];
  const float* base = L.ptr[0];
  float st = L.stride[0];
#pragma unroll
  for (int l = 1; l < B200YOLO_MAX_LEVELS; ++l) {
    if (l < L.n && a >= L.off[l]) {
      off = L.off[l]; lw = L.w[l]; cs = L.cstride[l]; bs = L.bstride[l]; base = L.ptr[l]; st = L.stride[l];
    }
  }
  AnchorRef r;
  r.i = a - off; r.lw = lw; r.cs = cs; r.st = st;
  r.p = base + (long long)b * bs + r.i;
  return r;
}

// DFL.forward for one box side: softmax over the 16 bins, then torch's 1x1 conv with arange weights
// (sequential fma), bit-identical to the torch CPU ops (expf_torch restates Sleef's expf_u10).
__device__ __forceinline__ float dfl_side(const float* __restrict__ p, long long cs) {
  float v[kReg];
#pragma unroll
  for (int k = 0; k < kReg; ++k) v[k] = p[(long long)k * cs];
  float mx = v[0], mn = v[0];
#pragma unroll
  for (int k = 1; k < kReg; ++k) { mx = fmaxf(mx, v[k]); mn = fminf(mn, v[k]); }
  float sum = 0.f, acc = 0.f;
  // softmax = exp(x - max) / sum with torch's Sleef expf and a true division.  sum is in [1, 16] (the maximum
  // contributes exactly 1).  When every logit is within 60 of the maximum (always, for a trained head) the cheap
  // bit-identical forms apply: exponent-add scaling and 16 quotients from one correctly rounded reciprocal.
  // Otherwise (tiny or subnormal exps; NaN/Inf logits fail the comparison) the general forms run.
  if (__fsub_rn(mn, mx) >= -60.0f) {
#pragma unroll
    for (int k = 0; k < kReg; ++k) { v[k] = expf_torch_m80_0(__fsub_rn(v[k], mx)); sum = __fadd_rn(sum, v[k]); }
    const float r = __frcp_rn(sum);
#pragma unroll
    for (int k = 0; k < kReg; ++k) acc = __fmaf_rn((float)k, div_by_rcp(v[k], sum, r), acc);
  } else {
#pragma unroll
    for (int k = 0; k < kReg; ++k) { v[k] = expf_torch(__fsub_rn(v[k], mx)); sum = __fadd_rn(sum, v[k]); }
#pragma unroll
    for (int k = 0; k < kReg; ++k) acc = __fmaf_rn((float)k, __fdiv_rn(v[k], sum), acc);
  }
  return acc;
}

// dist2bbox(xywh=True) * stride, then xywh2xyxy, in the reference's op order (compile with -fmad=false).
__device__ __forceinline__ float4 decode_box(const AnchorRef& r, float d0, float d1, float d2, float d3) {
  const float ax = (float)(r.i % r.lw) + 0.5f, ay = (float)(r.i / r.lw) + 0.5f;
  const float bx1 = ax - d0, by1 = ay - d1, bx2 = ax + d2, by2 = ay + d3;
  const float cx = ((bx1 + bx2) / 2.0f) * r.st, cy = ((by1 + by2) / 2.0f) * r.st;
  const float bw = (bx2 - bx1) * r.st, bh = (by2 - by1) * r.st;
  const float hw = bw / 2.0f, hh = bh / 2.0f;
  return make_float4(cx - hw, cy - hh, cx + hw, cy + hh);
}

}  // namespace b200
