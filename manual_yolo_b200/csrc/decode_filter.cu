// K2: YOLOv8 Detect-head decode + confidence filter + compaction.
//
// Replaces ultralytics Detect._inference (DFL softmax expectation over 16 bins, dist2bbox(xywh) *
// stride, class sigmoid; nn/modules/head.py, block.py, utils/tal.py) fused with the head of
// ops.non_max_suppression (amax > conf, xywh2xyxy, cls.max, conf filter, classes filter) -- reference
// entry detect.py:541 / yolo.py:361 / pipe.py:179.  Restated in oracle/head.py + oracle/nms.py.
//
// One thread per anchor.  The head is channel-major with the anchor axis contiguous, so a warp's
// load of one channel is one coalesced 128-byte line; channels are streamed with an unrolled loop
// (16 independent loads in flight per thread).  Work is lazy: the class channels are reduced to the
// best logit first, only anchors whose best sigmoid beats conf_thres (about 1% at conf 0.25, all of
// them at conf 0.001) read and decode the 64 DFL channels.  Survivors are compacted with a warp
// ballot and one atomicAdd per warp on the per-image counter.
// HBM-bound: algorithmic bytes per frame = (64+nc)*A*4 read + 28 B per survivor written.
// Compiled with -fmad=false: every add/mul below rounds separately, as the torch CPU ops do.

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kReg = B200YOLO_REG_MAX;

struct Levels {
  const float* ptr[B200YOLO_MAX_LEVELS];
  long long bstride[B200YOLO_MAX_LEVELS];
  long long cstride[B200YOLO_MAX_LEVELS];
  int w[B200YOLO_MAX_LEVELS];
  int off[B200YOLO_MAX_LEVELS + 1];  // anchor offsets, off[n_levels] = A
  float stride[B200YOLO_MAX_LEVELS];
  int n;
};

__device__ __forceinline__ float sigmoid_f32(float x) { return b200::sigmoid_torch(x); }

__device__ __forceinline__ bool class_allowed(const uint32_t* mask, int c) {
  return mask == nullptr || ((mask[c >> 5] >> (c & 31)) & 1u);
}

// Emit one candidate row; warp-aggregated slot allocation.  All 32 lanes must call.
__device__ __forceinline__ void emit(bool is_cand, int b, int a, float x1, float y1, float x2, float y2,
                                     float score, int cls, float* cand, int* cand_anchor, int* cand_count,
                                     int cap) {
  const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
  if (ballot == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == (__ffs(ballot) - 1)) base = atomicAdd(&cand_count[b], __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (is_cand) {
    const int slot = base + __popc(ballot & ((1u << lane) - 1u));
    if (slot < cap) {
      float2* row = reinterpret_cast<float2*>(cand + ((int64_t)b * cap + slot) * 6);
      row[0] = make_float2(x1, y1);
      row[1] = make_float2(x2, y2);
      row[2] = make_float2(score, (float)cls);
      cand_anchor[(int64_t)b * cap + slot] = a;
    }
  }
}

__global__ void __launch_bounds__(kThreads) decode_filter_kernel(const Levels L, int nc, float conf,
                                                                 const uint32_t* __restrict__ class_mask,
                                                                 float* __restrict__ cand,
                                                                 int* __restrict__ cand_anchor,
                                                                 int* __restrict__ cand_count, int cap) {
  const int A = L.off[B200YOLO_MAX_LEVELS];
  const int a = blockIdx.x * kThreads + threadIdx.x;
  const int b = blockIdx.y;
  const bool live = a < A;
  // level lookup with static indices only (dynamic indexing would spill the params to local memory)
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
  const int i = live ? a - off : 0;
  const float* p = base + (long long)b * bs + i;

  // ---- class channels: best logit m (first index j), and m2 = best logit among indices < j ----
  float m = -INFINITY, m2 = -INFINITY;
  int j = 0;
  if (live) {
    const float* pc = p + (long long)(4 * kReg) * cs;
    int c = 0;
    for (; c + 16 <= nc; c += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = b200::ldg_stream(pc + (long long)(c + u) * cs);
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (v[u] > m) { m2 = m; m = v[u]; j = c + u; }
    }
    for (; c < nc; ++c) {
      float v = b200::ldg_stream(pc + (long long)c * cs);
      if (v > m) { m2 = m; m = v; j = c; }
    }
  }
  const float score = sigmoid_f32(m);
  bool is_cand = live && (score > conf);

  float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
  if (is_cand) {
    // cls.max(1) returns the LOWEST index whose sigmoid equals the maximum: an earlier class with a
    // smaller logit can still tie after rounding/saturation.  m2 bounds every earlier logit, so the
    // rescan is needed only when sigmoid(m2) == sigmoid(m) (practically never).
    if (m2 > -INFINITY && sigmoid_f32(m2) == score) {
      const float* pc = p + (long long)(4 * kReg) * cs;
      for (int c = 0; c < j; ++c) {
        if (sigmoid_f32(pc[(long long)c * cs]) == score) { j = c; break; }
      }
    }
    if (!class_allowed(class_mask, j)) is_cand = false;
  }
  if (is_cand) {
    // ---- DFL: softmax over 16 bins per side, expectation with weights 0..15 (sequential) ----
    float d[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float v[kReg];
#pragma unroll
      for (int k = 0; k < kReg; ++k) v[k] = p[(long long)(s * kReg + k) * cs];
      float mx = v[0];
#pragma unroll
      for (int k = 1; k < kReg; ++k) mx = fmaxf(mx, v[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < kReg; ++k) { v[k] = b200::expf_torch(__fsub_rn(v[k], mx)); sum = __fadd_rn(sum, v[k]); }
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kReg; ++k) acc = __fmaf_rn((float)k, __fdiv_rn(v[k], sum), acc);
      d[s] = acc;
    }
    // ---- dist2bbox(xywh=True) * stride, then xywh2xyxy, in the reference's op order ----
    const int w = lw;
    const float ax = (float)(i % w) + 0.5f, ay = (float)(i / w) + 0.5f;
    const float bx1 = ax - d[0], by1 = ay - d[1], bx2 = ax + d[2], by2 = ay + d[3];
    const float cx = ((bx1 + bx2) / 2.0f) * st, cy = ((by1 + by2) / 2.0f) * st;
    const float bw = (bx2 - bx1) * st, bh = (by2 - by1) * st;
    const float hw = bw / 2.0f, hh = bh / 2.0f;
    x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
  }
  emit(is_cand, b, a, x1, y1, x2, y2, score, j, cand, cand_anchor, cand_count, cap);
}

// Already-decoded UL prediction (B, channels, A): rows 0-3 xywh, rows 4..4+nc scores.
__global__ void __launch_bounds__(kThreads) filter_decoded_kernel(const float* __restrict__ pred, int channels,
                                                                  int nc, int A, float conf,
                                                                  const uint32_t* __restrict__ class_mask,
                                                                  float* __restrict__ cand,
                                                                  int* __restrict__ cand_anchor,
                                                                  int* __restrict__ cand_count, int cap) {
  const int a = blockIdx.x * kThreads + threadIdx.x;
  const int b = blockIdx.y;
  const bool live = a < A;
  const float* p = pred + ((long long)b * channels) * A + a;
  float m = -INFINITY;
  int j = 0;
  if (live) {
    const float* pc = p + 4LL * A;
    int c = 0;
    for (; c + 16 <= nc; c += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = b200::ldg_stream(pc + (long long)(c + u) * A);
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (v[u] > m) { m = v[u]; j = c + u; }
    }
    for (; c < nc; ++c) {
      float v = b200::ldg_stream(pc + (long long)c * A);
      if (v > m) { m = v; j = c; }
    }
  }
  bool is_cand = live && (m > conf) && class_allowed(class_mask, j);
  float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
  if (is_cand) {
    const float cx = p[0], cy = p[A], bw = p[2LL * A], bh = p[3LL * A];
    const float hw = bw / 2.0f, hh = bh / 2.0f;
    x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
  }
  emit(is_cand, b, a, x1, y1, x2, y2, m, j, cand, cand_anchor, cand_count, cap);
}

}  // namespace

extern "C" int b200yolo_decode_filter(const b200yolo_level* levels, int n_levels, int B, int nc, float conf_thres,
                                      const uint32_t* class_mask, float* cand, int* cand_anchor,
                                      int* cand_count, int cap, void* stream) {
  B200_REQUIRE(levels && cand && cand_anchor && cand_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(n_levels >= 1 && n_levels <= B200YOLO_MAX_LEVELS, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(B > 0 && B <= 65535 && nc > 0 && cap > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(nc <= B200YOLO_MAX_CLASSES, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f, B200YOLO_ERR_RANGE);
  Levels L;
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
  B200_REQUIRE(off <= (1LL << 30), B200YOLO_ERR_UNSUPPORTED);
  for (int l = n_levels; l <= B200YOLO_MAX_LEVELS; ++l) L.off[l] = (int)off;
  dim3 grid((unsigned)((off + kThreads - 1) / kThreads), B);
  decode_filter_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(L, nc, conf_thres, class_mask, cand,
                                                                     cand_anchor, cand_count, cap);
  return b200_launch_status();
}

extern "C" int b200yolo_filter_decoded(const float* pred, int B, int channels, int nc, int A, float conf_thres,
                                       const uint32_t* class_mask, float* cand, int* cand_anchor,
                                       int* cand_count, int cap, void* stream) {
  B200_REQUIRE(pred && cand && cand_anchor && cand_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && B <= 65535 && nc > 0 && A > 0 && cap > 0 && channels >= 4 + nc, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(nc <= B200YOLO_MAX_CLASSES, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f, B200YOLO_ERR_RANGE);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 3) == 0, B200YOLO_ERR_ALIGN);
  dim3 grid((unsigned)((A + kThreads - 1) / kThreads), B);
  filter_decoded_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(pred, channels, nc, A, conf_thres,
                                                                      class_mask, cand, cand_anchor, cand_count,
                                                                      cap);
  return b200_launch_status();
}
