// K2: YOLOv8 Detect-head decode + confidence filter + compaction.
//
// Replaces ultralytics Detect._inference (DFL softmax expectation over 16 bins, dist2bbox(xywh) *
// stride, class sigmoid; nn/modules/head.py, block.py, utils/tal.py) fused with the head of
// ops.non_max_suppression (amax > conf, xywh2xyxy, cls.max, conf filter, classes filter) -- reference
// entry detect.py:541 / yolo.py:361 / pipe.py:179.  Restated in oracle/head.py + oracle/nms.py.
//
// Two kernels make up the fast path (see the block comment above decode_vec_kernel):
//   decode_vec_kernel  streams the class channels with 128-bit loads (4 anchors per lane, channel
//                      quarters per warp pair), thresholds the best sigmoid and compacts the survivors;
//   box_decode_kernel  decodes the 64 DFL channels of the survivors only (4 lanes per box, lane = side).
// Work is lazy: about 1% of the anchors survive conf 0.25, so the box channels of the other 99% are never
// read.  decode_filter_kernel (first in this file) is the scalar fallback for shapes that are not
// 16-byte aligned: 8 warps cover 64 anchors, lane = anchor, warp q reduces channel quarter q and, for the
// survivors, decodes box side q.
// HBM-bound: algorithmic bytes per frame = (64+nc)*A*4 read + 28 B per survivor written.
// Compiled with -fmad=false: every add/mul below rounds separately, as the torch CPU ops do.

#include "levels.cuh"
#include "nms_common.cuh"

namespace {

constexpr int kThreads = 256;
using b200::kReg;
using b200::Levels;

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

// Survivor whose box is decoded later (raw-head fast path): slot, score, class, anchor only.
__device__ __forceinline__ void emit_pending(bool is_cand, int b, int a, float score, int cls, float* cand,
                                             int* cand_anchor, int* cand_count, int cap) {
  const unsigned ballot = __ballot_sync(0xffffffffu, is_cand);
  if (ballot == 0) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == (__ffs(ballot) - 1)) base = atomicAdd(&cand_count[b], __popc(ballot));
  base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
  if (is_cand) {
    const int slot = base + __popc(ballot & ((1u << lane) - 1u));
    if (slot < cap) {
      reinterpret_cast<float2*>(cand + ((int64_t)b * cap + slot) * 6)[2] = make_float2(score, (float)cls);
      cand_anchor[(int64_t)b * cap + slot] = a;
    }
  }
}

constexpr int kQ = 4;                       // warps per anchor group (channel quarters / box sides)
constexpr int kGroups = kThreads / 32 / kQ; // anchor groups per CTA (2)
constexpr int kAnchorsPerCta = kGroups * 32;

struct Part { float m, m2; int j; };

// RAW = true : head (B, 64+nc, A) per level -> DFL decode;  RAW = false: decoded prediction (B, ch, A).
template <bool RAW>
__global__ void __launch_bounds__(kThreads) decode_filter_kernel(const Levels L, const float* __restrict__ pred,
                                                                 int channels, int nc, float conf,
                                                                 const uint32_t* __restrict__ class_mask,
                                                                 float* __restrict__ cand,
                                                                 int* __restrict__ cand_anchor,
                                                                 int* __restrict__ cand_count, int cap, int defer) {
  __shared__ Part part[kGroups][kQ][32];
  __shared__ float dist[kGroups][kQ][32];
  __shared__ unsigned char flag[kGroups][32];
  const int A = L.off[B200YOLO_MAX_LEVELS];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = wid / kQ, q = wid % kQ;
  const int a = (blockIdx.x * kGroups + g) * 32 + lane;
  const int b = blockIdx.y;
  const bool live = a < A;
  // per-lane level lookup with static indices only (dynamic indexing would spill the params)
  int off = 0, lw = L.w[0];
  long long cs = L.cstride[0], bs = L.bstride[0];
  const float* base = L.ptr[0];
  float st = L.stride[0];
  if (RAW) {
#pragma unroll
    for (int l = 1; l < B200YOLO_MAX_LEVELS; ++l) {
      if (l < L.n && a >= L.off[l]) {
        off = L.off[l]; lw = L.w[l]; cs = L.cstride[l]; bs = L.bstride[l]; base = L.ptr[l]; st = L.stride[l];
      }
    }
  } else {
    cs = A; bs = (long long)channels * A; base = pred;
  }
  const int i = live ? a - off : 0;
  const float* p = base + (long long)b * bs + i;
  const float* pc = p + (long long)(RAW ? 4 * kReg : 4) * cs;   // first class channel

  // ---- phase 1: this warp's quarter of the class channels, one load round ----
  const int cq = (nc + kQ - 1) / kQ;
  const int c0 = q * cq, c1 = min(nc, c0 + cq);
  float m = -INFINITY, m2 = -INFINITY;
  int j = c0;
  if (live) {
    int c = c0;
    for (; c + 16 <= c1; c += 16) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = b200::ldg_stream(pc + (long long)(c + u) * cs);
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (v[u] > m) { m2 = m; m = v[u]; j = c + u; }
    }
    if (c < c1) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (c + u < c1) ? b200::ldg_stream(pc + (long long)(c + u) * cs) : -INFINITY;
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (v[u] > m) { m2 = m; m = v[u]; j = c + u; }
    }
  }
  part[g][q][lane] = Part{m, m2, j};
  __syncthreads();

  // ---- phase 2: merge quarters in class order (first maximum wins), threshold, class filter ----
  float score = 0.f;
  int cls = 0;
  bool is_cand = false;
  if (q == 0) {
    Part t = part[g][0][lane];
    m = t.m; m2 = t.m2; j = t.j;
#pragma unroll
    for (int k = 1; k < kQ; ++k) {
      t = part[g][k][lane];
      if (t.m > m) { m2 = fmaxf(m, t.m2); m = t.m; j = t.j; }   // everything before t.j: earlier quarters + t.m2
    }
    score = RAW ? b200::sigmoid_torch(m) : m;
    is_cand = live && (score > conf);
    cls = j;
    if (RAW && is_cand && m2 > -INFINITY && b200::sigmoid_torch(m2) == score) {
      // cls.max(1) returns the LOWEST index whose sigmoid equals the maximum: an earlier class with a
      // smaller logit can still tie after rounding/saturation (practically never: rescan is rare).
      for (int c = 0; c < j; ++c)
        if (b200::sigmoid_torch(pc[(long long)c * cs]) == score) { cls = c; break; }
    }
    if (is_cand && !class_allowed(class_mask, cls)) is_cand = false;
    flag[g][lane] = is_cand;
  }
  if (!__syncthreads_or(is_cand)) return;   // no survivor among this CTA's 64 anchors (the common case)
  if (RAW && defer) {                        // boxes decoded later (b200yolo_class_filter)
    if (q == 0) emit_pending(is_cand, b, a, score, cls, cand, cand_anchor, cand_count, cap);
    return;
  }

  float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
  if (RAW) {
    // ---- phase 3: warp q decodes box side q for the group's surviving lanes (DFL expectation) ----
    if (flag[g][lane]) {
      float v[kReg];
#pragma unroll
      for (int k = 0; k < kReg; ++k) v[k] = p[(long long)(q * kReg + k) * cs];
      float mx = v[0];
#pragma unroll
      for (int k = 1; k < kReg; ++k) mx = fmaxf(mx, v[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < kReg; ++k) { v[k] = b200::expf_torch(__fsub_rn(v[k], mx)); sum = __fadd_rn(sum, v[k]); }
      float acc = 0.f;   // torch's 1x1 conv with arange weights: sequential fma over the 16 bins
#pragma unroll
      for (int k = 0; k < kReg; ++k) acc = __fmaf_rn((float)k, __fdiv_rn(v[k], sum), acc);
      dist[g][q][lane] = acc;
    }
    __syncthreads();
    if (q != 0) return;
    if (is_cand) {
      // dist2bbox(xywh=True) * stride, then xywh2xyxy, in the reference's op order
      const float d0 = dist[g][0][lane], d1 = dist[g][1][lane], d2 = dist[g][2][lane], d3 = dist[g][3][lane];
      const float ax = (float)(i % lw) + 0.5f, ay = (float)(i / lw) + 0.5f;
      const float bx1 = ax - d0, by1 = ay - d1, bx2 = ax + d2, by2 = ay + d3;
      const float cx = ((bx1 + bx2) / 2.0f) * st, cy = ((by1 + by2) / 2.0f) * st;
      const float bw = (bx2 - bx1) * st, bh = (by2 - by1) * st;
      const float hw = bw / 2.0f, hh = bh / 2.0f;
      x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
    }
  } else {
    if (q != 0) return;
    if (is_cand) {
      const float cx = p[0], cy = p[cs], bw = p[2 * cs], bh = p[3 * cs];
      const float hw = bw / 2.0f, hh = bh / 2.0f;
      x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
    }
  }
  emit(is_cand, b, a, x1, y1, x2, y2, score, cls, cand, cand_anchor, cand_count, cap);
}

// =================================================================================================
// Vectorised variant (the fast path; the kernel above is the fallback for shapes that are not
// 16-byte aligned).  A CTA of 256 threads covers 256 consecutive anchors of one flat segment (the whole
// concatenated head, or one level tensor): thread = (channel quarter q, anchor quad), i.e. every load is
// a 128-bit LDG of 4 consecutive anchors of one class channel, 8 of them in flight per thread, with
// CTA-uniform strides so the address arithmetic is one pointer bump per load.  The partial
// (max, runner-up, argmax) triples go through shared memory (SoA, conflict-free) and thread t then owns
// anchor t: merge in class order, sigmoid, threshold.  Lazy DFL: only surviving anchors read their 64
// box channels (coalesced across the owners of a dense tile), softmax expectation per side, decode, emit.
// =================================================================================================
// Tuning hooks (defaults = the measured best; sweep recorded in profiles/README.md)
#ifndef B200_K2_T
#define B200_K2_T 256             // threads per CTA = anchors per CTA
#endif
#ifndef B200_K2_MINB
#define B200_K2_MINB 4            // resident CTAs per SM the register budget is held to
#endif
#ifndef B200_K2_INFLIGHT
#define B200_K2_INFLIGHT 8        // 128-bit loads issued per thread before the first compare
#endif
constexpr int kVA = B200_K2_T;    // anchors per CTA
constexpr int kVL = B200_K2_INFLIGHT;

struct FlatSegs {
  const float* ptr[B200YOLO_MAX_LEVELS];
  long long bstride[B200YOLO_MAX_LEVELS], cstride[B200YOLO_MAX_LEVELS];
  int count[B200YOLO_MAX_LEVELS], first[B200YOLO_MAX_LEVELS];
  int n;
};

template <bool RAW>
__global__ void __launch_bounds__(kVA, B200_K2_MINB) decode_vec_kernel(const FlatSegs S, const Levels L, int nc, int cls0,
                                                            float conf, const uint32_t* __restrict__ class_mask,
                                                            float* __restrict__ cand, int* __restrict__ cand_anchor,
                                                            int* __restrict__ cand_count, int cap) {
  __shared__ __align__(16) float pm[kQ][kVA], pm2[kQ][kVA];
  __shared__ __align__(16) int pj[kQ][kVA];
  const int sg = blockIdx.z, b = blockIdx.y, tid = threadIdx.x;
  // CTA-uniform segment parameters (static indexing only)
  const float* sptr = S.ptr[0];
  long long sbs = S.bstride[0], scs = S.cstride[0];
  int scount = S.count[0], sfirst = S.first[0];
#pragma unroll
  for (int l = 1; l < B200YOLO_MAX_LEVELS; ++l)
    if (sg == l) { sptr = S.ptr[l]; sbs = S.bstride[l]; scs = S.cstride[l]; scount = S.count[l]; sfirst = S.first[l]; }
  const int a_cta = blockIdx.x * kVA;
  if (a_cta >= scount) return;

  // ---- phase 1: thread = (quarter q, 4 consecutive anchors); 128-bit loads, 8 in flight ----
  {
    const int q = tid / (kVA / 4), t4 = (tid % (kVA / 4)) * 4;
    const bool live = a_cta + t4 < scount;                    // segment counts are multiples of 4
    const int cq = (nc + kQ - 1) / kQ;
    const int c0 = q * cq, c1 = min(nc, c0 + cq);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    float m2[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int j[4] = {c0, c0, c0, c0};
    if (live) {
      const float* p = sptr + (long long)b * sbs + (long long)(cls0 + c0) * scs + a_cta + t4;
      for (int c = c0; c < c1; c += kVL) {
        float4 v[kVL];
#pragma unroll
        for (int u = 0; u < kVL; ++u) {
          if (c + u < c1) {
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                         : "l"(p + (long long)u * scs));
          } else {
            v[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          }
        }
        p += kVL * scs;
#pragma unroll
        for (int u = 0; u < kVL; ++u) {
          if (v[u].x > m[0]) { m2[0] = m[0]; m[0] = v[u].x; j[0] = c + u; }
          if (v[u].y > m[1]) { m2[1] = m[1]; m[1] = v[u].y; j[1] = c + u; }
          if (v[u].z > m[2]) { m2[2] = m[2]; m[2] = v[u].z; j[2] = c + u; }
          if (v[u].w > m[3]) { m2[3] = m[3]; m[3] = v[u].w; j[3] = c + u; }
        }
      }
    }
    *reinterpret_cast<float4*>(&pm[q][t4]) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(&pm2[q][t4]) = make_float4(m2[0], m2[1], m2[2], m2[3]);
    *reinterpret_cast<int4*>(&pj[q][t4]) = make_int4(j[0], j[1], j[2], j[3]);
  }
  __syncthreads();

  // ---- phase 2: thread t owns anchor t: merge quarters in class order, threshold ----
  const int a_loc = a_cta + tid;
  const int a = sfirst + a_loc;                               // global anchor index
  const bool live = a_loc < scount;
  float m = pm[0][tid], m2 = pm2[0][tid];
  int j = pj[0][tid];
#pragma unroll
  for (int k = 1; k < kQ; ++k) {
    const float tm = pm[k][tid];
    if (tm > m) { m2 = fmaxf(m, pm2[k][tid]); m = tm; j = pj[k][tid]; }
  }
  const float score = RAW ? b200::sigmoid_torch(m) : m;
  bool is_cand = live && (score > conf);
  if (!__syncthreads_or(is_cand)) return;                     // no survivor among these 256 anchors

  // per-lane level parameters (decode geometry) and this anchor's channel-0 element
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
  int cls = j;
  if (RAW && is_cand && m2 > -INFINITY && b200::sigmoid_torch(m2) == score) {
    const float* pc = p + (long long)cls0 * cs;               // lowest class index whose sigmoid ties the maximum
    for (int c = 0; c < j; ++c)
      if (b200::sigmoid_torch(pc[(long long)c * cs]) == score) { cls = c; break; }
  }
  if (is_cand && !class_allowed(class_mask, cls)) is_cand = false;

  if (RAW) {
    // boxes are decoded by box_decode_kernel from the compacted list: the streaming CTAs never wait on
    // the (rare, latency-bound) DFL reads
    emit_pending(is_cand, b, a, score, cls, cand, cand_anchor, cand_count, cap);
  } else {
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (is_cand) {
      const float cx = p[0], cy = p[cs], bw = p[2 * cs], bh = p[3 * cs];
      const float hw = bw / 2.0f, hh = bh / 2.0f;
      x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
    }
    emit(is_cand, b, a, x1, y1, x2, y2, score, cls, cand, cand_anchor, cand_count, cap);
  }
}

// Second half of the raw-head path: DFL decode of the compacted survivors.  Four lanes per candidate
// (lane = box side): 16 bins each in one load round, softmax expectation exactly as torch computes it,
// quad shuffle, then dist2bbox(xywh) * stride and xywh2xyxy in the reference's op order.
__global__ void __launch_bounds__(256, 6) box_decode_kernel(const Levels L, float* __restrict__ cand,
                                                            const int* __restrict__ cand_anchor,
                                                            const int* __restrict__ cand_count, int cap, int B,
                                                            int per_image) {
  const int lane = threadIdx.x & 31, sd = threadIdx.x & 3;
  // flat grid-stride over (image, block of 64 survivors) pairs: the grid is one full wave of resident CTAs
  // whatever B is, and work stays balanced when images hold different numbers of survivors
  const int total = B * per_image;
  for (int f = blockIdx.x; f < total; f += gridDim.x) {
    const int b = f / per_image, blk = f - b * per_image;
    const int n = min(cand_count[b], cap);
    if (blk * 64 >= n) continue;
    const int slot = blk * 64 + (threadIdx.x >> 2);
    const bool act = slot < n;
    const int a = act ? cand_anchor[(int64_t)b * cap + slot] : 0;
    const b200::AnchorRef r = b200::anchor_ref(L, b, a);
    const float d = act ? b200::dfl_side(r.p + (long long)(sd * kReg) * r.cs, r.cs) : 0.f;
    const int q0 = lane & ~3;
    const float d0 = __shfl_sync(0xffffffffu, d, q0), d1 = __shfl_sync(0xffffffffu, d, q0 + 1);
    const float d2 = __shfl_sync(0xffffffffu, d, q0 + 2), d3 = __shfl_sync(0xffffffffu, d, q0 + 3);
    if (act && sd == 0) {
      const float4 bx = b200::decode_box(r, d0, d1, d2, d3);
      float2* row = reinterpret_cast<float2*>(cand + ((int64_t)b * cap + slot) * 6);
      row[0] = make_float2(bx.x, bx.y);
      row[1] = make_float2(bx.z, bx.w);
    }
  }
}

// Dense-regime form of the box decode (b200yolo_postprocess_dense): only the best kPre = 512 candidates of an image
// (one NMS window: what greedy NMS with max_det = 300 typically consumes; anything beyond is decoded by the NMS
// kernel itself, window by window) are decoded here.  They are picked by the threshold the sort published in the
// workspace header and visited in SLOT order (= runs of consecutive anchors,
// as the class filter compacted them), so every 32-byte sector of the DFL channels is fetched once, by one warp
// group; walking order[] instead would touch one sector per 4 useful bytes.  A CTA takes 256 slots, compacts the
// selected ones in shared memory and decodes them 64 at a time (4 lanes per box).  pass 1 is the fallback for
// images whose NMS ran out of ordered entries: everything is decoded.
__global__ void __launch_bounds__(256, 6) box_decode_selected_kernel(const Levels L, float* __restrict__ cand,
                                                                     const int* __restrict__ cand_anchor,
                                                                     const int* __restrict__ cand_count, int cap, int B,
                                                                     int per_image, const int* __restrict__ hdr,
                                                                     int pass) {
  __shared__ int sel[256];
  __shared__ int wcnt[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, sd = tid & 3;
  const int total = B * per_image;
  for (int f = blockIdx.x; f < total; f += gridDim.x) {
    const int b = f / per_image, blk = f - b * per_image;
    if (pass == 1 && !hdr[B + b]) continue;
    const int n = min(cand_count[b], cap);
    if (blk * 256 >= n) continue;
    // the entries decoded ahead of the NMS: the best hdr[4B + b] of the image (threshold published by the sort)
    const unsigned long long thr48 = pass == 1 ? ~0ull
        : (((unsigned long long)(uint32_t)hdr[6 * B + b] << 32) | (unsigned long long)(uint32_t)hdr[5 * B + b]);
    const int s = blk * 256 + tid;
    bool pick = false;
    if (s < n) {
      const float score = cand[((int64_t)b * cap + s) * 6 + 4];
      pick = (b200::make_key(score, cand_anchor[(int64_t)b * cap + s], s) >> 16) <= thr48;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pick);
    if (lane == 0) wcnt[wid] = __popc(bal);
    __syncthreads();
    int base = 0, m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < wid) base += wcnt[w]; m += wcnt[w]; }
    if (pick) sel[base + __popc(bal & ((1u << lane) - 1u))] = s;
    __syncthreads();
    for (int c0 = 0; c0 < m; c0 += 64) {
      const int c = c0 + (tid >> 2);
      const bool act = c < m;
      const int slot = act ? sel[c] : 0;
      const int a = act ? cand_anchor[(int64_t)b * cap + slot] : 0;
      const b200::AnchorRef ar = b200::anchor_ref(L, b, a);
      const float d = act ? b200::dfl_side(ar.p + (long long)(sd * kReg) * ar.cs, ar.cs) : 0.f;
      const int q0 = lane & ~3;
      const float d0 = __shfl_sync(0xffffffffu, d, q0), d1 = __shfl_sync(0xffffffffu, d, q0 + 1);
      const float d2 = __shfl_sync(0xffffffffu, d, q0 + 2), d3 = __shfl_sync(0xffffffffu, d, q0 + 3);
      if (act && sd == 0) {
        const float4 bx = b200::decode_box(ar, d0, d1, d2, d3);
        float2* row = reinterpret_cast<float2*>(cand + ((int64_t)b * cap + slot) * 6);
        row[0] = make_float2(bx.x, bx.y);
        row[1] = make_float2(bx.z, bx.w);
      }
    }
    __syncthreads();                      // sel[] / wcnt[] are reused by the next (image, block) pair
  }
}

// Try the vectorised path; returns -1000 if the shape is not eligible (caller falls back to the scalar kernel).
template <bool RAW>
static int launch_vec(const Levels& L, int B, int nc, int cls0, float conf, const uint32_t* class_mask, float* cand,
                      int* cand_anchor, int* cand_count, int cap, bool defer_boxes, cudaStream_t stream) {
  const int A = L.off[B200YOLO_MAX_LEVELS];
  FlatSegs S;
  // one flat segment if the levels are views of one concatenated tensor, else one per level
  bool concat = true;
  for (int l = 1; l < L.n; ++l)
    concat = concat && L.ptr[l] == L.ptr[0] + L.off[l] && L.cstride[l] == L.cstride[0] && L.bstride[l] == L.bstride[0];
  S.n = concat ? 1 : L.n;
  int maxcount = 0;
  for (int l = 0; l < B200YOLO_MAX_LEVELS; ++l) {
    const bool used = l < S.n;
    S.ptr[l] = used ? L.ptr[l] : L.ptr[0];
    S.bstride[l] = used ? L.bstride[l] : 0;
    S.cstride[l] = used ? L.cstride[l] : 0;
    S.count[l] = used ? (concat ? A : L.off[l + 1] - L.off[l]) : 0;
    S.first[l] = used ? (concat ? 0 : L.off[l]) : 0;
    if (used) {
      if ((reinterpret_cast<uintptr_t>(S.ptr[l]) & 15) || (S.bstride[l] & 3) || (S.cstride[l] & 3) || (S.count[l] & 3))
        return -1000;
      if (S.count[l] > maxcount) maxcount = S.count[l];
    }
  }
  dim3 grid((unsigned)((maxcount + kVA - 1) / kVA), B, S.n);
  decode_vec_kernel<RAW><<<grid, kVA, 0, stream>>>(S, L, nc, cls0, conf, class_mask, cand, cand_anchor, cand_count, cap);
  if (RAW && !defer_boxes) {
    const int amax = cap < A ? cap : A;               // an image has at most min(cap, A) stored survivors
    const int per_image = (amax + 63) / 64;           // blocks of 64 survivors
    const long long total = (long long)B * per_image;
    const int wave = 6 * B200_NUM_SMS;                // resident CTAs (6 per SM at 40 registers)
    box_decode_kernel<<<(unsigned)(total < wave ? total : wave), 256, 0, stream>>>(L, cand, cand_anchor, cand_count, cap, B,
                                                                                 per_image);
  }
  return b200_launch_status();
}

}  // namespace

int b200_box_decode_sorted_launch(const b200yolo_level* levels, int n_levels, float* cand, const int* cand_anchor,
                                  const int* cand_count, const int* order, int B, int cap, int max_nms, const int* hdr,
                                  int pass, cudaStream_t s) {
  (void)order; (void)max_nms;             // selection is by the published threshold, in slot order
  Levels L;
  const int st = b200::build_levels(levels, n_levels, L);
  if (st != B200YOLO_OK) return st;
  const int A = L.off[B200YOLO_MAX_LEVELS];
  const int amax = cap < A ? cap : A;
  const int per_image = (amax + 255) / 256;
  const long long total = (long long)B * per_image;
  const int wave = 6 * B200_NUM_SMS;
  box_decode_selected_kernel<<<(unsigned)(total < wave ? total : wave), 256, 0, s>>>(L, cand, cand_anchor, cand_count, cap, B,
                                                                                   per_image, hdr, pass);
  return b200_launch_status();
}

static int decode_filter_impl(const b200yolo_level* levels, int n_levels, int B, int nc, float conf_thres,
                              const uint32_t* class_mask, float* cand, int* cand_anchor, int* cand_count, int cap,
                              bool defer_boxes, void* stream) {
  B200_REQUIRE(levels && cand && cand_anchor && cand_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && B <= 65535 && nc > 0 && cap > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(nc <= B200YOLO_MAX_CLASSES, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f, B200YOLO_ERR_RANGE);
  Levels L;
  const int st = b200::build_levels(levels, n_levels, L);
  if (st != B200YOLO_OK) return st;
  const int A = L.off[B200YOLO_MAX_LEVELS];
  {
    const int rc = launch_vec<true>(L, B, nc, 4 * kReg, conf_thres, class_mask, cand, cand_anchor, cand_count, cap,
                                    defer_boxes, (cudaStream_t)stream);
    if (rc != -1000) return rc;
  }
  dim3 grid((unsigned)((A + kAnchorsPerCta - 1) / kAnchorsPerCta), B);
  decode_filter_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(L, nullptr, 0, nc, conf_thres, class_mask,
                                                                           cand, cand_anchor, cand_count, cap,
                                                                           defer_boxes ? 1 : 0);
  return b200_launch_status();
}

extern "C" int b200yolo_decode_filter(const b200yolo_level* levels, int n_levels, int B, int nc, float conf_thres,
                                      const uint32_t* class_mask, float* cand, int* cand_anchor,
                                      int* cand_count, int cap, void* stream) {
  return decode_filter_impl(levels, n_levels, B, nc, conf_thres, class_mask, cand, cand_anchor, cand_count, cap, false,
                            stream);
}

extern "C" int b200yolo_class_filter(const b200yolo_level* levels, int n_levels, int B, int nc, float conf_thres,
                                     const uint32_t* class_mask, float* cand, int* cand_anchor, int* cand_count,
                                     int cap, void* stream) {
  return decode_filter_impl(levels, n_levels, B, nc, conf_thres, class_mask, cand, cand_anchor, cand_count, cap, true,
                            stream);
}

extern "C" int b200yolo_filter_decoded(const float* pred, int B, int channels, int nc, int A, float conf_thres,
                                       const uint32_t* class_mask, float* cand, int* cand_anchor,
                                       int* cand_count, int cap, void* stream) {
  B200_REQUIRE(pred && cand && cand_anchor && cand_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && B <= 65535 && nc > 0 && A > 0 && cap > 0 && channels >= 4 + nc, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(nc <= B200YOLO_MAX_CLASSES && A <= B200YOLO_MAX_ANCHORS, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(conf_thres >= 0.f && conf_thres <= 1.f, B200YOLO_ERR_RANGE);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 3) == 0, B200YOLO_ERR_ALIGN);
  Levels L;
  L.n = 1;
  for (int l = 0; l < B200YOLO_MAX_LEVELS; ++l) {
    L.ptr[l] = pred; L.bstride[l] = (long long)channels * A; L.cstride[l] = A; L.w[l] = 1; L.stride[l] = 1.f;
    L.off[l] = l == 0 ? 0 : A;
  }
  L.off[B200YOLO_MAX_LEVELS] = A;
  {
    const int rc = launch_vec<false>(L, B, nc, 4, conf_thres, class_mask, cand, cand_anchor, cand_count, cap, false,
                                     (cudaStream_t)stream);
    if (rc != -1000) return rc;
  }
  dim3 grid((unsigned)((A + kAnchorsPerCta - 1) / kAnchorsPerCta), B);
  decode_filter_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(L, pred, channels, nc, conf_thres,
                                                                            class_mask, cand, cand_anchor, cand_count,
                                                                            cap, 0);
  return b200_launch_status();
}

// ---- device self-test of the fast math used by the DFL decode ------------------------------------------
// mode 0: expf_torch_m80_0(d) == expf_torch(d) for EVERY float d in [-80, 0] (the 0x42A00001 bit patterns of the
//         negative floats up to 80.0f in magnitude; the kernel walks them all);
// mode 1: div_by_rcp(a, b, rcp_rn(b)) == a / b for a = exp outputs in (1e-26, 1], b in [1, 16], n pairs from a
//         counter-based generator.  mismatches: device int64, incremented per disagreeing input.
namespace {
__global__ void selftest_math_kernel(int mode, unsigned long long n, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (mode == 0) {
      // negative floats: sign bit set, magnitude bits 0 .. 0x42A00000 (= 80.0f)
      const unsigned int bits = 0x80000000u | (unsigned int)i;
      const float d = __uint_as_float(bits);
      const float a = b200::expf_torch_m80_0(d), b = b200::expf_torch(d);
      bad += __float_as_uint(a) != __float_as_uint(b);
    } else {
      // splitmix64 -> two floats
      unsigned long long z = i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
      const float u1 = (float)(unsigned int)(z & 0xffffffu) * (1.0f / 16777216.0f);
      const unsigned int hi = (unsigned int)(z >> 32);
      const float bden = __uint_as_float(0x3f800000u + hi % (0x41800000u - 0x3f800000u + 1u));   // [1, 16]
      const float a = b200::expf_torch_m80_0(-60.0f * u1);                                      // (8.7e-27, 1]
      const float q1 = b200::div_by_rcp(a, bden, __frcp_rn(bden)), q2 = __fdiv_rn(a, bden);
      bad += __float_as_uint(q1) != __float_as_uint(q2);
      // also the raw mantissa sweep: a any normal float in [2^-80, 1]
      const float a2 = __uint_as_float(0x17800000u + (unsigned int)((z >> 8) % (0x3f800000u - 0x17800000u + 1u)));
      const float q3 = b200::div_by_rcp(a2, bden, __frcp_rn(bden)), q4 = __fdiv_rn(a2, bden);
      bad += __float_as_uint(q3) != __float_as_uint(q4);
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int b200yolo_selftest_math(int mode, uint64_t n, uint64_t* mismatches, void* stream) {
  B200_REQUIRE(mismatches, B200YOLO_ERR_NULL);
  B200_REQUIRE(mode == 0 || mode == 1, B200YOLO_ERR_RANGE);
  if (mode == 0) n = 0x42A00000ull + 1ull;
  selftest_math_kernel<<<B200_NUM_SMS * 8, 256, 0, (cudaStream_t)stream>>>(mode, (unsigned long long)n,
                                                                          (unsigned long long*)mismatches);
  return b200_launch_status();
}
