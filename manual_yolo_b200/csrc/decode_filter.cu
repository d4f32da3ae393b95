// K2: YOLOv8 Detect-head decode + confidence filter + compaction.
//
// Replaces ultralytics Detect._inference (DFL softmax expectation over 16 bins, dist2bbox(xywh) *
// stride, class sigmoid; nn/modules/head.py, block.py, utils/tal.py) fused with the head of
// ops.non_max_suppression (amax > conf, xywh2xyxy, cls.max, conf filter, classes filter) -- reference
// entry detect.py:541 / yolo.py:361 / pipe.py:179.  Restated in oracle/head.py + oracle/nms.py.
//
// A CTA of 8 warps covers 64 consecutive anchors: two groups of 4 warps, each group owning 32 anchors
// (lane = anchor, so every channel load is one coalesced 128-byte line).  Within a group the channel
// axis is split four ways: warp q reduces class channels [q*nc/4, (q+1)*nc/4) -- all of a thread's loads
// are independent and issued in ONE round (16 in flight for nc=64) -- and the partial (max, argmax)
// pairs are merged through shared memory in class order (first-maximum semantics of cls.max(1)).
// Work is lazy: only if some anchor of the CTA beats conf_thres (about 1% of anchors at conf 0.25, all of
// them at conf 0.001) do the four warps of a group decode the four box sides (warp q = side q: 16 DFL
// bins, softmax expectation), again one load round.  Survivors are compacted with a warp ballot and one
// atomicAdd per group on the per-image counter.
// HBM-bound: algorithmic bytes per frame = (64+nc)*A*4 read + 28 B per survivor written.
// Compiled with -fmad=false: every add/mul below rounds separately, as the torch CPU ops do.

#include <cuda.h>

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
                                                                 int* __restrict__ cand_count, int cap) {
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
// TMA-staged variant (the fast path; the kernel above is the fallback for unaligned shapes).
//
// The head is channel-major with the anchor axis contiguous, so a [128 anchors x 16 channels] box is a
// 2-D tile of the (A, channels, B) tensor: one cp.async.bulk.tensor instruction (SASS UTMALDG) per 8 KB
// chunk, completion on an mbarrier, zero fill past A.  A persistent CTA (256 threads) walks tiles of 128
// anchors; the class chunks of tile t+1 are in flight in the second buffer set while tile t is reduced
// (thread = anchor x channel-half, conflict-free LDS columns).  Survivors are rare at conf 0.25, so the 64
// DFL channels are read lazily: a handful of candidates fetch them straight from global memory; when a
// tile is dense (eval regime, conf 0.001) its [128 x 64] box tile is pulled through TMA as well and
// decoded from shared memory by all 256 threads (thread = anchor x side-pair).
// =================================================================================================
constexpr int kTA = 128;          // anchors per tile
constexpr int kCC = 16;           // class channels per TMA chunk (8 KB)
constexpr int kDenseMin = 16;     // candidates in a tile from which the box tile is TMA-staged

struct Segs {
  int n;                          // tensor maps in use (1 = concatenated head, else one per level)
  int tiles[B200YOLO_MAX_LEVELS]; // tiles per image in each segment
  int first[B200YOLO_MAX_LEVELS]; // first global anchor index of the segment
  int count[B200YOLO_MAX_LEVELS]; // anchors in the segment
  int tiles_per_image;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(b200::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2),
      "r"(b200::smem_u32(bar))
      : "memory");
}

template <bool RAW>
__global__ void __launch_bounds__(256, 2) decode_tma_kernel(const __grid_constant__ CUtensorMap map0,
                                                            const __grid_constant__ CUtensorMap map1,
                                                            const __grid_constant__ CUtensorMap map2, const Segs S,
                                                            const Levels L, int B, int nc, int cls0, float conf,
                                                            const uint32_t* __restrict__ class_mask,
                                                            float* __restrict__ cand, int* __restrict__ cand_anchor,
                                                            int* __restrict__ cand_count, int cap) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  __shared__ __align__(8) uint64_t full[2], boxbar;
  __shared__ Part part[kTA];
  __shared__ float dist[4][kTA];
  __shared__ unsigned char flag[kTA];
  const int nchunk = (nc + kCC - 1) / kCC;
  const int set_bytes = nchunk * kTA * kCC * 4;
  float* cls_buf[2] = {reinterpret_cast<float*>(dsm), reinterpret_cast<float*>(dsm + set_bytes)};
  float* box_buf = reinterpret_cast<float*>(dsm + 2 * set_bytes);     // [64][kTA] (RAW only)
  const int tid = threadIdx.x, lane = tid & 31;
  const int al = tid & (kTA - 1), half = tid >> 7;                    // anchor-in-tile, channel half / side pair
  const int total_tiles = S.tiles_per_image * B;

  if (tid == 0) {
    b200::mbar_init(&full[0], 1);
    b200::mbar_init(&full[1], 1);
    b200::mbar_init(&boxbar, 1);
    b200::mbar_fence_init();
  }
  __syncthreads();

  auto locate = [&](int t, int& b, int& seg, int& a0) {     // tile id -> image, segment, first local anchor
    b = t / S.tiles_per_image;
    int r = t - b * S.tiles_per_image;
    seg = 0;
    if (S.n > 1 && r >= S.tiles[0]) { r -= S.tiles[0]; seg = 1; if (S.n > 2 && r >= S.tiles[1]) { r -= S.tiles[1]; seg = 2; } }
    a0 = r * kTA;
  };
  auto map_of = [&](int seg) -> const CUtensorMap* { return seg == 0 ? &map0 : (seg == 1 ? &map1 : &map2); };
  auto issue = [&](int t, int set) {                         // thread 0: all class chunks of tile t -> buffer set
    int b, seg, a0;
    locate(t, b, seg, a0);
    b200::mbar_expect_tx(&full[set], (uint32_t)set_bytes);
    for (int k = 0; k < nchunk; ++k)
      tma_load_3d(cls_buf[set] + k * kTA * kCC, map_of(seg), a0, cls0 + k * kCC, b, &full[set]);
  };

  int t = blockIdx.x;
  if (tid == 0) {
    if (t < total_tiles) issue(t, 0);
    if (t + (int)gridDim.x < total_tiles) issue(t + gridDim.x, 1);
  }
  int it = 0, box_phase = 0;
  for (; t < total_tiles; t += gridDim.x, ++it) {
    const int set = it & 1;
    int b, seg, a0;
    locate(t, b, seg, a0);
    const int seg_first = seg == 0 ? S.first[0] : (seg == 1 ? S.first[1] : S.first[2]);
    const int seg_count = seg == 0 ? S.count[0] : (seg == 1 ? S.count[1] : S.count[2]);
    const int a_loc = a0 + al;                       // anchor index inside the segment
    const int a = seg_first + a_loc;                 // global anchor index
    const bool live = a_loc < seg_count;

    b200::mbar_wait(&full[set], (it >> 1) & 1);
    // ---- class reduction: thread = (anchor, channel half); columns of the staged chunks ----
    float m = -INFINITY, m2 = -INFINITY;
    int j = 0;
    {
      const float* col = cls_buf[set] + al;
      const int cbeg = half * ((nchunk * kCC) >> 1), cend = cbeg + ((nchunk * kCC) >> 1);
      j = cbeg;
#pragma unroll 4
      for (int c = cbeg; c < cend; c += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = col[(c + u) * kTA];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (c + u < nc && v[u] > m) { m2 = m; m = v[u]; j = c + u; }
      }
    }
    if (half == 1) part[al] = Part{m, m2, j};
    __syncthreads();                                  // (1) every thread is done reading cls_buf[set]
    if (tid == 0 && t + 2 * (int)gridDim.x < total_tiles) issue(t + 2 * gridDim.x, set);
    float score = 0.f;
    int cls = 0;
    bool is_cand = false;
    if (half == 0) {
      const Part o = part[al];
      if (o.m > m) { m2 = fmaxf(m, o.m2); m = o.m; j = o.j; }
      score = RAW ? b200::sigmoid_torch(m) : m;
      is_cand = live && (score > conf);
      cls = j;
    }
    // per-lane level parameters (RAW decode) and element pointer of this anchor's channel 0
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
    if (half == 0) {
      if (RAW && is_cand && m2 > -INFINITY && b200::sigmoid_torch(m2) == score) {
        const float* pc = p + (long long)cls0 * cs;   // lowest class index whose sigmoid ties the maximum
        for (int c = 0; c < j; ++c)
          if (b200::sigmoid_torch(pc[(long long)c * cs]) == score) { cls = c; break; }
      }
      if (is_cand && !class_allowed(class_mask, cls)) is_cand = false;
      flag[al] = is_cand;
    }
    const int ncand = __syncthreads_count(is_cand);   // (2) also publishes flag[]
    if (ncand == 0) continue;

    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    if (RAW) {
      const bool dense = ncand >= kDenseMin;
      if (dense) {
        // box tile [128 anchors x 64 DFL channels] through TMA, decoded from shared memory
        if (tid == 0) {
          b200::mbar_expect_tx(&boxbar, kTA * 4 * kReg * 4);
          for (int k = 0; k < 4 * kReg / kCC; ++k)
            tma_load_3d(box_buf + k * kTA * kCC, map_of(seg), a0, k * kCC, b, &boxbar);
        }
        b200::mbar_wait(&boxbar, box_phase & 1);
        ++box_phase;
      }
      if (flag[al]) {
        for (int sd = half * 2; sd < half * 2 + 2; ++sd) {       // this thread's two box sides
          float v[kReg];
          if (dense) {
#pragma unroll
            for (int k = 0; k < kReg; ++k) v[k] = box_buf[(sd * kReg + k) * kTA + al];
          } else {
#pragma unroll
            for (int k = 0; k < kReg; ++k) v[k] = p[(long long)(sd * kReg + k) * cs];
          }
          float mx = v[0];
#pragma unroll
          for (int k = 1; k < kReg; ++k) mx = fmaxf(mx, v[k]);
          float sum = 0.f;
#pragma unroll
          for (int k = 0; k < kReg; ++k) { v[k] = b200::expf_torch(__fsub_rn(v[k], mx)); sum = __fadd_rn(sum, v[k]); }
          float acc = 0.f;   // torch's 1x1 conv with arange weights: sequential fma over the 16 bins
#pragma unroll
          for (int k = 0; k < kReg; ++k) acc = __fmaf_rn((float)k, __fdiv_rn(v[k], sum), acc);
          dist[sd][al] = acc;
        }
      }
      __syncthreads();                                // (3) dist[] complete; box_buf free again
      if (half == 0 && is_cand) {
        const float d0 = dist[0][al], d1 = dist[1][al], d2 = dist[2][al], d3 = dist[3][al];
        const float ax = (float)(i % lw) + 0.5f, ay = (float)(i / lw) + 0.5f;
        const float bx1 = ax - d0, by1 = ay - d1, bx2 = ax + d2, by2 = ay + d3;
        const float cx = ((bx1 + bx2) / 2.0f) * st, cy = ((by1 + by2) / 2.0f) * st;
        const float bw = (bx2 - bx1) * st, bh = (by2 - by1) * st;
        const float hw = bw / 2.0f, hh = bh / 2.0f;
        x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
      }
    } else if (half == 0 && is_cand) {
      const float cx = p[0], cy = p[cs], bw = p[2 * cs], bh = p[3 * cs];
      const float hw = bw / 2.0f, hh = bh / 2.0f;
      x1 = cx - hw; y1 = cy - hh; x2 = cx + hw; y2 = cy + hh;
    }
    if (half == 0) emit(is_cand, b, a, x1, y1, x2, y2, score, cls, cand, cand_anchor, cand_count, cap);
    (void)lane;
  }
}

// ---- host: tensor-map construction through the driver entry point (no libcuda link dependency) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;          // resolved once; read-only afterwards
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// (anchors, channels, B) fp32 tensor -> map with a [kTA x kCC x 1] box.  Returns false if TMA cannot express it.
static bool make_map(CUtensorMap* map, const float* ptr, long long anchors, long long channels, long long B,
                     long long chan_stride, long long batch_stride) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (chan_stride * 4) % 16 || (batch_stride * 4) % 16) return false;
  cuuint64_t dims[3] = {(cuuint64_t)anchors, (cuuint64_t)channels, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)chan_stride * 4, (cuuint64_t)batch_stride * 4};
  cuuint32_t box[3] = {kTA, kCC, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Try the TMA path; returns -1000 if the shape is not eligible (caller falls back to the plain kernel).
template <bool RAW>
static int launch_tma(const Levels& L, int B, int channels, int nc, int cls0, float conf, const uint32_t* class_mask,
                      float* cand, int* cand_anchor, int* cand_count, int cap, cudaStream_t stream) {
  const int A = L.off[B200YOLO_MAX_LEVELS];
  const int nchunk = (nc + kCC - 1) / kCC;
  if ((nchunk * kCC) % 8 != 0) return -1000;
  const size_t smem = 2 * (size_t)nchunk * kTA * kCC * 4 + (RAW ? (size_t)kTA * 4 * kReg * 4 : 0) + 1024;
  if (smem > 200 * 1024) return -1000;
  CUtensorMap maps[3];
  Segs S;
  // one map if the levels are views of one concatenated tensor, else one per level
  bool concat = true;
  for (int l = 1; l < L.n; ++l)
    concat = concat && L.ptr[l] == L.ptr[0] + L.off[l] && L.cstride[l] == L.cstride[0] && L.bstride[l] == L.bstride[0];
  if (concat) {
    if (!make_map(&maps[0], L.ptr[0], A, channels, B, L.cstride[0], L.bstride[0])) return -1000;
    maps[1] = maps[0]; maps[2] = maps[0];
    S.n = 1; S.tiles[0] = (A + kTA - 1) / kTA; S.first[0] = 0; S.count[0] = A;
    S.tiles[1] = S.tiles[2] = 0; S.first[1] = S.first[2] = 0; S.count[1] = S.count[2] = 0;
    S.tiles_per_image = S.tiles[0];
  } else {
    S.n = L.n; S.tiles_per_image = 0;
    for (int l = 0; l < B200YOLO_MAX_LEVELS; ++l) {
      if (l < L.n) {
        const int cnt = L.off[l + 1] - L.off[l];
        if (!make_map(&maps[l], L.ptr[l], cnt, channels, B, L.cstride[l], L.bstride[l])) return -1000;
        S.tiles[l] = (cnt + kTA - 1) / kTA; S.first[l] = L.off[l]; S.count[l] = cnt;
        S.tiles_per_image += S.tiles[l];
      } else {
        maps[l] = maps[0]; S.tiles[l] = 0; S.first[l] = 0; S.count[l] = 0;
      }
    }
  }
  auto kern = decode_tma_kernel<RAW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long total = (long long)S.tiles_per_image * B;
  const int ctas_per_sm = smem <= 110 * 1024 ? 2 : 1;
  const int grid = (int)(total < (long long)B200_NUM_SMS * ctas_per_sm ? total : B200_NUM_SMS * ctas_per_sm);
  kern<<<grid, 256, smem, stream>>>(maps[0], maps[1], maps[2], S, L, B, nc, cls0, conf, class_mask, cand, cand_anchor,
                                    cand_count, cap);
  return b200_launch_status();
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
  {
    const int rc = launch_tma<true>(L, B, 4 * kReg + nc, nc, 4 * kReg, conf_thres, class_mask, cand, cand_anchor,
                                    cand_count, cap, (cudaStream_t)stream);
    if (rc != -1000) return rc;
  }
  dim3 grid((unsigned)((off + kAnchorsPerCta - 1) / kAnchorsPerCta), B);
  decode_filter_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(L, nullptr, 0, nc, conf_thres, class_mask,
                                                                           cand, cand_anchor, cand_count, cap);
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
  Levels L;
  L.n = 1;
  for (int l = 0; l < B200YOLO_MAX_LEVELS; ++l) {
    L.ptr[l] = pred; L.bstride[l] = (long long)channels * A; L.cstride[l] = A; L.w[l] = 1; L.stride[l] = 1.f;
    L.off[l] = l == 0 ? 0 : A;
  }
  L.off[B200YOLO_MAX_LEVELS] = A;
  {
    const int rc = launch_tma<false>(L, B, channels, nc, 4, conf_thres, class_mask, cand, cand_anchor, cand_count, cap,
                                     (cudaStream_t)stream);
    if (rc != -1000) return rc;
  }
  dim3 grid((unsigned)((A + kAnchorsPerCta - 1) / kAnchorsPerCta), B);
  decode_filter_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(L, pred, channels, nc, conf_thres,
                                                                            class_mask, cand, cand_anchor, cand_count,
                                                                            cap);
  return b200_launch_status();
}
