// K3: per-image candidate sort (score descending, anchor ascending on ties) + max_nms cap.
//
// Replaces the stable descending `scores.sort()` inside torchvision.ops.nms and the `n > max_nms`
// cap of ultralytics ops.non_max_suppression (reference entry detect.py:541).  Candidates arrive
// from K2 in arbitrary slot order, so the reference's "stable sort over anchor-ordered rows" is
// reproduced with a composite 64-bit key:
//     key = (~score_bits) << 32 | anchor << 16 | slot          (anchor, slot < 65536)
// Ascending key order == score descending, anchor ascending; the low 16 bits carry the payload
// (slot), so no separate value array moves through the sort.
//
// One CTA per image, keys resident in shared memory:
//   n <= 512 : enumeration sort (rank = number of smaller keys; keys are unique)
//   n >  512 : LSD radix sort, 8-bit digits over key bits 16..63, per-warp histograms; each warp
//              owns a contiguous key segment so the scatter is stable (match_any ranks within a
//              32-key row); passes whose digit is uniform across the image are skipped.
// Latency-bound: reported in microseconds, not GB/s.

#include "nms_common.cuh"

namespace {

using b200::make_key;

constexpr int kEnumMax = 512;
constexpr int kSmemKeysMax = 12288;  // 2 * 8 B * 12288 + 32 KB histograms = 224 KB


template <int NT>
__global__ void __launch_bounds__(NT) sort_topk_kernel(const float* __restrict__ cand,
                                                       const int* __restrict__ cand_anchor,
                                                       const int* __restrict__ cand_count, int cap, int max_nms,
                                                       int* __restrict__ order, uint64_t* __restrict__ ws,
                                                       int smem_keys) {
  constexpr int W = NT / 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* sA = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* sB = sA + smem_keys;
  uint32_t* hist = reinterpret_cast<uint32_t*>(sB + smem_keys);  // [W][256]
  __shared__ uint32_t warp_tot[32];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = min(min(cand_count[b], cap), B200YOLO_MAX_SORT);
  if (n <= 0) return;
  const int n_out = min(n, max_nms);
  const float* crow = cand + (int64_t)b * cap * 6;
  const int* arow = cand_anchor + (int64_t)b * cap;
  int* orow = order + (int64_t)b * cap;

  uint64_t* src = sA;
  uint64_t* dst = sB;
  if (n > smem_keys) {  // image does not fit the shared-memory path: ping-pong in the L2-resident workspace
    src = ws + (int64_t)b * 2 * cap;
    dst = src + cap;
  }
  for (int i = tid; i < n; i += NT) src[i] = make_key(crow[i * 6 + 4], arow[i], i);
  __syncthreads();

  if (n <= kEnumMax) {
    for (int i = tid; i < n; i += NT) {
      const uint64_t k = src[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (src[j] < k);
      if (rank < n_out) orow[rank] = (int)(k & 0xffff);
    }
    return;
  }

  // ---- LSD radix sort over key bits [16, 64) ----
  const int seg = (((n + W - 1) / W) + 31) & ~31;  // keys per warp, multiple of 32
  const int w_beg = min(n, wid * seg), w_end = min(n, (wid + 1) * seg);
  for (int shift = 16; shift < 64; shift += 8) {
    // phase 0: uniform-digit test + histogram
    for (int i = tid; i < W * 256; i += NT) hist[i] = 0;
    const uint32_t d0 = (uint32_t)(src[0] >> shift) & 0xff;
    __syncthreads();
    int same = 1;
    for (int i = w_beg + lane; i < w_end; i += 32) {
      const uint32_t d = (uint32_t)(src[i] >> shift) & 0xff;
      same &= (d == d0);
      atomicAdd(&hist[wid * 256 + d], 1u);
    }
    if (__syncthreads_and(same)) continue;  // every key has the same digit: pass is the identity

    // phase 1: exclusive scan over (digit major, warp minor)
    {
      constexpr int PER = (W * 256) / NT;  // entries per thread (8)
      const int e0 = tid * PER;
      uint32_t v[PER], sum = 0;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int e = e0 + k, d = e / W, w = e % W;
        v[k] = hist[w * 256 + d];
        sum += v[k];
      }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) warp_tot[wid] = inc;
      __syncthreads();
      if (wid == 0) {
        uint32_t t = lane < W ? warp_tot[lane] : 0, ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
          if (lane >= o) ti += u;
        }
        warp_tot[lane] = ti - t;  // exclusive
      }
      __syncthreads();
      uint32_t run = warp_tot[wid] + inc - sum;
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        const int e = e0 + k, d = e / W, w = e % W;
        hist[w * 256 + d] = run;
        run += v[k];
      }
    }
    __syncthreads();

    // phase 2: stable scatter; each warp walks its own segment in order
    for (int base = w_beg; base < w_end; base += 32) {
      const int i = base + lane;
      const bool act = i < w_end;
      const unsigned amask = __ballot_sync(0xffffffffu, act);
      if (act) {
        const uint64_t k = src[i];
        const uint32_t d = (uint32_t)(k >> shift) & 0xff;
        const unsigned peers = __match_any_sync(amask, d);
        const int leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if (lane == leader) {
          pos = hist[wid * 256 + d];
          hist[wid * 256 + d] = pos + __popc(peers);
        }
        pos = __shfl_sync(peers, pos, leader);
        dst[pos + __popc(peers & ((1u << lane) - 1u))] = k;
      }
      __syncwarp();
    }
    __syncthreads();
    uint64_t* t = src; src = dst; dst = t;
  }
  for (int r = tid; r < n_out; r += NT) orow[r] = (int)(src[r] & 0xffff);
}

}  // namespace

extern "C" size_t b200yolo_workspace_bytes(int B, int cap) {
  if (B <= 0 || cap <= 0) return 0;
  if (cap <= kSmemKeysMax) return 16;  // shared-memory paths only; keep a non-null minimum
  // sort: 2 key buffers of cap u64; nms: cap float4 boxes + cap flag bytes -- the larger of the two
  return (size_t)B * ((size_t)cap + (cap + 15) / 16) * 16;
}

extern "C" int b200yolo_sort_topk(const float* cand, const int* cand_anchor, const int* cand_count, int B,
                                  int cap, int max_nms, int* order, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  B200_REQUIRE(cand && cand_anchor && cand_count && order, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && cap > 0 && max_nms > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(cap <= B200YOLO_MAX_SORT, B200YOLO_ERR_UNSUPPORTED);
  if (cap > kSmemKeysMax) {
    B200_REQUIRE(workspace, B200YOLO_ERR_NULL);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, B200YOLO_ERR_ALIGN);
    B200_REQUIRE(workspace_bytes >= b200yolo_workspace_bytes(B, cap), B200YOLO_ERR_WORKSPACE);
  }
  const int smem_keys = cap < kSmemKeysMax ? cap : kSmemKeysMax;
  cudaStream_t s = (cudaStream_t)stream;
  if (cap <= 1024) {
    constexpr int NT = 256;
    const size_t smem = 2 * sizeof(uint64_t) * (size_t)smem_keys + (NT / 32) * 256 * sizeof(uint32_t);
    sort_topk_kernel<NT><<<B, NT, smem, s>>>(cand, cand_anchor, cand_count, cap, max_nms, order,
                                             (uint64_t*)workspace, smem_keys);
  } else {
    constexpr int NT = 1024;
    const size_t smem = 2 * sizeof(uint64_t) * (size_t)smem_keys + (NT / 32) * 256 * sizeof(uint32_t);
    auto kern = sort_topk_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<B, NT, smem, s>>>(cand, cand_anchor, cand_count, cap, max_nms, order, (uint64_t*)workspace, smem_keys);
  }
  return b200_launch_status();
}
