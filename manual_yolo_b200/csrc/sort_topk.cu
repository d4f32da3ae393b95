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
//   cap > 2048 (dense regime): sort_select_kernel -- greedy NMS consumes the sorted list from the top and stops
//              at max_det keeps, so only the best kSelK = 2048 keys are put in order: a 6-pass MSD radix SELECT
//              finds the kSelK-th smallest key (early exit when the selected bin is taken whole), the keys up to it
//              are compacted and bitonic-sorted (two keys per thread in registers: the steps at distance <= 32 are
//              warp shuffles, only the 15 steps at distance >= 64 go through shared memory).  The number
//              of sorted entries goes to the workspace header; if the NMS ever runs out of them before max_det
//              keeps, it raises a per-image flag and b200yolo_nms re-runs that image with the full sort above
//              (exact in every case, one cheap path in the common one).
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
                                                       int smem_keys, int* __restrict__ hdr, int B, int pass) {
  constexpr int W = NT / 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* sA = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* sB = sA + smem_keys;
  uint32_t* hist = reinterpret_cast<uint32_t*>(sB + smem_keys);  // [W][256]
  __shared__ uint32_t warp_tot[32];

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (pass == 1 && !(hdr && hdr[B + b])) return;           // fallback launch: only the images the NMS flagged
  const int n = min(min(cand_count[b], cap), B200YOLO_MAX_SORT);
  const int n_out = min(max(n, 0), max_nms);
  if (hdr && tid == 0) {   // every entry sorted, no fallback pending, thresholds = "every key"
    hdr[b] = n_out; hdr[2 * B + b] = -1; hdr[3 * B + b] = -1;
    // boxes decoded ahead of the NMS: pass 0 -> all of them (the decode kernel takes every key); pass 1 (fallback
    // re-sort): none is assumed -- the NMS decodes every window it loads
    hdr[4 * B + b] = pass == 0 ? n_out : 0; hdr[5 * B + b] = -1; hdr[6 * B + b] = -1;
    if (pass == 0) hdr[B + b] = 0;
  }
  if (n <= 0) return;
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

// ---- dense regime: select the best kSelK keys, sort only those ----------------------------------------
constexpr int kSelK = 2048;
constexpr int kPre = 512;          // entries whose boxes b200yolo_postprocess_dense decodes ahead of the NMS (one NMS window)

// Bitonic network over the kSelK = 2048 keys in s[], 1024 threads.  Thread t keeps elements 2t and 2t+1 in registers:
// a compare-exchange at distance j <= 32 pairs element i with i ^ j, i.e. the same register of lane ^ (j / 2) (or the
// thread's other register for j = 1) -- 51 of the network's 66 steps are warp shuffles without a CTA barrier; only the
// 15 steps at distance >= 64 go through shared memory.
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void bitonic_reg_steps(uint64_t& e0, uint64_t& e1, int i0, int k, int jstart) {
  const bool up = (i0 & k) == 0;
  for (int j = jstart; j >= 2; j >>= 1) {
    const uint64_t p0 = shfl_xor_u64(e0, j >> 1), p1 = shfl_xor_u64(e1, j >> 1);
    const bool keep_min = ((i0 & j) == 0) == up;
    e0 = (keep_min == (p0 < e0)) ? p0 : e0;
    e1 = (keep_min == (p1 < e1)) ? p1 : e1;
  }
  if ((e0 > e1) == up) { const uint64_t t = e0; e0 = e1; e1 = t; }
}
__device__ __forceinline__ void bitonic_sort_2048(uint64_t* s, int tid, int nt) {
  (void)nt;                                                   // launched with kSelK / 2 threads
  const int i0 = 2 * tid;
  uint64_t e0 = s[i0], e1 = s[i0 + 1];
  for (int k = 2; k <= 64; k <<= 1) bitonic_reg_steps(e0, e1, i0, k, k >> 1);
  for (int k = 128; k <= kSelK; k <<= 1) {
    s[i0] = e0; s[i0 + 1] = e1;
    __syncthreads();
    for (int j = k >> 1; j >= 64; j >>= 1) {
      const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1)), l = i | j;
      const uint64_t a = s[i], c = s[l];
      if ((a > c) == ((i & k) == 0)) { s[i] = c; s[l] = a; }
      __syncthreads();
    }
    e0 = s[i0]; e1 = s[i0 + 1];
    bitonic_reg_steps(e0, e1, i0, k, 32);
  }
  s[i0] = e0; s[i0 + 1] = e1;
  __syncthreads();
}

__global__ void __launch_bounds__(1024, 2) sort_select_kernel(const float* __restrict__ cand,
                                                              const int* __restrict__ cand_anchor,
                                                              const int* __restrict__ cand_count, int cap, int max_nms,
                                                              int* __restrict__ order, uint64_t* __restrict__ ws,
                                                              int keys_in_smem, int* __restrict__ hdr, int B) {
  constexpr int NT = 1024;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* sB = reinterpret_cast<uint64_t*>(smem_raw);          // [kSelK] the selected keys
  uint64_t* src = sB + kSelK;                                     // [cap] all keys (shared memory, or the workspace)
  __shared__ uint32_t hist[256];
  __shared__ unsigned long long sel_prefix;
  __shared__ int sel_k, cursor, sel_done;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = min(min(cand_count[b], cap), B200YOLO_MAX_SORT);
  const int n_out = min(max(n, 0), max_nms);
  if (tid == 0) {
    hdr[b] = min(n_out, kSelK); hdr[B + b] = 0; hdr[2 * B + b] = -1; hdr[3 * B + b] = -1;
    hdr[4 * B + b] = min(n_out, kPre); hdr[5 * B + b] = -1; hdr[6 * B + b] = -1;     // (all ones: every key) for n <= kPre
  }
  if (n <= 0) return;
  if (!keys_in_smem) src = ws + (int64_t)b * 2 * cap;
  const float* crow = cand + (int64_t)b * cap * 6;
  const int* arow = cand_anchor + (int64_t)b * cap;
  int* orow = order + (int64_t)b * cap;

  if (n <= kEnumMax) {
    // sparse image inside a dense-capable launch: enumeration sort (rank = number of smaller keys; keys are unique)
    for (int i = tid; i < n; i += NT) sB[i] = make_key(crow[i * 6 + 4], arow[i], i);
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
      const uint64_t k = sB[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (sB[j] < k);
      if (rank < n_out) orow[rank] = (int)(k & 0xffff);
    }
    return;
  }
  if (n <= kSelK) {
    // everything fits the bitonic network: pad with the largest key
    for (int i = tid; i < kSelK; i += NT) sB[i] = i < n ? make_key(crow[i * 6 + 4], arow[i], i) : ~0ull;
    __syncthreads();
  } else {
    for (int i = tid; i < n; i += NT) src[i] = make_key(crow[i * 6 + 4], arow[i], i);
    if (tid == 0) cursor = 0;
    // MSD radix select over key bits [16, 64): prefix = top bits of the kSelK-th smallest key so far
    unsigned long long prefix = 0;
    int kk = kSelK;
    for (int shift = 56; shift >= 16; shift -= 8) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      // (scores cluster: most keys of a warp fall into a few bins -- one shared-memory atomic per distinct bin and warp)
      // (atomicAdd(.., 1) on shared memory compiles to ATOMS.POPC.INC, which already merges the lanes of a warp that
      // hit the same bin: explicit warp aggregation -- match_any, or a uniform-warp test -- measured 10-18 us slower)
      for (int i = tid; i < n; i += NT) {
        const uint64_t k = src[i];
        if (shift == 56 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(uint32_t)(k >> shift) & 0xff], 1u);
      }
      __syncthreads();
      if (wid == 0) {                       // 8 bins per lane: the bin where the cumulative count reaches kk
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; sum += c[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const uint32_t excl = inc - sum;
        if (excl < (uint32_t)kk && (uint32_t)kk <= inc) {
          uint32_t run = excl;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (run < (uint32_t)kk && (uint32_t)kk <= run + c[j]) {
              sel_prefix = (prefix << 8) | (unsigned long long)(lane * 8 + j);
              sel_k = kk - (int)run;
              sel_done = (uint32_t)kk == run + c[j];    // the whole bin is taken: the lower digits decide nothing
            }
            run += c[j];
          }
        }
      }
      __syncthreads();
      prefix = sel_prefix;
      kk = sel_k;
      if (sel_done) {                                   // (with distinct scores: after the 32 score bits, 4 passes of 6)
        prefix = (prefix << (shift - 16)) | ((1ull << (shift - 16)) - 1ull);
        break;
      }
    }
    // keys are unique in their top 48 bits (score, anchor): exactly kSelK keys satisfy (key >> 16) <= prefix.
    // The threshold goes to the header: the box decode of b200yolo_postprocess_dense selects by it, in slot order.
    if (tid == 0) { hdr[2 * B + b] = (int)(uint32_t)(prefix & 0xffffffffu); hdr[3 * B + b] = (int)(uint32_t)(prefix >> 32); }
    for (int base = 0; base < n; base += NT) {
      const int i = base + tid;
      const uint64_t k = i < n ? src[i] : 0ull;
      const bool in = i < n && (k >> 16) <= prefix;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      int pos = 0;
      if (lane == 0 && bal) pos = atomicAdd(&cursor, __popc(bal));
      pos = __shfl_sync(0xffffffffu, pos, 0);
      B200_CHECK(!in || pos + __popc(bal & ((1u << lane) - 1u)) < kSelK);     // exactly kSelK keys pass the threshold
      if (in) sB[pos + __popc(bal & ((1u << lane) - 1u))] = k;
    }
    __syncthreads();
  }
  bitonic_sort_2048(sB, tid, NT);
  const int m = min(n_out, kSelK);
  for (int r = tid; r < m; r += NT) {
    B200_CHECK((int)(sB[r] & 0xffff) < n && (r == 0 || sB[r - 1] < sB[r]));      // a slot of this image, keys ascending
    orow[r] = (int)(sB[r] & 0xffff);
  }
  if (tid == 0 && n > kPre) {
    // threshold of the kPre best entries: the box decode ahead of the NMS takes exactly those (in slot order)
    const unsigned long long t = sB[min(m, kPre) - 1] >> 16;
    hdr[5 * B + b] = (int)(uint32_t)(t & 0xffffffffu); hdr[6 * B + b] = (int)(uint32_t)(t >> 32);
  }
}

}  // namespace

// Workspace layout: [header: 7 x B ints -- entries of order[] that are sorted | image needs the full-sort fallback |
// low 32 / high 16 bits of the selection threshold (key >> 16 of the last ordered entry; all ones = every key) |
// entries whose boxes are decoded ahead of the NMS | low / high bits of their threshold -- padded to 16 B]
// [payload: per image 2 * cap u64 keys, used when cap exceeds the shared-memory paths].
static size_t ws_header_bytes(int B) { return (((size_t)B * 7 * sizeof(int)) + 15) & ~(size_t)15; }

extern "C" size_t b200yolo_workspace_bytes(int B, int cap) {
  if (B <= 0 || cap <= 0) return 0;
  const size_t payload = cap <= kSmemKeysMax ? 16 : (size_t)B * ((size_t)cap + (cap + 15) / 16) * 16;
  return ws_header_bytes(B) + payload;
}

// pass 0: the regular sort; pass 1: full sort of the images whose fallback flag is set (launched by b200yolo_nms)
int b200_sort_launch(const float* cand, const int* cand_anchor, const int* cand_count, int B, int cap, int max_nms,
                     int* order, void* workspace, size_t workspace_bytes, int pass, cudaStream_t s) {
  const bool have_hdr = workspace != nullptr && workspace_bytes >= b200yolo_workspace_bytes(B, cap);
  int* hdr = have_hdr ? reinterpret_cast<int*>(workspace) : nullptr;
  uint64_t* payload = have_hdr ? reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(workspace) + ws_header_bytes(B))
                               : reinterpret_cast<uint64_t*>(workspace);
  if (cap > kSmemKeysMax) B200_REQUIRE(have_hdr, B200YOLO_ERR_WORKSPACE);
  if (pass == 1 && !have_hdr) return B200YOLO_OK;
  if (pass == 0 && have_hdr && cap > kSelK) {
    const int in_smem = cap <= kSmemKeysMax ? 1 : 0;
    const size_t smem = sizeof(uint64_t) * ((size_t)kSelK + (in_smem ? (size_t)cap : 0));
    cudaError_t e = cudaFuncSetAttribute(sort_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    sort_select_kernel<<<B, 1024, smem, s>>>(cand, cand_anchor, cand_count, cap, max_nms, order, payload, in_smem, hdr, B);
    return b200_launch_status();
  }
  const int smem_keys = cap < kSmemKeysMax ? cap : kSmemKeysMax;
  if (cap <= 1024) {
    constexpr int NT = 256;
    const size_t smem = 2 * sizeof(uint64_t) * (size_t)smem_keys + (NT / 32) * 256 * sizeof(uint32_t);
    sort_topk_kernel<NT><<<B, NT, smem, s>>>(cand, cand_anchor, cand_count, cap, max_nms, order, payload, smem_keys, hdr, B,
                                             pass);
  } else {
    constexpr int NT = 1024;
    const size_t smem = 2 * sizeof(uint64_t) * (size_t)smem_keys + (NT / 32) * 256 * sizeof(uint32_t);
    auto kern = sort_topk_kernel<NT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<B, NT, smem, s>>>(cand, cand_anchor, cand_count, cap, max_nms, order, payload, smem_keys, hdr, B, pass);
  }
  return b200_launch_status();
}

extern "C" int b200yolo_sort_topk(const float* cand, const int* cand_anchor, const int* cand_count, int B,
                                  int cap, int max_nms, int* order, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  B200_REQUIRE(cand && cand_anchor && cand_count && order, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && cap > 0 && max_nms > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(cap <= B200YOLO_MAX_SORT, B200YOLO_ERR_UNSUPPORTED);
  if (workspace) B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, B200YOLO_ERR_ALIGN);
  if (cap > kSmemKeysMax) B200_REQUIRE(workspace, B200YOLO_ERR_NULL);
  return b200_sort_launch(cand, cand_anchor, cand_count, B, cap, max_nms, order, workspace, workspace_bytes, 0,
                          (cudaStream_t)stream);
}
