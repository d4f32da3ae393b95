// N3: merge step of the SAHI-style sliced prediction the reference runs through
// sahi.predict.get_sliced_prediction (pipe.py:183-194: 640x640 slices, overlap ratio 0.2).  Every slice of a frame
// went through the same path as a frame (K1 slice mode -> K2 -> K3 -> K4, boxes in SLICE pixels); this kernel
// concatenates the kept detections of a frame's slices (slice-major, rank order kept), adds each slice's origin
// (SAHI's shift_amount) and emits the candidate arrays of one more class-aware NMS over the whole frame
// (b200yolo_sort_topk + b200yolo_nms), whose overlap regions would otherwise hold every object twice.
#include "common.cuh"

namespace {

constexpr int kMaxSlices = 64;
struct SliceOrigins { int n; int x[kMaxSlices], y[kMaxSlices]; };

__global__ void __launch_bounds__(256) gather_slices_kernel(const float* __restrict__ det,
                                                            const int* __restrict__ det_count, int max_det,
                                                            const SliceOrigins so, float* __restrict__ cand,
                                                            int* __restrict__ cand_anchor, int* __restrict__ cand_count,
                                                            int cap, const float* __restrict__ full_det,
                                                            const int* __restrict__ full_count) {
  __shared__ int offs[kMaxSlices + 1];
  const int f = blockIdx.x, tid = threadIdx.x, ns = so.n;
  if (tid < 32) {
    // exclusive prefix over the slice counts: two slices per lane
    const int s0 = 2 * tid, s1 = 2 * tid + 1;
    const int c0 = s0 < ns ? min(det_count[f * ns + s0], max_det) : 0;
    const int c1 = s1 < ns ? min(det_count[f * ns + s1], max_det) : 0;
    int inc = c0 + c1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (tid >= o) inc += t;
    }
    const int excl = inc - c0 - c1;
    if (s0 <= ns) offs[s0] = excl;
    if (s1 <= ns) offs[s1] = excl + c0;
    if (tid == 31) offs[kMaxSlices] = inc;
  }
  __syncthreads();
  // SAHI's perform_standard_pred=True (its default, pipe.py:186 passes none): the full-frame prediction is appended
  // to the slice predictions (frame pixels already: no shift); provenance n_slices * max_det + rank
  const int nfull = full_det ? min(full_count[f], max_det) : 0;
  if (tid == 0) cand_count[f] = offs[kMaxSlices] + nfull;    // not clamped: the host can detect overflow past cap
  for (int i = tid; i < nfull; i += blockDim.x) {
    const int slot = offs[kMaxSlices] + i;
    if (slot >= cap) break;
    const float* r = full_det + ((int64_t)f * max_det + i) * 6;
    float2* out = reinterpret_cast<float2*>(cand + ((int64_t)f * cap + slot) * 6);
    out[0] = make_float2(r[0], r[1]); out[1] = make_float2(r[2], r[3]); out[2] = make_float2(r[4], r[5]);
    cand_anchor[(int64_t)f * cap + slot] = ns * max_det + i;
  }
  for (int s = 0; s < ns; ++s) {
    const int base = offs[s], n = (s + 1 <= ns ? offs[s + 1] : offs[kMaxSlices]) - base;
    const float ox = (float)so.x[s], oy = (float)so.y[s];
    const float* rows = det + ((int64_t)(f * ns + s) * max_det) * 6;
    for (int i = tid; i < n; i += blockDim.x) {
      const int slot = base + i;
      if (slot >= cap) break;
      const float2 a = *reinterpret_cast<const float2*>(rows + i * 6), b = *reinterpret_cast<const float2*>(rows + i * 6 + 2),
                   c = *reinterpret_cast<const float2*>(rows + i * 6 + 4);
      float2* out = reinterpret_cast<float2*>(cand + ((int64_t)f * cap + slot) * 6);
      out[0] = make_float2(__fadd_rn(a.x, ox), __fadd_rn(a.y, oy));
      out[1] = make_float2(__fadd_rn(b.x, ox), __fadd_rn(b.y, oy));
      out[2] = c;
      cand_anchor[(int64_t)f * cap + slot] = s * max_det + i;   // provenance: slice and rank inside the slice
    }
  }
}

}  // namespace

extern "C" int b200yolo_gather_slice_detections(const float* det, const int* det_count, int n_frames, int n_slices,
                                                int max_det, const int* slice_xy, const float* full_det,
                                                const int* full_count, float* cand, int* cand_anchor,
                                                int* cand_count, int cap, void* stream) {
  B200_REQUIRE(det && det_count && slice_xy && cand && cand_anchor && cand_count, B200YOLO_ERR_NULL);
  B200_REQUIRE((full_det == nullptr) == (full_count == nullptr), B200YOLO_ERR_NULL);
  B200_REQUIRE(n_frames > 0 && n_slices > 0 && max_det > 0 && cap > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(n_slices <= kMaxSlices, B200YOLO_ERR_UNSUPPORTED);
  SliceOrigins so;
  so.n = n_slices;
  for (int i = 0; i < kMaxSlices; ++i) {
    so.x[i] = i < n_slices ? slice_xy[2 * i] : 0;
    so.y[i] = i < n_slices ? slice_xy[2 * i + 1] : 0;
  }
  gather_slices_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(det, det_count, max_det, so, cand, cand_anchor, cand_count,
                                                                    cap, full_det, full_count);
  return b200_launch_status();
}

// ---------------------------------------------------------------------------------------------------------------
// SAHI's default merge of the slice predictions: postprocess_type = "GREEDYNMM", match_metric = "IOS",
// match_threshold = 0.5, class-aware (sahi/postprocess/combine.py: batched_greedy_nmm + GreedyNMMPostprocess.__call__,
// reached from get_sliced_prediction as the reference calls it, pipe.py:186-193).  Unlike NMS, matched boxes are
// MERGED into the kept one (union box, max score) instead of being dropped:
//   greedy_nmm      candidates in descending score order; the best unconsumed one is kept; every later unconsumed
//                   candidate (of its class) whose metric against the kept box's ORIGINAL box is >= thr is consumed
//                   and queued for merging (fp32 arithmetic, as the torch ops of greedy_nmm);
//   __call__        the queued ones, best first, are merged into the kept prediction one by one while has_match()
//                   holds against the CURRENT (already grown) box: metric > thr in float64 (numpy on python floats);
//                   merged box = union, score = max, category = that of the higher score.
// metric: IOS = inter / min(area_a, area_b), IOU = inter / union.  One CTA per frame; candidates stay in global memory
// (L2), keys are bitonic-sorted in shared memory.  Output in descending score of the kept boxes (SAHI groups its
// list by category first: same set, different order), cut at max_det.  sahi is not installed: parity unpinned.
namespace {

constexpr int kNmmMax = 8192;     // candidates per frame ((n_slices + 1) * max_det, padded to a power of two)
constexpr int kNmmFast = 512;     // frames with at most this many candidates take the parallel form (bit rows in shared memory)
constexpr int kNmmW = kNmmFast / 64;

__device__ __forceinline__ float nmm_metric_f32(const float* a, float area_a, const float* b, float area_b, int ios) {
  const float w = fmaxf(__fsub_rn(fminf(a[2], b[2]), fmaxf(a[0], b[0])), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a[3], b[3]), fmaxf(a[1], b[1])), 0.f);
  const float inter = __fmul_rn(w, h);
  if (ios) return __fdiv_rn(inter, fminf(area_b, area_a));
  return __fdiv_rn(inter, __fadd_rn(__fsub_rn(area_b, inter), area_a));       // (rem_areas - inter) + areas[idx]
}

__device__ __forceinline__ double nmm_metric_f64(const double* a, const float* b, int ios) {
  const double b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3];
  const double area_a = __dmul_rn(__dsub_rn(a[2], a[0]), __dsub_rn(a[3], a[1]));
  const double area_b = __dmul_rn(__dsub_rn(b2, b0), __dsub_rn(b3, b1));
  const double w = fmax(__dsub_rn(fmin(a[2], b2), fmax(a[0], b0)), 0.0);
  const double h = fmax(__dsub_rn(fmin(a[3], b3), fmax(a[1], b1)), 0.0);
  const double inter = __dmul_rn(w, h);
  if (ios) return __ddiv_rn(inter, fmin(area_a, area_b));
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_a, area_b), inter));
}

__global__ void __launch_bounds__(256) greedy_nmm_kernel(const float* __restrict__ cand, const int* __restrict__ cand_src,
                                                         const int* __restrict__ cand_count, int cap, int ios, float thr32,
                                                         double thr64, int agnostic, int max_det, float* __restrict__ out,
                                                         int* __restrict__ out_src, int* __restrict__ out_count,
                                                         const uint32_t* __restrict__ roi_mask, int roi_nc,
                                                         int* __restrict__ roi_cnt) {
  extern __shared__ __align__(16) unsigned char nmm_smem[];
  unsigned long long* key = reinterpret_cast<unsigned long long*>(nmm_smem);          // [kNmmMax]
  int* mlist = reinterpret_cast<int*>(key + kNmmMax);                                // [kNmmMax]
  unsigned char* consumed = reinterpret_cast<unsigned char*>(mlist + kNmmMax);       // [kNmmMax]
  __shared__ int mcount, kept_s, roi_s;
  const int f = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
  const int n = min(min(cand_count[f], cap), kNmmMax);
  const float* rows = cand + (int64_t)f * cap * 6;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  // descending score; among equal scores the later candidate first (greedy_nmm pops the END of an ascending argsort)
  for (int i = tid; i < np2; i += NT) {
    key[i] = i < n ? (((unsigned long long)(~__float_as_uint(rows[i * 6 + 4])) << 32) | (unsigned)(0xffffu - (unsigned)i)) : ~0ull;
    if (i < kNmmMax) consumed[i] = 0;
  }
  if (tid == 0) { kept_s = 0; roi_s = 0; }
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < np2 / 2; t += NT) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
        const unsigned long long a = key[i], c = key[l];
        if ((a > c) == ((i & k) == 0)) { key[i] = c; key[l] = a; }
      }
      __syncthreads();
    }
  auto idx_of = [&](int r) { return (int)(0xffffu - (unsigned)(key[r] & 0xffffu)); };
  if (n <= kNmmFast) {
    // ---- fast path (what a frame's slices + full frame produce after their own NMS: a few hundred candidates) ----
    // greedy_nmm is greedy NMS with a merge step: candidate q is a keeper iff no EARLIER KEEPER r matches it
    // (match = the fp32 test on the ORIGINAL boxes), and a keeper merges the candidates it matches that no earlier
    // keeper took, in rank order.  So: (1) all pair matches up front, in parallel, as bit rows; (2) the keeper set as
    // the fixed point K = all & ~S(K), S(K) = union of the keepers' rows (a candidate is final after the length of its
    // chain of matches: 2-3 rounds); (3) one thread per keeper does its own merge list.  The serial form below spends
    // three CTA barriers per keeper.
    float* sbox = reinterpret_cast<float*>(key + kNmmFast);                                   // [kNmmFast][6] rank order
    unsigned long long* M = reinterpret_cast<unsigned long long*>(sbox + kNmmFast * 6);     // [kNmmFast][kNmmW] match rows
    __shared__ unsigned long long Kw[kNmmW], Sw[kNmmW];
    __shared__ int kbase[kNmmW + 1], changed, nact;
    __shared__ short act[kNmmFast];
    for (int r = tid; r < n; r += NT) {
      const float* b = rows + idx_of(r) * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c) sbox[r * 6 + c] = b[c];
    }
    if (tid < kNmmW) {
      const int lo = tid * 64;
      Kw[tid] = n >= lo + 64 ? ~0ull : (n > lo ? ((1ull << (n - lo)) - 1ull) : 0ull);
      Sw[tid] = 0ull;
    }
    if (tid == 0) nact = 0;
    __syncthreads();
    for (int r = tid; r < n; r += NT) {                        // (1) row r: the later candidates r matches
      const float* bi = sbox + r * 6;
      const float area_i = __fmul_rn(__fsub_rn(bi[2], bi[0]), __fsub_rn(bi[3], bi[1]));
      bool any = false;
      for (int w = 0; w < kNmmW; ++w) {
        unsigned long long bits = 0ull;
        const int q1 = min(n, (w + 1) * 64);
        for (int q = max(r + 1, w * 64); q < q1; ++q) {
          const float* bj = sbox + q * 6;
          if (!agnostic && bj[5] != bi[5]) continue;
          const float area_j = __fmul_rn(__fsub_rn(bj[2], bj[0]), __fsub_rn(bj[3], bj[1]));
          if (!(nmm_metric_f32(bi, area_i, bj, area_j, ios) < thr32)) bits |= 1ull << (q & 63);   // NaN counts as matched
        }
        M[r * kNmmW + w] = bits;
        any = any || bits != 0ull;
      }
      if (any) act[atomicAdd(&nact, 1)] = (short)r;            // rows that can consume anything (few)
    }
    __syncthreads();
    const int na = nact;
    for (;;) {                                                 // (2) fixed point over the keeper set
      if (tid == 0) changed = 0;
      for (int a = tid; a < na; a += NT) {
        const int r = act[a];
        if ((Kw[r >> 6] >> (r & 63)) & 1ull)
          for (int w = r >> 6; w < kNmmW; ++w) { const unsigned long long v = M[r * kNmmW + w]; if (v) atomicOr(&Sw[w], v); }
      }
      __syncthreads();
      if (tid < kNmmW) {
        const int lo = tid * 64;
        const unsigned long long valid = n >= lo + 64 ? ~0ull : (n > lo ? ((1ull << (n - lo)) - 1ull) : 0ull);
        const unsigned long long kn = valid & ~Sw[tid];
        if (kn != Kw[tid]) { Kw[tid] = kn; changed = 1; }
        Sw[tid] = 0ull;
      }
      __syncthreads();
      if (!changed) break;
      __syncthreads();
    }
    if (tid == 0) {
      int acc = 0;
      for (int w = 0; w < kNmmW; ++w) { kbase[w] = acc; acc += __popcll(Kw[w]); }
      kbase[kNmmW] = acc;
    }
    __syncthreads();
    for (int r = tid; r < n; r += NT) {                        // (3) one thread per keeper
      const unsigned long long kw = Kw[r >> 6];
      if (!((kw >> (r & 63)) & 1ull)) continue;
      const int kk = kbase[r >> 6] + __popcll(kw & ((1ull << (r & 63)) - 1ull));
      if (kk >= max_det) continue;
      const float* bi = sbox + r * 6;
      double mb[4] = {(double)bi[0], (double)bi[1], (double)bi[2], (double)bi[3]};
      float mscore = bi[4], mcls = bi[5];
      bool mine = false;
      for (int w = r >> 6; w < kNmmW; ++w) mine = mine || M[r * kNmmW + w] != 0ull;
      if (mine) {
        for (int w = r >> 6; w < kNmmW; ++w) {
          unsigned long long bits = M[r * kNmmW + w];
          if (!bits) continue;
          for (int a = 0; a < na; ++a) {                       // minus what earlier keepers consumed
            const int r2 = act[a];
            if (r2 < r && ((Kw[r2 >> 6] >> (r2 & 63)) & 1ull)) bits &= ~M[r2 * kNmmW + w];
          }
          while (bits) {                                       // rank ascending = best first
            const int q = w * 64 + __ffsll((long long)bits) - 1;
            bits &= bits - 1ull;
            const float* bj = sbox + q * 6;
            if (nmm_metric_f64(mb, bj, ios) > thr64) {
              mb[0] = fmin(mb[0], (double)bj[0]); mb[1] = fmin(mb[1], (double)bj[1]);
              mb[2] = fmax(mb[2], (double)bj[2]); mb[3] = fmax(mb[3], (double)bj[3]);
              if (!(mscore > bj[4])) mcls = bj[5];             // category of the higher score (the later one on a tie)
              mscore = fmaxf(mscore, bj[4]);
            }
          }
        }
      }
      float* o = out + ((int64_t)f * max_det + kk) * 6;
      o[0] = (float)mb[0]; o[1] = (float)mb[1]; o[2] = (float)mb[2]; o[3] = (float)mb[3]; o[4] = mscore; o[5] = mcls;
      out_src[(int64_t)f * max_det + kk] = cand_src[(int64_t)f * cap + idx_of(r)];
      if (roi_cnt) {
        const int c = (int)mcls;
        if (c >= 0 && c < roi_nc && ((roi_mask[c >> 5] >> (c & 31)) & 1u)) atomicAdd(&roi_s, 1);
      }
    }
    __syncthreads();
    if (tid == 0) { out_count[f] = min(kbase[kNmmW], max_det); if (roi_cnt) roi_cnt[f] = roi_s; }
    return;
  }
  for (int r = 0; r < n; ++r) {
    const int i = idx_of(r);
    if (consumed[i]) continue;                                   // uniform: read after a barrier
    if (kept_s >= max_det) break;
    const float* bi = rows + i * 6;
    const float area_i = __fmul_rn(__fsub_rn(bi[2], bi[0]), __fsub_rn(bi[3], bi[1]));
    if (tid == 0) mcount = 0;
    __syncthreads();
    for (int q = r + 1 + tid; q < n; q += NT) {
      const int j = idx_of(q);
      if (consumed[j]) continue;
      const float* bj = rows + j * 6;
      if (!agnostic && bj[5] != bi[5]) continue;
      const float area_j = __fmul_rn(__fsub_rn(bj[2], bj[0]), __fsub_rn(bj[3], bj[1]));
      const float v = nmm_metric_f32(bi, area_i, bj, area_j, ios);
      if (!(v < thr32)) mlist[atomicAdd(&mcount, 1)] = q;       // mask = value < thr; NaN (0/0) counts as matched, as upstream
    }
    __syncthreads();
    if (tid == 0) {
      const int m = mcount;
      for (int a = 1; a < m; ++a) {                              // matched ranks ascending = best first (tiny lists)
        const int v = mlist[a];
        int b = a - 1;
        while (b >= 0 && mlist[b] > v) { mlist[b + 1] = mlist[b]; --b; }
        mlist[b + 1] = v;
      }
      double mb[4] = {(double)bi[0], (double)bi[1], (double)bi[2], (double)bi[3]};
      float mscore = bi[4], mcls = bi[5];
      for (int a = 0; a < m; ++a) {
        const int j = idx_of(mlist[a]);
        consumed[j] = 1;
        const float* bj = rows + j * 6;
        if (nmm_metric_f64(mb, bj, ios) > thr64) {
          mb[0] = fmin(mb[0], (double)bj[0]); mb[1] = fmin(mb[1], (double)bj[1]);
          mb[2] = fmax(mb[2], (double)bj[2]); mb[3] = fmax(mb[3], (double)bj[3]);
          if (!(mscore > bj[4])) mcls = bj[5];                   // category of the higher score (the later one on a tie)
          mscore = fmaxf(mscore, bj[4]);
        }
      }
      const int kk = kept_s;
      float* o = out + ((int64_t)f * max_det + kk) * 6;
      o[0] = (float)mb[0]; o[1] = (float)mb[1]; o[2] = (float)mb[2]; o[3] = (float)mb[3]; o[4] = mscore; o[5] = mcls;
      out_src[(int64_t)f * max_det + kk] = cand_src[(int64_t)f * cap + i];
      if (roi_cnt) {
        const int c = (int)mcls;
        if (c >= 0 && c < roi_nc && ((roi_mask[c >> 5] >> (c & 31)) & 1u)) ++roi_s;
      }
      kept_s = kk + 1;
    }
    __syncthreads();
  }
  if (tid == 0) { out_count[f] = kept_s; if (roi_cnt) roi_cnt[f] = roi_s; }
}

}  // namespace

extern "C" int b200yolo_greedy_nmm(const float* cand, const int* cand_src, const int* cand_count, int n_frames, int cap,
                                   int match_metric, double match_threshold, int agnostic, int max_det, float* out,
                                   int* out_src, int* out_count, const uint32_t* roi_class_mask, int roi_nc, int* roi_cnt,
                                   void* stream) {
  B200_REQUIRE(cand && cand_src && cand_count && out && out_src && out_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(n_frames > 0 && cap > 0 && max_det > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(cap <= kNmmMax, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(match_metric == 0 || match_metric == 1, B200YOLO_ERR_RANGE);
  B200_REQUIRE(match_threshold >= 0.0 && match_threshold <= 1.0, B200YOLO_ERR_RANGE);
  B200_REQUIRE(roi_cnt == nullptr || (roi_class_mask != nullptr && roi_nc > 0), B200YOLO_ERR_NULL);
  const size_t smem = (size_t)kNmmMax * (sizeof(unsigned long long) + sizeof(int) + 1);
  cudaError_t e = cudaFuncSetAttribute(greedy_nmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  greedy_nmm_kernel<<<n_frames, 256, smem, (cudaStream_t)stream>>>(cand, cand_src, cand_count, cap, match_metric,
                                                                (float)match_threshold, match_threshold, agnostic, max_det,
                                                                out, out_src, out_count, roi_class_mask, roi_nc, roi_cnt);
  return b200_launch_status();
}
