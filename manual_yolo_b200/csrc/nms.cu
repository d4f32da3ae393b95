// K4: class-aware greedy NMS with max_det cap (+ optional rescale to source pixels).
//
// Replaces, in ultralytics ops.non_max_suppression (reference entry detect.py:541 / yolo.py:361 /
// pipe.py:179):   c = x[:, 5:6] * (0 if agnostic else max_wh); boxes = x[:, :4] + c
//                 i = torchvision.ops.nms(boxes, scores, iou_thres); i = i[:max_det]; x[i]
// and optionally ops.scale_boxes + clip_boxes.  Restated in oracle/nms.py (nms_numpy_restated follows
// torchvision's CPU nms_kernel: SURVEY.md Appendix A.9 / B.2 / B.7).
//
// Bit-exactness rules implemented here:
//   * offset boxes are formed in fp32, fl32(box + fl32(cls * max_wh)), BEFORE areas / IoU;
//   * inter = max(0, xx2-xx1) * max(0, yy2-yy1); ovr = inter / (ai + aj - inter) with IEEE
//     division; the fp32 ovr is compared against the DOUBLE threshold; NaN (0/0) never suppresses;
//   * kept boxes come out in score order, so stopping after max_det keeps == `i[:max_det]`.
//
// One CTA per image.  The sorted list is consumed in WINDOWS of kWindow boxes: greedy NMS in score order only
// ever needs the prefix it takes to collect max_det keeps (a few hundred boxes even when 8400 pass the confidence
// filter), so neither the gather of the sorted rows nor the suppression sweep touches the rest.  A window lives in
// shared memory (21 B per box); the kept boxes of earlier windows (<= max_det, 20 B each) stay resident and a new
// window is first tested against them.  Inside a window boxes are processed in chunks of 64 (sorted order): (A) the
// 64x64 intra-chunk IoU bitmask is built in parallel (4 pairs per thread, rows merged with warp shuffles), (B) warp 0
// resolves the chunk as a fixed-point iteration over the mask rows (K = alive & ~S(K): 2-4 warp-wide steps of two
// REDUX.OR each, instead of a serial sweep with one dependent load per suppressing box), (C) the chunk's kept boxes
// are applied to every later box of the window in parallel.  ~19 KB of shared memory per CTA: several images per
// SM, so a batch of 256 images is one wave.  Compiled with -fmad=false.

#include "levels.cuh"
#include "nms_common.cuh"

namespace {

using b200::Levels;
using b200::box_meta;
using b200::iou_suppresses;
using b200::may_overlap;
using b200::scale_clip;

constexpr int kChunk = 64;
constexpr int kWindow = 512;           // sorted boxes resident at a time

template <int NT>
__global__ void __launch_bounds__(NT) nms_kernel(const float* __restrict__ cand, const int* __restrict__ cand_anchor,
                                                 const int* __restrict__ cand_count, const int* __restrict__ order,
                                                 int cap, int max_nms, double thr, float max_wh, int agnostic,
                                                 int max_det, const float* __restrict__ scale,
                                                 float* __restrict__ out, int* __restrict__ out_anchor,
                                                 int* __restrict__ out_count, const uint32_t* __restrict__ roi_mask,
                                                 int roi_nc, int* __restrict__ roi_cnt, int* __restrict__ hdr, int B,
                                                 int pass, const Levels L, int decode, float* __restrict__ cand_rw) {
  // decode != 0 (b200yolo_postprocess_dense): the candidate rows hold score and class; the boxes of the first
  // hdr[4B + b] sorted entries were decoded ahead of this kernel (in slot order: DRAM-friendly), the boxes of any
  // later entry are DFL-decoded HERE, when its window is loaded -- so exactly the prefix of the sorted list the greedy
  // NMS consumes is ever read from the head tensor (~500 of 8400 candidates per image in the conf = 0.001 regime).
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float4* box = reinterpret_cast<float4*>(smem_raw);                  // [kWindow] class-offset boxes of the window
  float4* kbox = box + kWindow;                                       // [max_det] kept boxes (all windows)
  int* meta = reinterpret_cast<int*>(kbox + ((max_det + 3) & ~3));    // [kWindow] class id, or -1
  int* kmeta = meta + kWindow;                                        // [max_det]
  int* keep = kmeta + ((max_det + 3) & ~3);                           // [max_det] sorted index of each kept box
  uint8_t* rem = reinterpret_cast<uint8_t*>(keep + ((max_det + 3) & ~3));   // [kWindow]
  __shared__ unsigned long long mask[kChunk];
  __shared__ unsigned rem_bits[2];
  __shared__ unsigned long long kept_bits_s;
  __shared__ int kcount_s;
  __shared__ int roi_s;
  __shared__ unsigned long long cmask[128];   // per chunk: kept boxes by (class & 127); see phase (C)
  __shared__ unsigned long long always_s;      // kept boxes that must be tested against every class (meta -1)

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (pass == 1 && !(hdr && hdr[B + b])) return;       // fallback launch: only the images flagged by pass 0
  const int n_all = min(min(min(cand_count[b], cap), max_nms), B200YOLO_MAX_SORT);
  // hdr[b] = how many entries of order[] the sort put in order (dense regime: the best 2048 only)
  const int n = hdr ? min(n_all, hdr[b]) : n_all;
  const float* crow = cand + (int64_t)b * cap * 6;
  const int* orow = order + (int64_t)b * cap;
  if (n <= 0) {
    if (tid == 0) { out_count[b] = 0; if (roi_cnt) roi_cnt[b] = 0; }
    return;
  }
  if (tid == 0) { kcount_s = 0; roi_s = 0; }
  __syncthreads();

  for (int w0 = 0; w0 < n; w0 += kWindow) {
    const int wn = min(kWindow, n - w0);
    const int kprev = kcount_s;               // keeps collected in earlier windows (uniform: read after a barrier)
    // ---- load the window (decode mode: DFL-decode the boxes that were not decoded ahead, 4 lanes per box) ----
    const int n_pre = decode ? (pass == 0 ? hdr[4 * B + b] : 0) : n;     // sorted entries whose rows hold boxes already
    const bool dec_win = decode && w0 + wn > n_pre;
    if (dec_win) {
      const int r_lo = max(0, n_pre - w0);
      for (int q = tid; q < (((wn - r_lo) * 4 + 31) & ~31); q += NT) {   // whole warps: the quad shuffles need every lane
        const int r = r_lo + (q >> 2), sd = q & 3;
        const bool act = r < wn;
        const int slot = act ? orow[w0 + r] : 0;
        B200_CHECK(slot >= 0 && slot < min(cand_count[b], cap));
        const b200::AnchorRef ar = b200::anchor_ref(L, b, act ? cand_anchor[(int64_t)b * cap + slot] : 0);
        const float d = act ? b200::dfl_side(ar.p + (long long)(sd * b200::kReg) * ar.cs, ar.cs) : 0.f;
        const int q0 = lane & ~3;
        const float d0 = __shfl_sync(0xffffffffu, d, q0), d1 = __shfl_sync(0xffffffffu, d, q0 + 1);
        const float d2 = __shfl_sync(0xffffffffu, d, q0 + 2), d3 = __shfl_sync(0xffffffffu, d, q0 + 3);
        if (act && sd == 0) {
          const float4 bx = b200::decode_box(ar, d0, d1, d2, d3);
          float2* g = reinterpret_cast<float2*>(cand_rw + ((int64_t)b * cap + slot) * 6);   // keep the global rows complete
          g[0] = make_float2(bx.x, bx.y);
          g[1] = make_float2(bx.z, bx.w);
          const float cls = crow[(int64_t)slot * 6 + 5];
          const float rowv[6] = {bx.x, bx.y, bx.z, bx.w, 0.f, cls};
          const float c = agnostic ? 0.f : __fmul_rn(cls, max_wh);
          box[r] = make_float4(__fadd_rn(bx.x, c), __fadd_rn(bx.y, c), __fadd_rn(bx.z, c), __fadd_rn(bx.w, c));
          meta[r] = box_meta(rowv, max_wh, agnostic);
        }
      }
      __syncthreads();
    }
    // ---- test the window against the kept boxes of earlier windows ----
    for (int r = tid; r < wn; r += NT) {
      float4 bj;
      int mj;
      const bool have = dec_win && w0 + r >= n_pre;          // decoded above, already in shared memory
      if (have) {
        bj = box[r]; mj = meta[r];
      } else {
        const float* row = crow + (int64_t)orow[w0 + r] * 6;
        const float c = agnostic ? 0.f : __fmul_rn(row[5], max_wh);
        bj = make_float4(__fadd_rn(row[0], c), __fadd_rn(row[1], c), __fadd_rn(row[2], c), __fadd_rn(row[3], c));
        mj = box_meta(row, max_wh, agnostic);
      }
      bool dead = false;
      for (int k = 0; k < kprev && !dead; ++k) {
        if (!may_overlap(kmeta[k], mj)) continue;
        const float4 bi = kbox[k];
        const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
        dead = iou_suppresses(bi, ai, bj, thr);
      }
      if (!have) { box[r] = bj; meta[r] = mj; }
      rem[r] = dead ? 1 : 0;
    }
    __syncthreads();

    bool done = false;
    for (int s = 0; s < wn; s += kChunk) {
      const int m = min(kChunk, wn - s);
      // ---- (A) intra-chunk mask: thread -> row i = t/16, columns j = (t%16)*4 .. +3 ----
      if (tid < 128) cmask[tid] = 0ull;                    // phase (C)'s class buckets: idle until this chunk's (B) is done
      if (tid == 128 % NT) always_s = 0ull;
      for (int t = tid; t < kChunk * 16; t += NT) {
        const int i = t >> 4, jg = t & 15;
        unsigned nib = 0;
        if (i < m) {
          const float4 bi = box[s + i];
          const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = jg * 4 + u;
            if (j > i && j < m && may_overlap(meta[s + i], meta[s + j]) && iou_suppresses(bi, ai, box[s + j], thr))
              nib |= 1u << u;
          }
        }
        unsigned lo = jg < 8 ? nib << (jg * 4) : 0u;
        unsigned hi = jg >= 8 ? nib << ((jg - 8) * 4) : 0u;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          lo |= __shfl_xor_sync(0xffffffffu, lo, o);
          hi |= __shfl_xor_sync(0xffffffffu, hi, o);
        }
        if (jg == 0) mask[i] = ((unsigned long long)hi << 32) | lo;
      }
      if (wid < 2) {
        const int j = wid * 32 + lane;
        const unsigned bits = __ballot_sync(0xffffffffu, j < m && rem[s + j] != 0);
        if (lane == 0) rem_bits[wid] = bits;
      }
      __syncthreads();
      // ---- (B) resolve of the chunk by warp 0, as a fixed-point iteration instead of a serial sweep: with
      //      S(K) = union of the mask rows of the boxes in K, the greedy result is the unique K with
      //      K = alive & ~S(K) (box j depends on earlier boxes only).  K_0 = alive, K_{t+1} = alive & ~S(K_t): box j is
      //      final after at most (length of its suppression chain) steps -- 2-4 warp-wide steps (two REDUX.OR each)
      //      for a typical chunk, against one dependent shared-memory load per suppressing box. ----
      if (tid < 32) {
        const unsigned long long r0 = lane < m ? mask[lane] : 0ull, r1 = lane + 32 < m ? mask[lane + 32] : 0ull;
        const unsigned long long valid = m == 64 ? ~0ull : ((1ull << m) - 1ull);
        const unsigned long long alive = valid & ~(((unsigned long long)rem_bits[1] << 32) | rem_bits[0]);
        unsigned long long K = alive;
        for (;;) {
          const unsigned long long sl = (((K >> lane) & 1ull) ? r0 : 0ull) | (((K >> (lane + 32)) & 1ull) ? r1 : 0ull);
          const unsigned long long S = ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(sl >> 32)) << 32) |
                                       __reduce_or_sync(0xffffffffu, (unsigned)sl);
          const unsigned long long Kn = alive & ~S;
          if (Kn == K) break;
          K = Kn;
        }
        int kc = kcount_s;
        int over = kc + __popcll(K) - max_det;           // keeps beyond max_det: drop the last ones (score order)
        while (over > 0) { K &= ~(1ull << (63 - __clzll((long long)K))); --over; }
        if (lane == 0) { kept_bits_s = K; kcount_s = kc + __popcll(K); }
      }
      __syncthreads();
      {   // record the chunk's keeps in parallel: rank inside the chunk = number of kept boxes before it
        const unsigned long long kept = kept_bits_s;
        if (tid < kChunk && ((kept >> tid) & 1ull)) {
          const int r = kcount_s - __popcll(kept) + __popcll(kept & ((1ull << tid) - 1ull));
          B200_CHECK(r >= 0 && r < max_det && s + tid < wn);
          keep[r] = w0 + s + tid;
          kbox[r] = box[s + tid];
          kmeta[r] = meta[s + tid];
        }
      }
      if (kcount_s >= max_det) { done = true; break; }
      // ---- (C) apply this chunk's kept boxes to all later boxes of the window ----
      // Class-aware shortcut (exact, see box_meta): a later box only needs the kept boxes of its own class.
      // The chunk's kept boxes are bucketed by (class & 127) into 64-bit masks, so a thread visits 0-2
      // kept boxes per later box instead of up to 64.
      const unsigned long long kept = kept_bits_s;
      if (kept && s + kChunk < wn) {
        if (tid < kChunk && ((kept >> tid) & 1ull)) {     // (cmask / always_s were cleared in phase (A))
          const int mi = meta[s + tid];
          if (mi < 0) atomicOr(&always_s, 1ull << tid);
          else atomicOr(&cmask[mi & 127], 1ull << tid);
        }
        __syncthreads();
        for (int j = s + kChunk + tid; j < wn; j += NT) {
          if (rem[j]) continue;
          const float4 bj = box[j];
          const int mj = meta[j];
          unsigned long long kb = mj >= 0 ? (cmask[mj & 127] | always_s) : kept;
          while (kb) {
            const int i = __ffsll((long long)kb) - 1;
            kb &= kb - 1;
            if (!may_overlap(meta[s + i], mj)) continue;   // bucket collision (class & 127)
            const float4 bi = box[s + i];
            const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
            if (iou_suppresses(bi, ai, bj, thr)) { rem[j] = 1; break; }
          }
        }
      }
      __syncthreads();
    }
    if (done) break;
    __syncthreads();
  }
  __syncthreads();
  const int kc = kcount_s;
  if (n < n_all && kc < max_det) {
    // the ordered prefix ran out before max_det keeps: this image needs the full sort.  Flag it; b200yolo_nms
    // launches the fallback pair (full sort + this kernel with pass = 1) right behind this launch.
    if (tid == 0) hdr[B + b] = 1;
    return;
  }
  // ---- outputs: un-offset rows in kept order ----
  float gain = 1.f, padx = 0.f, pady = 0.f, w0s = 0.f, h0s = 0.f;
  if (scale) { gain = scale[b * 5 + 0]; padx = scale[b * 5 + 1]; pady = scale[b * 5 + 2]; w0s = scale[b * 5 + 3]; h0s = scale[b * 5 + 4]; }
  for (int r = tid; r < kc; r += NT) {
    const int slot = orow[keep[r]];
    B200_CHECK(keep[r] >= 0 && keep[r] < n && slot >= 0 && slot < min(cand_count[b], cap));
    const float* row = crow + (int64_t)slot * 6;
    float x1 = row[0], y1 = row[1], x2 = row[2], y2 = row[3];
    if (scale) {
      x1 = scale_clip(x1, padx, gain, w0s); y1 = scale_clip(y1, pady, gain, h0s);
      x2 = scale_clip(x2, padx, gain, w0s); y2 = scale_clip(y2, pady, gain, h0s);
    }
    float* o = out + ((int64_t)b * max_det + r) * 6;
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2; o[4] = row[4]; o[5] = row[5];
    out_anchor[(int64_t)b * max_det + r] = cand_anchor[(int64_t)b * cap + slot];
    if (roi_cnt) {  // kept detections whose class feeds the ROI stage (the *_rank ids)
      const int c = (int)row[5];
      if (c >= 0 && c < roi_nc && ((roi_mask[c >> 5] >> (c & 31)) & 1u)) atomicAdd(&roi_s, 1);
    }
  }
  __syncthreads();
  if (tid == 0) { out_count[b] = kc; if (roi_cnt) roi_cnt[b] = roi_s; }
}

__global__ void scale_boxes_kernel(float* boxes, int n, int row_stride, float gain, float padx, float pady,
                                   float w0, float h0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* r = boxes + (int64_t)i * row_stride;
  r[0] = scale_clip(r[0], padx, gain, w0); r[1] = scale_clip(r[1], pady, gain, h0);
  r[2] = scale_clip(r[2], padx, gain, w0); r[3] = scale_clip(r[3], pady, gain, h0);
}

}  // namespace

struct NmsArgs {
  const float* cand; const int* cand_anchor; const int* cand_count; const int* order;
  int B, cap, max_nms; double iou_thres; float max_wh; int agnostic, max_det;
  const float* scale; float* out; int* out_anchor; int* out_count;
  const uint32_t* roi_class_mask; int roi_nc; int* roi_cnt;
};

static int nms_check(const NmsArgs& a) {
  B200_REQUIRE(a.cand && a.cand_anchor && a.cand_count && a.order && a.out && a.out_anchor && a.out_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(a.B > 0 && a.cap > 0 && a.max_nms > 0 && a.max_det > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(a.cap <= B200YOLO_MAX_SORT && a.max_det <= 4096, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(a.iou_thres >= 0.0 && a.iou_thres <= 1.0, B200YOLO_ERR_RANGE);
  B200_REQUIRE(a.roi_cnt == nullptr || (a.roi_class_mask != nullptr && a.roi_nc > 0), B200YOLO_ERR_NULL);
  return B200YOLO_OK;
}

static int nms_launch(const NmsArgs& a, int* hdr, int pass, cudaStream_t s, const Levels* levels = nullptr) {
  Levels L;
  if (levels) L = *levels;
  else {
    for (int l = 0; l < B200YOLO_MAX_LEVELS; ++l) {
      L.ptr[l] = nullptr; L.bstride[l] = 0; L.cstride[l] = 0; L.w[l] = 1; L.stride[l] = 1.f; L.off[l] = 0;
    }
    L.off[B200YOLO_MAX_LEVELS] = 0; L.n = 1;
  }
  const int decode = levels ? 1 : 0;
  float* cand_rw = const_cast<float*>(a.cand);
  const size_t md4 = ((size_t)a.max_det + 3) & ~(size_t)3;
  const size_t smem = (size_t)kWindow * 16 + md4 * 16 + (size_t)kWindow * 4 + md4 * 4 + md4 * 4 + (size_t)kWindow;
  if (a.cap <= 1024) {
    constexpr int NT = 256;
    auto kern = nms_kernel<NT>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    kern<<<a.B, NT, smem, s>>>(a.cand, a.cand_anchor, a.cand_count, a.order, a.cap, a.max_nms, a.iou_thres, a.max_wh,
                               a.agnostic, a.max_det, a.scale, a.out, a.out_anchor, a.out_count, a.roi_class_mask,
                               a.roi_nc, a.roi_cnt, hdr, a.B, pass, L, decode, cand_rw);
  } else {
    constexpr int NT = 512;
    auto kern = nms_kernel<NT>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    kern<<<a.B, NT, smem, s>>>(a.cand, a.cand_anchor, a.cand_count, a.order, a.cap, a.max_nms, a.iou_thres, a.max_wh,
                               a.agnostic, a.max_det, a.scale, a.out, a.out_anchor, a.out_count, a.roi_class_mask,
                               a.roi_nc, a.roi_cnt, hdr, a.B, pass, L, decode, cand_rw);
  }
  return b200_launch_status();
}

extern "C" int b200yolo_nms(const float* cand, const int* cand_anchor, const int* cand_count, const int* order,
                            int B, int cap, int max_nms, double iou_thres, float max_wh, int agnostic, int max_det,
                            const float* scale, float* out, int* out_anchor, int* out_count,
                            const uint32_t* roi_class_mask, int roi_nc, int* roi_cnt, void* workspace,
                            size_t workspace_bytes, void* stream) {
  const NmsArgs a{cand, cand_anchor, cand_count, order, B, cap, max_nms, iou_thres, max_wh, agnostic, max_det,
                  scale, out, out_anchor, out_count, roi_class_mask, roi_nc, roi_cnt};
  int rc = nms_check(a);
  if (rc != B200YOLO_OK) return rc;
  // workspace header (see sort_topk.cu): present iff the same workspace was given to b200yolo_sort_topk
  const bool have_hdr = workspace != nullptr && workspace_bytes >= b200yolo_workspace_bytes(B, cap);
  if (workspace) B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, B200YOLO_ERR_ALIGN);
  int* hdr = have_hdr ? reinterpret_cast<int*>(workspace) : nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  rc = nms_launch(a, hdr, 0, s);
  if (rc != B200YOLO_OK || !have_hdr || cap <= 2048) return rc;
  // dense regime: the sort ordered only the best 2048 entries per image.  Images whose NMS ran out of them (flag
  // raised by pass 0) are re-done exactly: full sort, then NMS again -- both launches exit at once otherwise.
  rc = b200_sort_launch(cand, cand_anchor, cand_count, B, cap, max_nms, const_cast<int*>(order), workspace,
                        workspace_bytes, 1, s);
  if (rc != B200YOLO_OK) return rc;
  return nms_launch(a, hdr, 1, s);
}

// K2b + K3 + K4 for the dense regime (cap > 1024, e.g. conf = 0.001 evaluation: every anchor is a candidate).
// The candidates come from b200yolo_class_filter (scores and classes only).  Sorting needs no boxes, and greedy NMS
// stops at max_det keeps, so the work is ordered to touch only what the NMS consumes: select + sort the best 2048
// entries per image, then the NMS kernel DFL-decodes each window of 512 sorted candidates as it loads it -- about
// 500 boxes per image in the conf = 0.001 regime instead of all 8400 (or the 2048 ordered ones).  An image whose NMS
// runs out of ordered entries is flagged and re-done after the full sort (same stream, launches exit at once otherwise).
extern "C" int b200yolo_postprocess_dense(const b200yolo_level* levels, int n_levels, float* cand,
                                          const int* cand_anchor, const int* cand_count, int B, int cap, int max_nms,
                                          double iou_thres, float max_wh, int agnostic, int max_det,
                                          const float* scale, float* out, int* out_anchor, int* out_count,
                                          const uint32_t* roi_class_mask, int roi_nc, int* roi_cnt, int* order,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  const NmsArgs a{cand, cand_anchor, cand_count, order, B, cap, max_nms, iou_thres, max_wh, agnostic, max_det,
                  scale, out, out_anchor, out_count, roi_class_mask, roi_nc, roi_cnt};
  int rc = nms_check(a);
  if (rc != B200YOLO_OK) return rc;
  B200_REQUIRE(levels && workspace, B200YOLO_ERR_NULL);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, B200YOLO_ERR_ALIGN);
  B200_REQUIRE(workspace_bytes >= b200yolo_workspace_bytes(B, cap), B200YOLO_ERR_WORKSPACE);
  int* hdr = reinterpret_cast<int*>(workspace);
  cudaStream_t s = (cudaStream_t)stream;
  Levels L;
  rc = b200::build_levels(levels, n_levels, L);
  if (rc != B200YOLO_OK) return rc;
  for (int pass = 0; pass < (cap > 2048 ? 2 : 1); ++pass) {
    rc = b200_sort_launch(cand, cand_anchor, cand_count, B, cap, max_nms, order, workspace, workspace_bytes, pass, s);
    if (rc != B200YOLO_OK) return rc;
    if (pass == 0) {   // boxes of the best 512 entries per image, in slot order
      rc = b200_box_decode_sorted_launch(levels, n_levels, cand, cand_anchor, cand_count, order, B, cap, max_nms, hdr, 0, s);
      if (rc != B200YOLO_OK) return rc;
    }
    rc = nms_launch(a, hdr, pass, s, &L);       // decodes the boxes of any further window it consumes
    if (rc != B200YOLO_OK) return rc;
  }
  return B200YOLO_OK;
}

extern "C" int b200yolo_scale_boxes(float* boxes, int n, int row_stride, float gain, float pad_x, float pad_y,
                                    float w0, float h0, void* stream) {
  B200_REQUIRE(n >= 0 && gain > 0.f, B200YOLO_ERR_SHAPE);
  if (n == 0) return B200YOLO_OK;           // an empty tensor has a NULL data pointer: nothing to do, not an error
  B200_REQUIRE(boxes, B200YOLO_ERR_NULL);
  B200_REQUIRE(row_stride >= 4, B200YOLO_ERR_SHAPE);
  scale_boxes_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(boxes, n, row_stride, gain, pad_x, pad_y,
                                                                        w0, h0);
  return b200_launch_status();
}
