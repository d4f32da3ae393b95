// K5: batched ROI crop + resize to size x size (the rank-classifier input batch), and ROI selection.
//
// Replaces, per detection, the reference's   x1,y1,x2,y2 = map(int, xyxy)      (detect.py:581)
//                                            crop = safe_crop(frame, .., pad=6)   (detect.py:100-113,586)
//                                            rank_model(crop)  preprocessing      (detect.py:121)
// where the preprocessing is ultralytics ClassificationPredictor.preprocess with the checkpoint's
// transforms: BGR->RGB, torchvision Resize(size) on a PIL image (short side -> size, bilinear with
// antialias), CenterCrop(size), ToTensor (/255), Normalize(0,1).  Pillow resamples 8-bit images in
// two passes (horizontal then vertical), each with 22-bit fixed-point coefficients and a round to
// uint8 in between; oracle/roi.py::pil_resize_restated is the specification followed here
// (SURVEY.md Appendix A.11 / B.4).  All size arithmetic (double division, banker's rounding of the
// crop offset) is done on the device exactly as Python/C do it on the host.
//
// One CTA per ROI: 128 threads build the two coefficient tables (only the 64 output columns/rows
// that survive the centre crop), the horizontal pass writes a uint8 strip into shared memory, the
// vertical pass streams coalesced fp32 rows to the planar (3,size,size) output.
// HBM-bound: algorithmic bytes per ROI = crop_h*crop_w*3 read + 3*size*size*4 written.

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSize = 64;     // output side supported by the shared-memory plan
constexpr int kMaxTaps = 64;     // ksize = ceil(support)*2+1 <= 64  <=> scale <= 31
constexpr int kRowsMax = 192;    // uint8 strip rows staged per vertical tile
constexpr int kPrec = 22;        // Pillow PRECISION_BITS = 32 - 8 - 2

struct AxisPlan {
  int bounds[kMaxSize][2];          // xmin, count
  int coef[kMaxSize][kMaxTaps];
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output index `o` (triangle filter).
__device__ void pil_axis(int o, int in_size, int out_size, int* bounds, int* coef) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = fscale;  // bilinear support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, fscale);
  const double center = __dmul_rn((double)o + 0.5, scale);
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kMaxTaps) xmax = kMaxTaps;  // unreachable inside the supported envelope
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double v = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (v < 0.0) v = -v;
    const double w = v < 1.0 ? __dsub_rn(1.0, v) : 0.0;
    ww = __dadd_rn(ww, w);
  }
  for (int x = 0; x < xmax; ++x) {
    double v = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (v < 0.0) v = -v;
    double w = v < 1.0 ? __dsub_rn(1.0, v) : 0.0;
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    coef[x] = __double2int_rz(__dadd_rn(0.5, __dmul_rn(w, (double)(1 << kPrec))));
  }
  bounds[0] = xmin;
  bounds[1] = xmax;
}

__device__ __forceinline__ int clip8(int v) {
  v >>= kPrec;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// python round() of d/2 for integer d >= 0 (round half to even)
__device__ __forceinline__ int half_round_even(int d) {
  const int k = d >> 1;
  return (d & 1) ? ((k & 1) ? k + 1 : k) : k;
}

__global__ void __launch_bounds__(kThreads) roi_kernel(const uint8_t* __restrict__ frames, int B, int H, int W,
                                                       int64_t pitch, int64_t bstride,
                                                       const float* __restrict__ boxes,
                                                       const int* __restrict__ batch_idx,
                                                       const int* __restrict__ roi_count, int pad, int S,
                                                       float* __restrict__ dst, int* __restrict__ valid) {
  extern __shared__ __align__(16) uint8_t roi_smem[];
  AxisPlan& px = *reinterpret_cast<AxisPlan*>(roi_smem);
  AxisPlan& py = *reinterpret_cast<AxisPlan*>(roi_smem + sizeof(AxisPlan));
  uint8_t (*strip)[kMaxSize][3] = reinterpret_cast<uint8_t (*)[kMaxSize][3]>(roi_smem + 2 * sizeof(AxisPlan));
  const int r = blockIdx.x, tid = threadIdx.x;
  if (roi_count != nullptr && r >= *roi_count) return;
  float* out = dst + (int64_t)r * 3 * S * S;

  // ---- integer crop geometry: int() truncation then safe_crop (detect.py:100-113) ----
  const int bi = batch_idx[r];
  const int bx1 = __float2int_rz(boxes[r * 4 + 0]), by1 = __float2int_rz(boxes[r * 4 + 1]);
  const int bx2 = __float2int_rz(boxes[r * 4 + 2]), by2 = __float2int_rz(boxes[r * 4 + 3]);
  const int cx1 = max(0, min(W - 1, bx1 - pad)), cx2 = max(0, min(W, bx2 + pad));
  const int cy1 = max(0, min(H - 1, by1 - pad)), cy2 = max(0, min(H, by2 + pad));
  const int cw = cx2 - cx1, ch = cy2 - cy1;
  bool ok = (bi >= 0 && bi < B && cw > 0 && ch > 0), unsupported = false;
  // torchvision Resize(int): short side -> S, long side -> int(S * long / short)
  int new_w = S, new_h = S;
  if (ok) {
    if (cw <= ch) new_h = __double2int_rz(__ddiv_rn((double)(S * (int64_t)ch), (double)cw));
    else new_w = __double2int_rz(__ddiv_rn((double)(S * (int64_t)cw), (double)ch));
    // supported envelope of the shared-memory plan: ksize = ceil(support)*2+1 <= kMaxTaps on both axes
    const double sx = (double)cw / (double)new_w, sy = (double)ch / (double)new_h;
    const double sm = fmax(fmax(sx, sy), 1.0);
    if (2 * (int)ceil(sm) + 1 > kMaxTaps) { ok = false; unsupported = true; }
  }
  if (!ok) {
    for (int i = tid; i < 3 * S * S; i += kThreads) out[i] = 0.f;
    if (tid == 0) valid[r] = unsupported ? -1 : 0;
    return;
  }
  const int left = half_round_even(new_w - S), top = half_round_even(new_h - S);

  if (tid < S) pil_axis(left + tid, cw, new_w, px.bounds[tid], px.coef[tid]);
  else if (tid >= 64 && tid < 64 + S) pil_axis(top + (tid - 64), ch, new_h, py.bounds[tid - 64], py.coef[tid - 64]);
  __syncthreads();

  const uint8_t* crop = frames + (int64_t)bi * bstride + (int64_t)cy1 * pitch + (int64_t)cx1 * 3;
  int t0 = 0;
  while (t0 < S) {
    // vertical tile [t0, t1): input rows [rmin, rmax) must fit the strip
    const int rmin = py.bounds[t0][0];
    int t1 = t0 + 1;
    while (t1 < S && py.bounds[t1][0] + py.bounds[t1][1] - rmin <= kRowsMax) ++t1;
    const int rmax = py.bounds[t1 - 1][0] + py.bounds[t1 - 1][1];
    const int rows = rmax - rmin;
    // ---- horizontal pass: strip[row][xx][c] = clip8(2^21 + sum px * k) ----
    for (int e = tid; e < rows * S * 3; e += kThreads) {
      const int c = e % 3, xx = (e / 3) % S, rr = e / (3 * S);
      const int xmin = px.bounds[xx][0], cnt = px.bounds[xx][1];
      const uint8_t* p = crop + (int64_t)(rmin + rr) * pitch + xmin * 3 + c;
      int acc = 1 << (kPrec - 1);
      for (int x = 0; x < cnt; ++x) acc += (int)__ldg(p + x * 3) * px.coef[xx][x];
      strip[rr][xx][c] = (uint8_t)clip8(acc);
    }
    __syncthreads();
    // ---- vertical pass + BGR->RGB + /255 ----
    const int trows = t1 - t0;
    for (int e = tid; e < 3 * trows * S; e += kThreads) {
      const int xx = e % S, yy = t0 + (e / S) % trows, c = e / (S * trows);
      const int ymin = py.bounds[yy][0], cnt = py.bounds[yy][1];
      int acc = 1 << (kPrec - 1);
      for (int y = 0; y < cnt; ++y) acc += (int)strip[ymin - rmin + y][xx][c] * py.coef[yy][y];
      out[((2 - c) * S + yy) * S + xx] = b200::u8_div255(clip8(acc));
    }
    __syncthreads();
    t0 = t1;
  }
  if (tid == 0) valid[r] = 1;
}

// ---- ROI selection: detections of the allowed classes -> dense list, image-major, order kept ----
__global__ void __launch_bounds__(1024) select_rois_kernel(const float* __restrict__ det,
                                                           const int* __restrict__ det_count, int B, int max_det,
                                                           const uint32_t* __restrict__ class_mask, int nc,
                                                           float* __restrict__ roi_boxes, int* __restrict__ roi_batch,
                                                           int* __restrict__ roi_det, int* __restrict__ roi_count,
                                                           int roi_cap) {
  extern __shared__ int offs[];  // [B + 1]
  __shared__ int warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  auto wanted = [&](int b, int i) -> bool {
    const int c = (int)det[((int64_t)b * max_det + i) * 6 + 5];
    return c >= 0 && c < nc && ((class_mask[c >> 5] >> (c & 31)) & 1u);
  };
  // pass 1: per-image counts (one warp per image)
  for (int b = wid; b < B; b += 32) {
    const int n = min(det_count[b], max_det);
    int cnt = 0;
    for (int i = lane; i < n; i += 32) cnt += wanted(b, i);
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) offs[b] = cnt;
  }
  __syncthreads();
  // exclusive scan over images (B is small: chunked block scan)
  int carry = 0;
  for (int base = 0; base < B; base += 1024) {
    const int b = base + tid;
    const int v = b < B ? offs[b] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int t = warp_tot[lane];
      int ti = t;
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += u;
      }
      warp_tot[lane] = ti - t;
    }
    __syncthreads();
    const int excl = carry + warp_tot[wid] + inc - v;
    if (b < B) offs[b] = excl;
    __syncthreads();
    if (tid == 1023) warp_tot[0] = excl + v;  // running total
    __syncthreads();
    carry = warp_tot[0];
    __syncthreads();
  }
  if (tid == 0) *roi_count = min(carry, roi_cap);
  // pass 2: ordered scatter
  for (int b = wid; b < B; b += 32) {
    const int n = min(det_count[b], max_det);
    int pos = offs[b];
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      const bool w = i < n && wanted(b, i);
      const unsigned bal = __ballot_sync(0xffffffffu, w);
      if (w) {
        const int slot = pos + __popc(bal & ((1u << lane) - 1u));
        if (slot < roi_cap) {
          const float* row = det + ((int64_t)b * max_det + i) * 6;
          roi_boxes[slot * 4 + 0] = row[0]; roi_boxes[slot * 4 + 1] = row[1];
          roi_boxes[slot * 4 + 2] = row[2]; roi_boxes[slot * 4 + 3] = row[3];
          roi_batch[slot] = b;
          roi_det[slot] = i;
        }
      }
      pos += __popc(bal);
    }
  }
}

}  // namespace

extern "C" int b200yolo_roi_crop_resize(const uint8_t* frames, int B, int H, int W, int64_t pitch,
                                        int64_t batch_stride, const float* boxes, const int* batch_idx,
                                        const int* roi_count, int N, int pad, int size, float* dst, int* valid,
                                        void* stream) {
  B200_REQUIRE(frames && boxes && batch_idx && dst && valid, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && N >= 0 && pad >= 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(pitch >= (int64_t)W * 3 && (B == 1 || batch_stride >= pitch * (int64_t)(H - 1) + (int64_t)W * 3), B200YOLO_ERR_SHAPE);
  B200_REQUIRE(size > 0 && size <= kMaxSize, B200YOLO_ERR_UNSUPPORTED);
  if (N == 0) return B200YOLO_OK;
  const size_t smem = 2 * sizeof(AxisPlan) + (size_t)kRowsMax * kMaxSize * 3;
  cudaError_t e = cudaFuncSetAttribute(roi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  roi_kernel<<<N, kThreads, smem, (cudaStream_t)stream>>>(frames, B, H, W, pitch, batch_stride, boxes, batch_idx,
                                                       roi_count, pad, size, dst, valid);
  return b200_launch_status();
}

extern "C" int b200yolo_select_rois(const float* det, const int* det_count, int B, int max_det,
                                    const uint32_t* class_mask, int nc, float* roi_boxes, int* roi_batch,
                                    int* roi_det, int* roi_count, int roi_cap, void* stream) {
  B200_REQUIRE(det && det_count && class_mask && roi_boxes && roi_batch && roi_det && roi_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && max_det > 0 && nc > 0 && roi_cap > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(B <= 8192, B200YOLO_ERR_UNSUPPORTED);
  select_rois_kernel<<<1, 1024, (B + 1) * sizeof(int), (cudaStream_t)stream>>>(
      det, det_count, B, max_det, class_mask, nc, roi_boxes, roi_batch, roi_det, roi_count, roi_cap);
  return b200_launch_status();
}
