// K1: fused letterbox (cv2 INTER_LINEAR 8-bit fixed point) + pad + BGR->RGB + /255 + NCHW.
//
// Replaces ultralytics LetterBox.__call__ + BasePredictor.preprocess (reference entry
// detect.py:541, yolo.py:361, pipe.py:179); arithmetic restated in oracle/letterbox.py
// (SURVEY.md Appendix B.1).
//
// One CTA produces kRows (8) consecutive output rows of one frame; thread t owns output pixels 4t..4t+3 of
// every row, so its horizontal taps (double/float arithmetic exactly as cv::resize builds its tables)
// are derived ONCE and stay in registers.  The (at most two) source rows each output row references are
// staged in shared memory by the TMA engine (cp.async.bulk 1-D copies completing on an mbarrier;
// SASS UBLKCP) into a two-stage ring: row r+1 is in flight while row r is blended.  Rows that are not
// 16-byte aligned fall back to cooperative loads.  Per pixel: three 32-bit LDS + funnel shifts fetch the
// two BGR taps, cv2's fixed-point blend (or a pass-through when the scale is an exact integer and the
// weights degenerate to 2048/0), an exact u8 -> v/255 table lookup, three 128-bit streaming stores
// (planar RGB) -- or 12 interleaved bytes for the u8 variant.
// HBM-bound: algorithmic bytes per frame = referenced rows * W*3 + 3*outH*outW*4.

#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int kMaxSlices = 64;   // slices per frame in slice mode (a 1920x1200 frame at 640/0.2 has 12)

struct LbParams {
  const uint8_t* src;
  void* dst;
  int64_t pitch, bstride;
  double scale_x, scale_y;
  int H, W, outH, outW, new_w, new_h, top, left, pad_value, swap_rb;
  int row_bytes;      // W*3
  int row_smem;       // bytes reserved per staged row (>= row_bytes + 16, multiple of 16)
  int bulk_ok;        // rows are 16-byte aligned and row_bytes % 16 == 0
  // slice mode (SAHI-style tiling, pipe.py:183-194): batch item b is slice b % ns of frame b / ns, a H x W window
  // of the frame at (sx, sy); ns == 1 with a zero origin is the plain per-frame form
  int ns;
  int sx[kMaxSlices], sy[kMaxSlices];
};

// cv::resize table entry for one axis: source index and 11-bit weights (a0 for s, a1 for s+1).
__device__ __forceinline__ void cv_linear_tap(int d, double scale, int ssize, int& s, int& a0, int& a1) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (s < 0) { f = 0.f; s = 0; }
  if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// Vertical axis: cv::resize keeps the fractional weights of out-of-range taps and only clamps the
// source ROW indices in its row loop (matters on up-scales: the first/last rows blend one source
// row twice, with two separately truncated products).
__device__ __forceinline__ void cv_linear_tap_v(int d, double scale, int ssize, int& r0, int& r1, int& b0, int& b1) {
  float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  const int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  b1 = __float2int_rn(__fmul_rn(f, 2048.f));
  r0 = min(max(s, 0), ssize - 1);
  r1 = min(max(s + 1, 0), ssize - 1);
}

// 6 consecutive bytes (two BGR pixels) starting at byte offset `off` of a staged row.
__device__ __forceinline__ void load6(const uint8_t* row, int off, uint32_t& lo, uint32_t& hi) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(row) + (off >> 2);
  uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
  uint32_t sh = (off & 3) * 8;
  lo = __funnelshift_r(w0, w1, sh);
  hi = __funnelshift_r(w1, w2, sh);
}

template <typename OutT>
__device__ __forceinline__ OutT lb_cast(const float* lut, int v);
template <>
__device__ __forceinline__ float lb_cast<float>(const float* lut, int v) { return lut[v]; }   // exact v/255 table
template <>
__device__ __forceinline__ uint8_t lb_cast<uint8_t>(const float*, int v) { return (uint8_t)v; }
// half=True (UL `im.half(); im /= 255` on the device): torch divides in fp32 and rounds the quotient to fp16
template <>
__device__ __forceinline__ __half lb_cast<__half>(const float* lut, int v) { return __float2half_rn(lut[v]); }

#ifndef B200_LB_ROWS
// Rows per CTA, measured on 64 frames 1920x1200 -> 640x640 (us per launch of the kernel ALONE): 1: 111.5, 2: 87.0,
// 3: 82.4, 4: 76.8 (0.92 of the copy peak), 5: 78.7, 6: 78.2, 8: 79.0 (0.89), 16: 82.9, 32: 89.2.  The whole step
// (letterbox concurrent with class filter -> post-processing -> ROI crops, two batches in flight) is fastest with 8:
// 572 k frames/s against 570 k (6 rows) and 561 k (4 rows) -- the extra per-CTA set-up of short CTAs takes issue slots
// from the latency-bound kernels running beside them.  8 it is: the pipeline is the product, not the kernel alone.
#define B200_LB_ROWS 8
#endif
#ifndef B200_LB_STAGES
#define B200_LB_STAGES 2
#endif
constexpr int kRows = B200_LB_ROWS;       // output rows per CTA
constexpr int kStages = B200_LB_STAGES;   // ring slots: row r+1 is in flight while row r is blended

struct RowInfo { int r0, r1, b0, b1; };   // b0 < 0 marks a padding row

// Per-thread output cursor: three plane pointers (f32) or one interleaved pointer (u8), advanced per row.
template <typename OutT>
struct OutCursor;
template <>
struct OutCursor<float> {
  float *q0, *q1, *q2;     // planes of source channels 0,1,2 (B,G,R): swapped to R,G,B order when swap_rb
  int64_t row_stride;
  bool vec;
  __device__ __forceinline__ void init(const LbParams& p, int b, int oy, int ox, int n) {
    float* base = reinterpret_cast<float*>(p.dst) + ((int64_t)b * 3 * p.outH + oy) * p.outW + ox;
    const int64_t plane = (int64_t)p.outH * p.outW;
    q0 = base + (p.swap_rb ? 2 : 0) * plane;
    q1 = base + plane;
    q2 = base + (p.swap_rb ? 0 : 2) * plane;
    row_stride = p.outW;
    vec = (n == 4) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((plane & 3) == 0) && ((p.outW & 3) == 0);
  }
  __device__ __forceinline__ void store(const float (&v)[4][3], int n) {
    if (vec) {
      b200::stg_stream_f4(q0, make_float4(v[0][0], v[1][0], v[2][0], v[3][0]));
      b200::stg_stream_f4(q1, make_float4(v[0][1], v[1][1], v[2][1], v[3][1]));
      b200::stg_stream_f4(q2, make_float4(v[0][2], v[1][2], v[2][2], v[3][2]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n) { q0[k] = v[k][0]; q1[k] = v[k][1]; q2[k] = v[k][2]; }
    }
    q0 += row_stride; q1 += row_stride; q2 += row_stride;
  }
};
template <>
struct OutCursor<__half> {
  __half *q0, *q1, *q2;    // planes of source channels 0,1,2 (B,G,R): swapped to R,G,B order when swap_rb
  int64_t row_stride;
  bool vec;
  __device__ __forceinline__ void init(const LbParams& p, int b, int oy, int ox, int n) {
    __half* base = reinterpret_cast<__half*>(p.dst) + ((int64_t)b * 3 * p.outH + oy) * p.outW + ox;
    const int64_t plane = (int64_t)p.outH * p.outW;
    q0 = base + (p.swap_rb ? 2 : 0) * plane;
    q1 = base + plane;
    q2 = base + (p.swap_rb ? 0 : 2) * plane;
    row_stride = p.outW;
    vec = (n == 4) && ((reinterpret_cast<uintptr_t>(base) & 7) == 0) && ((plane & 3) == 0) && ((p.outW & 3) == 0);
  }
  __device__ __forceinline__ void store(const __half (&v)[4][3], int n) {
    if (vec) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        __half* q = c == 0 ? q0 : (c == 1 ? q1 : q2);
        const __half2 lo = __halves2half2(v[0][c], v[1][c]), hi = __halves2half2(v[2][c], v[3][c]);
        uint2 w;
        w.x = *reinterpret_cast<const uint32_t*>(&lo);
        w.y = *reinterpret_cast<const uint32_t*>(&hi);
        asm volatile("st.global.cs.v2.u32 [%0], {%1, %2};" ::"l"(q), "r"(w.x), "r"(w.y) : "memory");
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n) { q0[k] = v[k][0]; q1[k] = v[k][1]; q2[k] = v[k][2]; }
    }
    q0 += row_stride; q1 += row_stride; q2 += row_stride;
  }
};
template <>
struct OutCursor<uint8_t> {
  uint8_t* q;
  int64_t row_stride;
  bool vec;
  __device__ __forceinline__ void init(const LbParams& p, int b, int oy, int ox, int n) {
    q = reinterpret_cast<uint8_t*>(p.dst) + (((int64_t)b * p.outH + oy) * p.outW + ox) * 3;
    row_stride = (int64_t)p.outW * 3;
    vec = (n == 4) && ((reinterpret_cast<uintptr_t>(q) & 3) == 0) && ((row_stride & 3) == 0);
  }
  __device__ __forceinline__ void store(const uint8_t (&v)[4][3], int n) {
    if (vec) {
      uint32_t* qw = reinterpret_cast<uint32_t*>(q);
      qw[0] = v[0][0] | (v[0][1] << 8) | (v[0][2] << 16) | (v[1][0] << 24);
      qw[1] = v[1][1] | (v[1][2] << 8) | (v[2][0] << 16) | (v[2][1] << 24);
      qw[2] = v[2][2] | (v[3][0] << 8) | (v[3][1] << 16) | (v[3][2] << 24);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n) { q[k * 3 + 0] = v[k][0]; q[k * 3 + 1] = v[k][1]; q[k * 3 + 2] = v[k][2]; }
    }
    q += row_stride;
  }
};

template <typename OutT>
__global__ void __launch_bounds__(256, 4) letterbox_kernel(const LbParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[kStages];
  __shared__ float lut[256];
  __shared__ RowInfo rows[kRows];
  __shared__ int span[2];
  const int b = blockIdx.y, tid = threadIdx.x;
  const int oy0 = blockIdx.x * kRows;
  const int nrows = min(kRows, p.outH - oy0);
  const int ox_base = blockIdx.z * (int)blockDim.x * 4;          // this CTA's column chunk
  const int ox0 = ox_base + tid * 4;
  const bool have = ox0 < p.outW;
  const int npx = have ? min(4, p.outW - ox0) : 0;

  for (int v = tid; v < 256; v += blockDim.x) lut[v] = b200::u8_div255(v);
  if (tid < nrows) {            // vertical taps of this CTA's rows (shared by every thread)
    const int oy = oy0 + tid;
    RowInfo ri{0, 0, -1, 0};
    if (oy >= p.top && oy < p.top + p.new_h) cv_linear_tap_v(oy - p.top, p.scale_y, p.H, ri.r0, ri.r1, ri.b0, ri.b1);
    rows[tid] = ri;
  }
  if (tid == 32 || (blockDim.x <= 32 && tid == 0)) {
    // byte span of a source row this chunk's columns reference (16-byte granular for the bulk copy)
    const int lo_ox = max(ox_base, p.left);
    const int hi_ox = min(min(ox_base + (int)blockDim.x * 4, p.outW), p.left + p.new_w) - 1;
    int blo = 0, bhi = 0;
    if (lo_ox <= hi_ox) {
      int s_lo, s_hi, t0, t1;
      cv_linear_tap(lo_ox - p.left, p.scale_x, p.W, s_lo, t0, t1);
      cv_linear_tap(hi_ox - p.left, p.scale_x, p.W, s_hi, t0, t1);
      blo = (s_lo * 3) & ~15;
      bhi = p.bulk_ok ? min(p.row_bytes, ((min(s_hi + 2, p.W) * 3) + 15) & ~15) : min(s_hi + 2, p.W) * 3;
    }
    span[0] = blo; span[1] = bhi;
  }
  if (tid == 0 && p.bulk_ok) {
    for (int s = 0; s < kStages; ++s) b200::mbar_init(&bar[s], 1);
    b200::mbar_fence_init();
  }
  // horizontal taps of this thread's 4 pixels: derived once, kept in registers for all kRows rows
  int off[4], a0[4], a1[4];
  bool in[4], all_in = true, pass_h = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ox = ox0 + k;
    in[k] = have && (ox >= p.left) && (ox < p.left + p.new_w);
    off[k] = 0; a0[k] = 2048; a1[k] = 0;
    if (in[k]) cv_linear_tap(ox - p.left, p.scale_x, p.W, off[k], a0[k], a1[k]);
    all_in &= in[k];
    pass_h &= (a0[k] == 2048) && (a1[k] == 0);
  }
  const bool cta_pass_h = __syncthreads_and(pass_h);              // also publishes rows[], span[], lut[], barriers
  const int blo = span[0], nbytes = span[1] - span[0];
#pragma unroll
  for (int k = 0; k < 4; ++k) off[k] = in[k] ? off[k] * 3 - blo : 0;   // byte offset inside the staged span

  const int fi = b / p.ns, si = b - fi * p.ns;
  const uint8_t* frame = p.src + (int64_t)fi * p.bstride + (int64_t)p.sy[si] * p.pitch + (int64_t)p.sx[si] * 3 + blo;
  auto prefetch = [&](int r, int s) {        // stage the source rows of CTA row r into ring slot s
    const RowInfo ri = rows[r];
    if (ri.b0 < 0 || nbytes <= 0) return;
    const bool need1 = (ri.b1 != 0) && (ri.r1 != ri.r0);
    uint8_t* d0 = smem + (size_t)s * 2 * p.row_smem;
    uint8_t* d1 = d0 + p.row_smem;
    const uint8_t* g0 = frame + (int64_t)ri.r0 * p.pitch;
    const uint8_t* g1 = frame + (int64_t)ri.r1 * p.pitch;
    B200_CHECK(s >= 0 && s < kStages && nbytes <= p.row_smem);                 // the staged span fits its ring slot
    B200_CHECK(ri.r0 >= 0 && ri.r0 < p.H && ri.r1 >= 0 && ri.r1 < p.H);        // source rows of this frame / window
    if (p.bulk_ok) {
      B200_CHECK(((reinterpret_cast<uintptr_t>(g0) | reinterpret_cast<uintptr_t>(g1) | (uintptr_t)nbytes) & 15) == 0);
      if (tid == 0) {
        b200::mbar_expect_tx(&bar[s], need1 ? 2u * nbytes : (uint32_t)nbytes);
        b200::bulk_g2s(d0, g0, nbytes, &bar[s]);
        if (need1) b200::bulk_g2s(d1, g1, nbytes, &bar[s]);
      }
    } else {
      for (int i = tid; i < nbytes; i += blockDim.x) {
        d0[i] = g0[i];
        if (need1) d1[i] = g1[i];
      }
    }
  };

  OutCursor<OutT> cur;
  B200_CHECK(oy0 + nrows <= p.outH && (!have || ox0 + npx <= p.outW));         // this thread's output pixels exist
  if (have) cur.init(p, b, oy0, ox0, npx);
  const OutT padv = lb_cast<OutT>(lut, p.pad_value);
  unsigned uses = 0;                         // bit s = parity of completed phases of ring slot s (interior rows only)
#pragma unroll
  for (int r = 0; r < kStages - 1; ++r)
    if (r < nrows) prefetch(r, r);
  for (int r = 0; r < nrows; ++r) {
    const int s = r % kStages;
    const RowInfo ri = rows[r];
    // slot (r-1) % kStages was released by the barrier that ended row r-1: refill it with row r+kStages-1
#ifdef B200_CHECKS
    if (p.bulk_ok && r >= 1 && r + kStages - 1 < nrows) {
      // poison the released slot before its refill is issued: a blend that ran ahead of the refill's completion would
      // read 0xA5 bytes, and the parity tests compare bit for bit
      uint8_t* d = smem + (size_t)((r + kStages - 1) % kStages) * 2 * p.row_smem;
      for (int i = tid * 4; i < 2 * p.row_smem; i += (int)blockDim.x * 4) *reinterpret_cast<uint32_t*>(d + i) = 0xA5A5A5A5u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
#endif
    if (r + kStages - 1 < nrows) prefetch(r + kStages - 1, (r + kStages - 1) % kStages);
    OutT v[4][3];
    if (ri.b0 < 0 || nbytes <= 0) {                  // CTA-uniform: padding row / chunk fully in the side padding
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k][0] = v[k][1] = v[k][2] = padv;
      if (have) cur.store(v, npx);
      __syncthreads();
      continue;
    }
    if (p.bulk_ok) {
      b200::mbar_wait(&bar[s], (uses >> s) & 1u);
      uses ^= 1u << s;
    } else {
      __syncthreads();
    }
    const uint8_t* row0 = smem + (size_t)s * 2 * p.row_smem;
    if (have) {
      if (cta_pass_h && ri.b1 == 0 && ri.b0 == 2048) {
        // exact integer decimation (weights 2048/0 on both axes): the output IS a source pixel
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          B200_CHECK(off[k] >= 0 && (off[k] & ~3) + 8 <= p.row_smem);
          const uint32_t* w = reinterpret_cast<const uint32_t*>(row0 + (off[k] & ~3));
          const uint32_t lo = __funnelshift_r(w[0], w[1], (off[k] & 3) * 8);
          v[k][0] = lb_cast<OutT>(lut, (int)(lo & 0xff));
          v[k][1] = lb_cast<OutT>(lut, (int)((lo >> 8) & 0xff));
          v[k][2] = lb_cast<OutT>(lut, (int)((lo >> 16) & 0xff));
        }
      } else {
        const uint8_t* row1 = ((ri.b1 != 0) && (ri.r1 != ri.r0)) ? row0 + p.row_smem : row0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t lo, hi;
          B200_CHECK(off[k] >= 0 && ((off[k] >> 2) + 3) * 4 <= p.row_smem);      // load6 reads three words
          load6(row0, off[k], lo, hi);
          int S0[3], S1[3];
          S0[0] = (int)(lo & 0xff) * a0[k] + (int)(lo >> 24) * a1[k];
          S0[1] = (int)((lo >> 8) & 0xff) * a0[k] + (int)(hi & 0xff) * a1[k];
          S0[2] = (int)((lo >> 16) & 0xff) * a0[k] + (int)((hi >> 8) & 0xff) * a1[k];
          load6(row1, off[k], lo, hi);
          S1[0] = (int)(lo & 0xff) * a0[k] + (int)(lo >> 24) * a1[k];
          S1[1] = (int)((lo >> 8) & 0xff) * a0[k] + (int)(hi & 0xff) * a1[k];
          S1[2] = (int)((lo >> 16) & 0xff) * a0[k] + (int)((hi >> 8) & 0xff) * a1[k];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int q = (((ri.b0 * (S0[c] >> 4)) >> 16) + ((ri.b1 * (S1[c] >> 4)) >> 16) + 2) >> 2;
            v[k][c] = lb_cast<OutT>(lut, q);
          }
        }
      }
      if (!all_in) {                                  // side padding columns (edge threads only)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!in[k]) v[k][0] = v[k][1] = v[k][2] = padv;
      }
      cur.store(v, npx);
    }
    __syncthreads();   // every thread is done with ring slot s: it may be refilled for row r+2
  }
}

template <typename OutT>
int launch_letterbox(const uint8_t* src, int B, int H, int W, int64_t pitch, int64_t bstride, void* dst,
                     int outH, int outW, int new_w, int new_h, int top, int left, int pad_value, int swap_rb,
                     void* stream, const int* slice_xy = nullptr, int n_slices = 1) {
  B200_REQUIRE(src && dst, B200YOLO_ERR_NULL);
  B200_REQUIRE(n_slices >= 1 && n_slices <= kMaxSlices && B % n_slices == 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && outH > 0 && outW > 0 && new_w > 0 && new_h > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(top >= 0 && left >= 0 && top + new_h <= outH && left + new_w <= outW, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(pitch >= (int64_t)W * 3 && (B == 1 || bstride >= pitch * (int64_t)(H - 1) + (int64_t)W * 3), B200YOLO_ERR_SHAPE);
  B200_REQUIRE(pad_value >= 0 && pad_value <= 255, B200YOLO_ERR_RANGE);
  B200_REQUIRE(B <= 65535 && outH <= 2147483647, B200YOLO_ERR_UNSUPPORTED);
  LbParams p;
  p.src = src; p.dst = dst; p.pitch = pitch; p.bstride = bstride;
  p.H = H; p.W = W; p.outH = outH; p.outW = outW; p.new_w = new_w; p.new_h = new_h;
  p.top = top; p.left = left; p.pad_value = pad_value; p.swap_rb = swap_rb;
  // cv::resize: inv_scale = dsize/ssize (double); scale = 1/inv_scale
  p.scale_x = 1.0 / ((double)new_w / (double)W);
  p.scale_y = 1.0 / ((double)new_h / (double)H);
  p.row_bytes = W * 3;
  p.row_smem = ((p.row_bytes + 15) / 16) * 16 + 32;
  p.bulk_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (pitch % 16 == 0) && (bstride % 16 == 0) &&
              (p.row_bytes % 16 == 0);
  p.ns = n_slices;
  for (int i = 0; i < kMaxSlices; ++i) {
    p.sx[i] = (slice_xy && i < n_slices) ? slice_xy[2 * i] : 0;
    p.sy[i] = (slice_xy && i < n_slices) ? slice_xy[2 * i + 1] : 0;
    if (p.sx[i] % 16 != 0) p.bulk_ok = 0;          // sx * 3 bytes must keep the rows 16-byte aligned
  }
  const size_t smem = (size_t)kStages * 2 * (size_t)p.row_smem;
  B200_REQUIRE(smem <= 200 * 1024, B200YOLO_ERR_UNSUPPORTED);
  auto kern = letterbox_kernel<OutT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int groups = (outW + 3) / 4;
  int chunks = (groups + 255) / 256;                 // column chunks (grid.z): one pixel group per thread
  int threads = (((groups + chunks - 1) / chunks + 31) / 32) * 32;
  dim3 grid((outH + kRows - 1) / kRows, B, chunks);
  kern<<<grid, threads, smem, (cudaStream_t)stream>>>(p);
  return b200_launch_status();
}

}  // namespace

extern "C" int b200yolo_letterbox_u8_to_f32(const uint8_t* src, int B, int H, int W, int64_t src_pitch,
                                            int64_t src_batch_stride, float* dst, int outH, int outW,
                                            int new_w, int new_h, int top, int left, int pad_value,
                                            int swap_rb, void* stream) {
  return launch_letterbox<float>(src, B, H, W, src_pitch, src_batch_stride, dst, outH, outW, new_w, new_h, top,
                                 left, pad_value, swap_rb, stream);
}

extern "C" int b200yolo_letterbox_u8_to_f16(const uint8_t* src, int B, int H, int W, int64_t src_pitch,
                                            int64_t src_batch_stride, void* dst, int outH, int outW, int new_w,
                                            int new_h, int top, int left, int pad_value, int swap_rb, void* stream) {
  return launch_letterbox<__half>(src, B, H, W, src_pitch, src_batch_stride, dst, outH, outW, new_w, new_h, top, left,
                                  pad_value, swap_rb, stream);
}

extern "C" int b200yolo_letterbox_slices_u8_to_f32(const uint8_t* frames, int n_frames, int frame_h, int frame_w,
                                                   int64_t pitch, int64_t frame_stride, const int* slice_xy,
                                                   int n_slices, int slice_h, int slice_w, float* dst, int outH,
                                                   int outW, int new_w, int new_h, int top, int left, int pad_value,
                                                   int swap_rb, void* stream) {
  B200_REQUIRE(frames && slice_xy && dst, B200YOLO_ERR_NULL);
  B200_REQUIRE(n_frames > 0 && n_slices > 0 && slice_h > 0 && slice_w > 0 && frame_h >= slice_h && frame_w >= slice_w,
               B200YOLO_ERR_SHAPE);
  B200_REQUIRE(n_slices <= kMaxSlices, B200YOLO_ERR_UNSUPPORTED);
  B200_REQUIRE(pitch >= (int64_t)frame_w * 3 &&
                   (n_frames == 1 || frame_stride >= pitch * (int64_t)(frame_h - 1) + (int64_t)frame_w * 3),
               B200YOLO_ERR_SHAPE);
  for (int i = 0; i < n_slices; ++i)
    B200_REQUIRE(slice_xy[2 * i] >= 0 && slice_xy[2 * i + 1] >= 0 && slice_xy[2 * i] + slice_w <= frame_w &&
                     slice_xy[2 * i + 1] + slice_h <= frame_h,
                 B200YOLO_ERR_SHAPE);
  // the per-item shape check of launch_letterbox (pitch vs slice width, stride vs slice rows) is implied by the above
  return launch_letterbox<float>(frames, n_frames * n_slices, slice_h, slice_w, pitch, frame_stride, dst, outH, outW,
                                 new_w, new_h, top, left, pad_value, swap_rb, stream, slice_xy, n_slices);
}

extern "C" int b200yolo_letterbox_u8(const uint8_t* src, int B, int H, int W, int64_t src_pitch,
                                     int64_t src_batch_stride, uint8_t* dst, int outH, int outW, int new_w,
                                     int new_h, int top, int left, int pad_value, void* stream) {
  return launch_letterbox<uint8_t>(src, B, H, W, src_pitch, src_batch_stride, dst, outH, outW, new_w, new_h, top,
                                   left, pad_value, 0, stream);
}
