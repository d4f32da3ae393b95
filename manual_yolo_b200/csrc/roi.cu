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
// One CTA per ROI.  Two bodies: the fast path (192 threads, every crop up to resample scale 4: see roi_fast_body) and
// the general body (384 threads = 3 channels x 64 columns x 2 row phases, any scale <= 31, run by roi_big_kernel for
// the ROIs the fast path defers).  Both build the two coefficient tables for the 64 output columns/rows that survive
// the centre crop only, write Pillow's uint8 intermediate strip into shared memory (horizontal pass) and stream
// coalesced fp32 rows to the planar (3,size,size) output (vertical pass).
// Algorithmic bytes per ROI = crop_h*crop_w*3 read + 3*size*size*4 written; the arithmetic (one 8 x 22-bit
// integer product per tap) keeps the kernel issue-bound rather than HBM-bound: see DESIGN.md.

#include "common.cuh"

namespace {

constexpr int kS = 64;           // output side (the rank classifier's imgsz; class.py:26, args.yaml imgsz: 64)
constexpr int kRowPar = 2;       // row phases: threads sharing a (channel, column) interleave over rows
constexpr int kCols = 3 * kS;    // (channel, output column) pairs
constexpr int kThreads = kCols * kRowPar;
constexpr int kMaxTaps = 64;     // ksize = ceil(support)*2+1 <= 64  <=> scale <= 31
constexpr int kFastTaps = 8;     // taps kept in registers (scale <= 3.5: every rank-card crop)
constexpr int kRowsMax = 128;    // uint8 strip rows staged per vertical tile (>= kMaxTaps)
constexpr int kPrec = 22;        // Pillow PRECISION_BITS = 32 - 8 - 2
constexpr int kBigParts = 8;     // CTAs per large ROI: each produces kS/kBigParts output rows
constexpr int kBigCtas = 32;     // large-ROI launch width (grid-strides over the deferred list)
constexpr int kStageBytes = 48 * 1024;  // staged crop rows (a 115x105 rank crop needs ~37 KB)

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output index `o` (triangle filter):
// returns xmin, writes `cnt` fixed-point weights.
template <int MAXT>
__device__ int pil_axis(int o, int in_size, int out_size, int* coef, int cstride, int& cnt_out) {
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
  if (xmax > MAXT) xmax = MAXT;  // unreachable inside the supported envelope
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
    coef[x * cstride] = __double2int_rz(__dadd_rn(0.5, __dmul_rn(w, (double)(1 << kPrec))));
  }
  cnt_out = xmax;
  return xmin;
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

struct RoiSmem {
  int xb[kS][2];                 // horizontal bounds (xmin, count) per surviving output column
  int yb[kS][2];                 // vertical bounds per surviving output row
  int xkT[kMaxTaps][kS];         // horizontal weights, transposed: xkT[tap][column] (conflict-free per warp)
  int yk[kS][kMaxTaps];          // vertical weights (broadcast reads)
  uint8_t strip[kRowsMax][3][kS];  // horizontal-pass output, uint8 like Pillow's intermediate image
  int sel[4];
  __align__(16) uint8_t stage[kStageBytes];   // referenced crop rows, staged once with coalesced loads
};

// Word load that never touches bytes outside [lo, hi) (the caller's frame buffer).
__device__ __forceinline__ uint32_t load_word_guarded(const uint8_t* a, const uint8_t* lo, const uint8_t* hi) {
  if (a >= lo && a + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(a));
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (a + k >= lo && a + k < hi) w |= (uint32_t)__ldg(a + k) << (8 * k);
  return w;
}

// One ROI per CTA, one thread per (channel c, output column xx).  (bi, x1..y2): already int()-truncated
// box in source pixels.
__device__ void roi_body(const uint8_t* __restrict__ frames, const uint8_t* buf_hi, int B, int H, int W, int64_t pitch,
                         int64_t bstride,
                         int bi, int bx1, int by1, int bx2, int by2, int pad, float* __restrict__ out,
                         int* __restrict__ valid_out, RoiSmem& sm, int part, int nparts) {
  const int tid = threadIdx.x;
  const int y_begin = part * (kS / nparts), y_end = y_begin + kS / nparts;   // this CTA's output rows
  const int rp = tid / kCols, tcol = tid % kCols;
  const int c = tcol / kS, xx = tcol % kS;     // warp = 32 consecutive columns of one channel, one row phase
  // ---- safe_crop (detect.py:100-113) ----
  const int cx1 = max(0, min(W - 1, bx1 - pad)), cx2 = max(0, min(W, bx2 + pad));
  const int cy1 = max(0, min(H - 1, by1 - pad)), cy2 = max(0, min(H, by2 + pad));
  const int cw = cx2 - cx1, ch = cy2 - cy1;
  bool ok = (bi >= 0 && bi < B && cw > 0 && ch > 0), unsupported = false;
  // torchvision Resize(int): short side -> S, long side -> int(S * long / short)
  int new_w = kS, new_h = kS;
  if (ok) {
    if (cw <= ch) new_h = __double2int_rz(__ddiv_rn((double)(kS * (int64_t)ch), (double)cw));
    else new_w = __double2int_rz(__ddiv_rn((double)(kS * (int64_t)cw), (double)ch));
    // supported envelope: ksize = ceil(support)*2+1 <= kMaxTaps on both axes
    const double sx = (double)cw / (double)new_w, sy = (double)ch / (double)new_h;
    const double smx = fmax(fmax(sx, sy), 1.0);
    if (2 * (int)ceil(smx) + 1 > kMaxTaps) { ok = false; unsupported = true; }
  }
  if (!ok) {
    for (int c2 = 0; c2 < 3; ++c2)
      for (int i = tid; i < (y_end - y_begin) * kS; i += kThreads) out[(c2 * kS + y_begin) * kS + i] = 0.f;
    if (tid == 0 && part == 0) *valid_out = unsupported ? -1 : 0;
    return;
  }
  const int left = half_round_even(new_w - kS), top = half_round_even(new_h - kS);

  // ---- coefficient tables: threads 0..63 horizontal (own column), 64.. vertical (this CTA's rows only) ----
  if (tid < kS) {
    int cnt;
    const int xm = pil_axis<kMaxTaps>(left + tid, cw, new_w, &sm.xkT[0][tid], kS, cnt);
    sm.xb[tid][0] = xm; sm.xb[tid][1] = cnt;
  } else if (tid < kS + (y_end - y_begin)) {
    const int yy = y_begin + (tid - kS);
    int cnt;
    const int ymin = pil_axis<kMaxTaps>(top + yy, ch, new_h, sm.yk[yy], 1, cnt);
    sm.yb[yy][0] = ymin; sm.yb[yy][1] = cnt;
  }
  __syncthreads();
  const int xmin = sm.xb[xx][0], xcnt = sm.xb[xx][1];
  const bool fast_x = xcnt <= kFastTaps;
  int kx[kFastTaps];
#pragma unroll
  for (int x = 0; x < kFastTaps; ++x) kx[x] = (x < xcnt) ? sm.xkT[x][xx] : 0;

  const uint8_t* crop = frames + (int64_t)bi * bstride + (int64_t)cy1 * pitch + (int64_t)cx1 * 3;
  // horizontal span of source columns the 64 surviving output columns reference
  const int x_lo = sm.xb[0][0];
  const int span_bytes = (sm.xb[kS - 1][0] + sm.xb[kS - 1][1] - x_lo) * 3;
  const int row_words = (span_bytes + 3 + 3) / 4 + 3;        // any 4-byte phase fits, + slack for the 4-tap reads
  const int row_stride = row_words * 4;
  const int stage_rows = kStageBytes / row_stride;            // rows of this ROI one stage fill can hold (>= 1)
  const int cofs = (xmin - x_lo) * 3 + c;
  int t0 = y_begin;
  while (t0 < y_end) {
    // vertical tile [t0, t1): the input rows [rmin, rmin+rows) it references must fit the uint8 strip
    const int rmin = sm.yb[t0][0];
    int t1 = t0 + 1;
    while (t1 < y_end && sm.yb[t1][0] + sm.yb[t1][1] - rmin <= kRowsMax) ++t1;
    const int rows = sm.yb[t1 - 1][0] + sm.yb[t1 - 1][1] - rmin;
    // ---- horizontal pass, stage-sized row chunks: coalesced 32-bit loads (all in flight at once) into the
    //      stage, then every thread resamples its column for the chunk's rows ----
    for (int r0 = 0; r0 < rows; r0 += stage_rows) {
      const int nr = min(stage_rows, rows - r0);
      if (r0 > 0) __syncthreads();                            // previous chunk fully consumed
      {   // warp per row, lanes stride over the row's words (no integer division per element)
        const int lane = tid & 31, wid = tid >> 5;
        for (int rr = wid; rr < nr; rr += kThreads / 32) {
          const uint8_t* g = crop + (int64_t)(rmin + r0 + rr) * pitch + x_lo * 3;
          const uint8_t* ga = g - (reinterpret_cast<uintptr_t>(g) & 3);
          uint32_t* srow = reinterpret_cast<uint32_t*>(sm.stage + rr * row_stride);
          for (int wd = lane; wd < row_words; wd += 32) srow[wd] = load_word_guarded(ga + wd * 4, frames, buf_hi);
        }
      }
      __syncthreads();
      if (xcnt <= 4) {
        // <= 4 taps: every up-scale and down-scales to 1.5x (the rank-card case)
        for (int rr = rp; rr < nr; rr += kRowPar) {
          const uint8_t* g = crop + (int64_t)(rmin + r0 + rr) * pitch + x_lo * 3;
          const uint8_t* p = sm.stage + rr * row_stride + (reinterpret_cast<uintptr_t>(g) & 3) + cofs;
          // taps beyond xcnt have weight 0; their bytes are inside the staged row (+ slack words)
          const int acc = (1 << (kPrec - 1)) + (int)p[0] * kx[0] + (int)p[3] * kx[1] + (int)p[6] * kx[2] + (int)p[9] * kx[3];
          sm.strip[r0 + rr][c][xx] = (uint8_t)clip8(acc);
        }
      } else if (fast_x) {
        for (int rr = rp; rr < nr; rr += kRowPar) {
          const uint8_t* g = crop + (int64_t)(rmin + r0 + rr) * pitch + x_lo * 3;
          const uint8_t* p = sm.stage + rr * row_stride + (reinterpret_cast<uintptr_t>(g) & 3) + cofs;
          int acc = 1 << (kPrec - 1);
#pragma unroll
          for (int x = 0; x < kFastTaps; ++x)
            if (x < xcnt) acc += (int)p[x * 3] * kx[x];
          sm.strip[r0 + rr][c][xx] = (uint8_t)clip8(acc);
        }
      } else {
        for (int rr = rp; rr < nr; rr += kRowPar) {
          const uint8_t* g = crop + (int64_t)(rmin + r0 + rr) * pitch + x_lo * 3;
          const uint8_t* p = sm.stage + rr * row_stride + (reinterpret_cast<uintptr_t>(g) & 3) + cofs;
          int acc = 1 << (kPrec - 1);
#pragma unroll 4
          for (int x = 0; x < xcnt; ++x) acc += (int)p[x * 3] * sm.xkT[x][xx];
          sm.strip[r0 + rr][c][xx] = (uint8_t)clip8(acc);
        }
      }
    }
    __syncthreads();
    // ---- vertical pass + BGR->RGB + /255: thread = (channel, 4 adjacent columns), 8 row phases; one
    //      32-bit LDS feeds 4 accumulators per tap, one 128-bit store per 4 outputs ----
    {
      const int vq = tid % 48, vph = tid / 48;           // 48 column quads (3 channels x 16), 8 row phases
      const int vc = vq >> 4, vx = (vq & 15) * 4;
      for (int yy = t0 + vph; yy < t1; yy += kThreads / 48) {
        const int ymin = sm.yb[yy][0] - rmin, cnt = sm.yb[yy][1];
        int a0 = 1 << (kPrec - 1), a1 = a0, a2 = a0, a3 = a0;
        const uint8_t* sp = &sm.strip[ymin][vc][vx];
        const int* kp = sm.yk[yy];
#pragma unroll 4
        for (int y = 0; y < cnt; ++y) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(sp + y * (3 * kS));
          const int k = kp[y];
          a0 += (int)(w & 0xff) * k;
          a1 += (int)((w >> 8) & 0xff) * k;
          a2 += (int)((w >> 16) & 0xff) * k;
          a3 += (int)(w >> 24) * k;
        }
        float4 o = make_float4(b200::u8_div255(clip8(a0)), b200::u8_div255(clip8(a1)), b200::u8_div255(clip8(a2)),
                               b200::u8_div255(clip8(a3)));
        *reinterpret_cast<float4*>(out + ((2 - vc) * kS + yy) * kS + vx) = o;
      }
    }
    __syncthreads();
    t0 = t1;
  }
  if (tid == 0 && part == 0 && nparts == 1) *valid_out = 1;
}

// ---------------------------------------------------------------------------------------------------------
// Fast path (first launch): every ROI with <= kFTaps taps per axis (resample scale <= 4, i.e. short side <= 256)
// -- every rank card and everything near it.  One CTA of 192 threads per ROI, ~42 KB of shared memory, 5 CTAs
// (30 warps) per SM.
//
//  * staging: each warp streams its own source rows (w, w+6, ..) through a private ring of 2-4 row slots filled by
//    the TMA engine -- lane 0 issues one 1-D bulk copy per row (cp.async.bulk global->shared completing on the
//    slot's mbarrier; SASS UBLKCP), so staging costs a handful of instructions per row instead of per-lane
//    address arithmetic, and rows r+6 .. r+24 are in flight while row r is resampled.  No CTA barrier inside the
//    horizontal pass (a CTA-wide chunked stage with one barrier per chunk was measured: 183 us vs 143 us).
//  * horizontal pass: warp w resamples rows w, w+6, .. of the tile; lane owns output columns lane and lane+32
//    (3 channels each: conflict-free byte reads), weights in registers, taps unrolled to the crop's exact tap
//    class (2/3/4/5/6/8/10: up-scales have 2).  The weights are stored pre-shifted by 2 bits, so the rounded
//    uint8 result is the TOP BYTE of the 32-bit accumulator.
//  * vertical pass: thread = (channel, 4 adjacent columns) of one output row, one 32-bit LDS feeds 4 accumulators
//    per tap, taps unrolled to the tap class; v/255 exactly as torch divides, computed in the FMA pipe (no table:
//    no bank conflicts); one 128-bit streaming store per 4 outputs.
// Crops taller than the uint8 strip are produced in vertical tiles of output rows.  ROIs outside the envelope are
// marked valid = 2 and produced by roi_big_kernel (the general body above, split over kBigParts CTAs).
constexpr int kFT = 192;          // threads: 6 warps
constexpr int kFWarps = kFT / 32;
constexpr int kFTaps = 10;        // taps per axis (ksize = 2*ceil(scale)+1 <= 9 for scale <= 4)
constexpr int kFKy = 12;          // vertical weight row, padded to int4 multiples
constexpr int kFRows = 106;       // uint8 strip rows: source rows one vertical tile may reference (tallest rank crop: 105)
constexpr int kPool = 1248;       // per-warp row-slot pool: 4 rows of <= 304 B, 3 of <= 416 B, 2 of <= 624 B, 1 of <= 1248 B
constexpr int kRingMax = 4;       // rows in flight per warp
constexpr uint32_t kRnd4 = 1u << (kPrec + 1);   // Pillow's 2^(PRECISION_BITS-1) rounding term, in the <<2 domain
#ifndef B200_K5_SEG_SCALE
#define B200_K5_SEG_SCALE 1.3
#endif
#ifndef B200_K5_SEG_MIN
#define B200_K5_SEG_MIN 1024
#endif
constexpr double kSegScale = B200_K5_SEG_SCALE; // two-segment launches: crops down-scaled by more than this run first
constexpr int kSegMinRois = B200_K5_SEG_MIN;    // ... when the launch holds at least this many ROIs (more than one wave of CTAs)

struct FastSmem {
  uint32_t xk[kFTaps][kS];        // horizontal weights << 2, transposed (conflict-free per warp)
  uint32_t yk[kS][kFKy];          // vertical weights << 2 (broadcast reads)
  int xb[kS][2];                  // (xmin, count) per surviving output column
  int yb[kS][2];                  // per surviving output row
  uint8_t strip[kFRows + kFTaps][3][kS];  // horizontal-pass output (Pillow's uint8 intermediate image) + rows that
                                          // only zero-weight taps of the unrolled vertical pass may touch
  __align__(16) uint8_t ring[kFWarps][kPool];
  __align__(16) uint8_t slack[48];  // zero-weight taps of the last staged row may read (never use) bytes past it
  __align__(8) uint64_t bar[kFWarps][kRingMax];
  int sel[4];
};

// Pillow's clip8 clamps (acc >> 22) to [0,255].  For the bilinear (triangle) filter every weight is >= 0 and the
// fixed-point weights of one output sum to at most 2^22 + taps/2, so 0 <= acc = 2^21 + sum(p_i * k_i)
// <= 2^21 + 255 * (2^22 + 5) < 256 * 2^22: the clamp can never act and the fast path omits it (the general body
// keeps it; both are compared bit-for-bit against PIL in tests/test_gpu_roi.py).  Hence 4 * acc < 2^32: with the
// weights and the rounding term pre-shifted by two bits the unsigned accumulator holds acc >> 22 in its top byte.

// (a >> 24) / 255 exactly as torch's fp32 division rounds it, without a table: v = a >> 24 is built as the float
// 2^23 + v by one byte permute, v/255 = RN(v * c_hi + RN(v * c_lo)) with c_hi + c_lo = 1/255 to 48 bits -- equal to
// IEEE v / 255.0f for every v in [0, 255] (checked exhaustively on the host in tests/test_host_logic.py and, through
// every K5 output, against PIL on the device).
// Two accumulators at a time on Blackwell's packed fp32 pipe (SASS FADD2 / FMUL2 / FFMA2).
__device__ __forceinline__ float2 top_byte_div255_x2(uint32_t a, uint32_t b) {
  float2 d;
  asm("{\n\t"
      ".reg .b64 x, v, t, m, cl, ch;\n\t"
      "mov.b64 x, {%2, %3};\n\t"
      "mov.b64 m, {%4, %4};\n\t"
      "mov.b64 cl, {%5, %5};\n\t"
      "mov.b64 ch, {%6, %6};\n\t"
      "add.rn.f32x2 v, x, m;\n\t"
      "mul.rn.f32x2 t, v, cl;\n\t"
      "fma.rn.f32x2 t, v, ch, t;\n\t"
      "mov.b64 {%0, %1}, t;\n\t"
      "}"
      : "=f"(d.x), "=f"(d.y)
      : "r"(__byte_perm(a, 0x4B000000u, 0x7443)), "r"(__byte_perm(b, 0x4B000000u, 0x7443)), "f"(-8388608.0f),
        "f"(__uint_as_float(0xAF7F00BFu)), "f"(__uint_as_float(0x3B808081u)));
  return d;
}

// x is the same in every lane: returns it in a form the compiler knows to be warp-uniform (REDUX writes a uniform
// register), so the address arithmetic that feeds the bulk copies stays in the uniform datapath.
__device__ __forceinline__ uint32_t uni(uint32_t x) { return __reduce_max_sync(0xffffffffu, x); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for output index `o` (triangle filter), fast-path form: the
// (<= kFTaps) weights go straight to shared memory, pre-shifted, zero-filled up to `pad` entries.  Same double
// arithmetic, in the same order, as pil_axis above.
__device__ __forceinline__ int pil_axis_fast(int o, int in_size, int out_size, uint32_t* coef, int cstride, int pad,
                                             int& cnt_out) {
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double ss = __ddiv_rn(1.0, fscale);
  const double center = __dmul_rn((double)o + 0.5, scale);
  int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, fscale), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, fscale), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  if (xmax > kFTaps) xmax = kFTaps;  // unreachable inside the fast envelope (scale <= 4 -> <= 9 taps)
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
    coef[x * cstride] = (uint32_t)__double2int_rz(__dadd_rn(0.5, __dmul_rn(w, (double)(1 << kPrec)))) << 2;
  }
  for (int x = xmax; x < pad; ++x) coef[x * cstride] = 0u;
  cnt_out = xmax;
  return xmin;
}

// Everything the two passes of one vertical tile share (CTA-uniform).
struct FastTile {
  const uint8_t* row0;      // global address of the first referenced byte of the tile's first source row
  const uint8_t* frames;    // caller's buffer bounds (guarded path)
  const uint8_t* buf_hi;
  int64_t pitch;
  int rows;                 // source rows the tile references
  int span;                 // referenced bytes per row
  int row_stride;           // staged row pitch (worst 16-byte phase + span, multiple of 16)
  int nslots;               // ring slots per warp (2..kRingMax)
  bool inside;              // every 16-byte chunk of every row lies inside the caller's buffer
};

// One source row -> 6 strip bytes per lane (columns lane and lane+32, 3 channels), T taps unrolled; taps beyond a
// column's count have weight 0 (their bytes are whatever follows in the ring: read, never used).
template <int T>
__device__ __forceinline__ void hrow(const uint8_t* __restrict__ p, const uint8_t* __restrict__ q,
                                     const uint32_t (&k0)[T], const uint32_t (&k1)[T], uint8_t* __restrict__ so) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint32_t a = kRnd4, b = kRnd4;
#pragma unroll
    for (int t = 0; t < T; ++t) { a += (uint32_t)p[3 * t + c] * k0[t]; b += (uint32_t)q[3 * t + c] * k1[t]; }
    so[c * kS] = (uint8_t)(a >> 24);
    so[c * kS + 32] = (uint8_t)(b >> 24);
  }
}

// Horizontal pass of one vertical tile: warp w owns source rows w, w+6, ..  (sidx, parity) is the warp's ring cursor:
// slots are used round-robin, the cursor keeps running across tiles, and every slot completes one mbarrier phase per
// round, so one parity bit (flipped when the cursor wraps) serves all slots.
template <int T>
__device__ __forceinline__ void hpass(FastSmem& sm, const FastTile& t, int xo0, int xo1, uint32_t& sidx,
                                      uint32_t& parity, int tid) {
  const int lane = tid & 31;
  const uint32_t wid = uni((uint32_t)tid >> 5);
  uint32_t k0[T], k1[T];
#pragma unroll
  for (int i = 0; i < T; ++i) { k0[i] = sm.xk[i][lane]; k1[i] = sm.xk[i][lane + 32]; }
  // warp-uniform copies of the tile parameters (they come from shared / global loads: the compiler cannot tell)
  const uint32_t rows = uni((uint32_t)t.rows), span = uni((uint32_t)t.span), rstride = uni((uint32_t)t.row_stride);
  const uint32_t nslots = uni((uint32_t)t.nslots);
  const uint64_t row0 = ((uint64_t)uni((uint32_t)(reinterpret_cast<uintptr_t>(t.row0) >> 32)) << 32) |
                        uni((uint32_t)reinterpret_cast<uintptr_t>(t.row0));
  uint8_t* const ring = &sm.ring[0][0] + wid * kPool;
  const uint32_t ring_a = b200::smem_u32(&sm.ring[0][0]) + wid * kPool;
  const uint32_t bar_a = b200::smem_u32(&sm.bar[0][0]) + wid * (8 * kRingMax);
  const uint32_t pstep = (uint32_t)(t.pitch & 15), dph = (kFWarps * pstep) & 15u;
  uint32_t ph = ((uint32_t)(row0 & 15) + wid * pstep) & 15u;   // 16-byte phase of row j
  uint8_t* so = &sm.strip[0][0][lane] + wid * (3 * kS);
  const uint64_t step = (uint64_t)kFWarps * (uint64_t)t.pitch;
  if (uni(t.inside ? 1u : 0u)) {
    uint64_t gi = row0 + (uint64_t)wid * (uint64_t)t.pitch;    // next row to issue, its index and phase
    uint32_t ji = wid, phi = ph;
    auto issue = [&](uint32_t s) {
      if (ji < rows) {
        if (elect_one()) {
          const uint32_t nb = (phi + span + 15u) & ~15u;
          B200_CHECK(s < nslots && s * rstride + nb <= (uint32_t)kPool);                 // the copy stays inside the slot pool
          B200_CHECK(reinterpret_cast<const uint8_t*>(gi - phi) >= t.frames &&
                     reinterpret_cast<const uint8_t*>(gi - phi) + nb <= t.buf_hi);       // ... and inside the caller's buffer
          B200_CHECK(((gi - phi) & 15u) == 0 && ((ring_a + s * rstride) & 15u) == 0);    // bulk-copy alignment
          b200::mbar_expect_tx_addr(bar_a + 8u * s, nb);
          b200::bulk_g2s_addr(ring_a + s * rstride, reinterpret_cast<const void*>(gi - phi), nb, bar_a + 8u * s);
        }
      }
      gi += step; ji += kFWarps; phi = (phi + dph) & 15u;
    };
    {
      uint32_t s = sidx;
      for (uint32_t r = 0; r < nslots; ++r) { issue(s); s = s + 1u == nslots ? 0u : s + 1u; }
    }
    for (uint32_t j = wid; j < rows; j += kFWarps) {
      b200::mbar_wait_addr(bar_a + 8u * sidx, parity);
      const uint8_t* rb = ring + sidx * rstride + ph;
      // the widest tap window read (zero-weight taps included) ends inside the ring array + its slack
      B200_CHECK(rb + max(xo0, xo1) + 3 * (T - 1) + 2 < &sm.ring[0][0] + kFWarps * kPool + (int)sizeof(sm.slack));
      B200_CHECK(so + 2 * kS + 32 < &sm.strip[0][0][0] + sizeof(sm.strip) && so >= &sm.strip[0][0][0]);
      hrow<T>(rb + xo0, rb + xo1, k0, k1, so);
      __syncwarp();                                          // every lane is done with this slot before it is refilled
#ifdef B200_CHECKS
      // poison the consumed slot before its refill is issued: a read that ran ahead of the refill's completion (a
      // wrong parity, a missing wait) would resample 0xA5 bytes, and the parity tests compare bit for bit
      for (uint32_t o = lane * 4; o < rstride; o += 128) *reinterpret_cast<uint32_t*>(ring + sidx * rstride + o) = 0xA5A5A5A5u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
#endif
      issue(sidx);
      so += kFWarps * 3 * kS;
      ph = (ph + dph) & 15u;
      if (++sidx == nslots) { sidx = 0; parity ^= 1u; }
    }
  } else {
    // crop touching the first/last bytes of the caller's allocation: guarded word loads into slot 0, no bulk copy
    // (which could start before / end after the buffer), no barrier
    const uint8_t* g = t.row0 + (int64_t)wid * t.pitch;
    for (uint32_t j = wid; j < rows; j += kFWarps, g += step, so += kFWarps * 3 * kS, ph = (ph + dph) & 15u) {
      const int nchunk16 = (int)((ph + span + 15u) >> 4);
      for (int ck = lane; ck < nchunk16; ck += 32) {
        uint32_t* dw = reinterpret_cast<uint32_t*>(ring + ck * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u) dw[u] = load_word_guarded(g - ph + ck * 16 + 4 * u, t.frames, t.buf_hi);
      }
      __syncwarp();
      B200_CHECK(ring + ph + max(xo0, xo1) + 3 * (T - 1) + 2 < &sm.ring[0][0] + kFWarps * kPool + (int)sizeof(sm.slack));
      B200_CHECK(so + 2 * kS + 32 < &sm.strip[0][0][0] + sizeof(sm.strip) && so >= &sm.strip[0][0][0]);
      hrow<T>(ring + ph + xo0, ring + ph + xo1, k0, k1, so);
      __syncwarp();
    }
  }
}

// Vertical pass + BGR->RGB + /255 for output rows [t0, t1).  Per round of 4 rows: warps {0,1} / {3,4} take channels
// 0-1 of two rows (lane -> channel = lane/16, column quad = lane%16), warps 2 / 5 take channel 2 of both -- every
// LDS.32 of a warp is bank-conflict free (rows of one warp are identical or 48 words apart).
template <int T>
__device__ __forceinline__ void vpass(FastSmem& sm, int t0, int t1, int rmin, float* __restrict__ out, int tid) {
  const int lane = tid & 31, wid = tid >> 5;
  const int w3 = wid % 3, half = lane >> 4;
  const int vc = w3 == 2 ? 2 : half;
  const int rsel = 2 * (wid / 3) + (w3 == 2 ? half : w3);
  const int vx = (lane & 15) * 4;
  float* o = out + ((2 - vc) * kS + t0 + rsel) * kS + vx;
  const uint8_t* spb = &sm.strip[0][vc][vx] - rmin * (3 * kS);
  const int* ybp = &sm.yb[t0 + rsel][0];
  const uint32_t* kq = &sm.yk[t0 + rsel][0];
  for (int yy = t0 + rsel; yy < t1; yy += 4, o += 4 * kS, ybp += 4 * 2, kq += 4 * kFKy) {
    const uint8_t* sp = spb + *ybp * (3 * kS);
    // every tap row read (zero-weight ones included) lies inside the strip; the 4 outputs inside the ROI's plane
    B200_CHECK(sp >= &sm.strip[0][0][0] && sp + (T - 1) * (3 * kS) + 3 < &sm.strip[0][0][0] + sizeof(sm.strip));
    B200_CHECK(o >= out && o + 3 < out + 3 * kS * kS);
    uint32_t kv[((T + 3) / 4) * 4];
#pragma unroll
    for (int i = 0; i < (T + 3) / 4; ++i) {
      const uint4 v = *reinterpret_cast<const uint4*>(kq + 4 * i);
      kv[4 * i] = v.x; kv[4 * i + 1] = v.y; kv[4 * i + 2] = v.z; kv[4 * i + 3] = v.w;
    }
    uint32_t a0 = kRnd4, a1 = kRnd4, a2 = kRnd4, a3 = kRnd4;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(sp + t * (3 * kS));
      a0 += (w & 0xffu) * kv[t];
      a1 += __byte_perm(w, 0u, 0x4441) * kv[t];
      a2 += __byte_perm(w, 0u, 0x4442) * kv[t];
      a3 += (w >> 24) * kv[t];
    }
    const float2 lo = top_byte_div255_x2(a0, a1), hi = top_byte_div255_x2(a2, a3);
    b200::stg_stream_f4(o, make_float4(lo.x, lo.y, hi.x, hi.y));
  }
}

__device__ void roi_fast_body(const uint8_t* __restrict__ frames, const uint8_t* buf_hi, int B, int H, int W,
                              int64_t pitch, int64_t bstride, int bi, int bx1, int by1, int bx2, int by2, int pad,
                              float* __restrict__ out, int* __restrict__ valid_out, FastSmem& sm, int seg) {
  // seg: -1 = produce this ROI; 0 / 1 = the launch runs every ROI through two grid segments and this CTA belongs to
  // the first / second: the first produces only the long-running crops (down-scales beyond kSegScale), the second
  // the rest -- the hardware dispatches CTAs in index order, so the long ones start first and the short ones fill
  // in behind them (longest-processing-time-first: the tail of the launch is a short crop, not a long one).
  const int tid = threadIdx.x, lane = tid & 31;
  // ---- safe_crop (detect.py:100-113) ----
  const int cx1 = max(0, min(W - 1, bx1 - pad)), cx2 = max(0, min(W, bx2 + pad));
  const int cy1 = max(0, min(H - 1, by1 - pad)), cy2 = max(0, min(H, by2 + pad));
  const int cw = cx2 - cx1, ch = cy2 - cy1;
  bool ok = (bi >= 0 && bi < B && cw > 0 && ch > 0), unsupported = false, defer = false;
  int new_w = kS, new_h = kS;
  if (ok) {
    // torchvision Resize(int): short side -> S, long side -> int(S * long / short)
    if (cw <= ch) new_h = __double2int_rz(__ddiv_rn((double)(kS * (int64_t)ch), (double)cw));
    else new_w = __double2int_rz(__ddiv_rn((double)(kS * (int64_t)cw), (double)ch));
    const double sx = (double)cw / (double)new_w, sy = (double)ch / (double)new_h;
    const double smax = fmax(fmax(sx, sy), 1.0);
    const int cs = (int)ceil(smax);
    if (seg >= 0 && (seg == 0) != (smax > kSegScale)) return;          // the other segment's CTA produces this ROI
    if (2 * cs + 1 > kMaxTaps) { ok = false; unsupported = true; }     // beyond the general body's envelope too
    else if (2 * cs + 1 > kFTaps) defer = true;                        // scale > 4: general body
  } else if (seg == 0) {
    return;                                                            // invalid ROIs are zero-filled by segment 1
  }
  if (!ok) {
    for (int i = tid; i < 3 * kS * kS / 4; i += kFT) reinterpret_cast<float4*>(out)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) *valid_out = unsupported ? -1 : 0;
    return;
  }
  if (defer) { if (tid == 0) *valid_out = 2; return; }
  const int left = half_round_even(new_w - kS), top = half_round_even(new_h - kS);

  // ---- coefficient tables: threads 0..63 horizontal (own column), 64..127 vertical (own row); 128 the barriers ----
  if (tid < kS) {
    int cnt;
    const int xm = pil_axis_fast(left + tid, cw, new_w, &sm.xk[0][tid], kS, kFTaps, cnt);
    sm.xb[tid][0] = xm; sm.xb[tid][1] = cnt;
  } else if (tid < 2 * kS) {
    const int yy = tid - kS;
    int cnt;
    const int ymin = pil_axis_fast(top + yy, ch, new_h, &sm.yk[yy][0], 1, kFKy, cnt);
    sm.yb[yy][0] = ymin; sm.yb[yy][1] = cnt;
  } else if (tid - 2 * kS < kFWarps * kRingMax) {
    b200::mbar_init(&sm.bar[0][0] + (tid - 2 * kS), 1);
    b200::mbar_fence_init();
  }
  __syncthreads();
  FastTile t;
  t.frames = frames; t.buf_hi = buf_hi; t.pitch = pitch;
  const int x_lo = sm.xb[0][0];
  t.span = (sm.xb[kS - 1][0] + sm.xb[kS - 1][1] - x_lo) * 3;   // bounds are monotone
  t.row_stride = (15 + t.span + 15) & ~15;                     // worst 16-byte phase + referenced bytes, in chunks
  t.nslots = 4 * t.row_stride <= kPool ? 4 : (3 * t.row_stride <= kPool ? 3 : (2 * t.row_stride <= kPool ? 2 : (t.row_stride <= kPool ? 1 : 0)));
  if (t.nslots < 1) {                                          // uniform: row wider than the fast envelope (cannot happen for <= kFTaps taps)
    if (tid == 0) *valid_out = 2;
    return;
  }
  // per-thread horizontal set-up: lane owns output columns lane and lane+32 of every row its warp resamples
  const int xo0 = (sm.xb[lane][0] - x_lo) * 3, xo1 = (sm.xb[lane + 32][0] - x_lo) * 3;
  const int cmax = __reduce_max_sync(0xffffffffu, max(sm.xb[lane][1], sm.xb[lane + 32][1]));     // CTA-uniform
  const int ycmax = __reduce_max_sync(0xffffffffu, max(sm.yb[lane][1], sm.yb[lane + 32][1]));    // tap classes
  uint32_t sidx = 0, parity = 0;                               // this warp's ring cursor

  int t0 = 0;
  while (t0 < kS) {
    // ---- vertical tile [t0, t1): the source rows [rmin, rmin + rows) it references fit the strip ----
    const int rmin = sm.yb[t0][0];
    int t1 = kS;
    if (sm.yb[kS - 1][0] + sm.yb[kS - 1][1] - rmin > kFRows) {          // tall crop (not a rank card): several tiles
      t1 = t0 + 1;
      while (t1 < kS && sm.yb[t1][0] + sm.yb[t1][1] - rmin <= kFRows) ++t1;
    }
    t.rows = sm.yb[t1 - 1][0] + sm.yb[t1 - 1][1] - rmin;
    t.row0 = frames + (int64_t)bi * bstride + (int64_t)(cy1 + rmin) * pitch + (int64_t)(cx1 + x_lo) * 3;
    t.inside = (t.row0 - 15 >= frames) && (t.row0 + (int64_t)(t.rows - 1) * pitch + t.span + 31 <= buf_hi);
    switch (cmax) {
      case 1: case 2: hpass<2>(sm, t, xo0, xo1, sidx, parity, tid); break;
      case 3: hpass<3>(sm, t, xo0, xo1, sidx, parity, tid); break;
      case 4: hpass<4>(sm, t, xo0, xo1, sidx, parity, tid); break;
      case 5: hpass<5>(sm, t, xo0, xo1, sidx, parity, tid); break;
      case 6: hpass<6>(sm, t, xo0, xo1, sidx, parity, tid); break;
      case 7: case 8: hpass<8>(sm, t, xo0, xo1, sidx, parity, tid); break;
      default: hpass<kFTaps>(sm, t, xo0, xo1, sidx, parity, tid); break;
    }
    __syncthreads();                                       // the strip is complete
    switch (ycmax) {
      case 1: case 2: vpass<2>(sm, t0, t1, rmin, out, tid); break;
      case 3: vpass<3>(sm, t0, t1, rmin, out, tid); break;
      case 4: vpass<4>(sm, t0, t1, rmin, out, tid); break;
      case 5: case 6: vpass<6>(sm, t0, t1, rmin, out, tid); break;
      case 7: case 8: vpass<8>(sm, t0, t1, rmin, out, tid); break;
      default: vpass<kFTaps>(sm, t0, t1, rmin, out, tid); break;
    }
    if (t1 < kS) __syncthreads();                          // the strip is rewritten by the next tile
    t0 = t1;
  }
  if (tid == 0) *valid_out = 1;
}

// ROI list form: boxes (N,4) float + batch_idx (N).
__global__ void __launch_bounds__(kFT, 6) roi_kernel(const uint8_t* __restrict__ frames, const uint8_t* buf_hi, int B,
                                                     int H, int W, int64_t pitch, int64_t bstride,
                                                     const float* __restrict__ boxes, const int* __restrict__ batch_idx,
                                                     const int* __restrict__ roi_count, int pad,
                                                     float* __restrict__ dst, int* __restrict__ valid, int N) {
  extern __shared__ __align__(16) uint8_t roi_smem[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(roi_smem);
  // grid = N (one segment) or 2N (two segments: see roi_fast_body)
  const int seg = gridDim.x > (unsigned)N ? (int)(blockIdx.x >= (unsigned)N) : -1;
  const int r = blockIdx.x - (seg > 0 ? N : 0);
  if (roi_count != nullptr && r >= *roi_count) return;
  // int() truncation of the float box (detect.py:581)
  const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)r * 4);
  roi_fast_body(frames, buf_hi, B, H, W, pitch, bstride, batch_idx[r], __float2int_rz(bx.x), __float2int_rz(bx.y),
                __float2int_rz(bx.z), __float2int_rz(bx.w), pad, dst + (int64_t)r * 3 * kS * kS, valid + r, sm, seg);
}

// Detection form (pipeline): CTA g locates the g-th detection (image-major, rank order) whose class is in
// the allow-list, from the per-image counts the NMS kernel wrote -- no separate selection launch.  The search
// is done by warp 0 alone (shuffle scans, no CTA barriers) while the other warps wait at one barrier.
__global__ void __launch_bounds__(kFT, 6) roi_det_kernel(const uint8_t* __restrict__ frames, const uint8_t* buf_hi,
                                                         int B, int H, int W, int64_t pitch, int64_t bstride,
                                                         const float* __restrict__ det,
                                                         const int* __restrict__ det_count,
                                                         const int* __restrict__ roi_cnt, int max_det,
                                                         const uint32_t* __restrict__ class_mask, int nc, int pad,
                                                         float* __restrict__ dst, int* __restrict__ roi_batch,
                                                         int* __restrict__ roi_det, int* __restrict__ valid,
                                                         int* __restrict__ roi_total, int roi_cap) {
  extern __shared__ __align__(16) uint8_t roi_smem[];
  FastSmem& sm = *reinterpret_cast<FastSmem*>(roi_smem);
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  if (tid < 32) {
    // ---- (1) which image: running prefix over roi_cnt[0..B) in chunks of 32 ----
    int carry = 0, selb = -1, selk = 0;
    for (int base = 0; base < B && (selb < 0 || g == 0); base += 32) {     // CTA 0 also needs the grand total
      const int b = base + lane;
      const int v = b < B ? roi_cnt[b] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const int excl = carry + inc - v;
      const unsigned hit = __ballot_sync(0xffffffffu, selb < 0 && b < B && g >= excl && g < excl + v);
      if (hit) {
        const int src = __ffs(hit) - 1;
        selb = base + src;
        selk = g - __shfl_sync(0xffffffffu, excl, src);
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (g == 0 && lane == 0) *roi_total = carry;      // not clamped: > roi_cap tells the host that ROIs were dropped
    // ---- (2) the selk-th allowed detection of image selb, in kept (score) order ----
    int seli = -1;
    if (selb >= 0) {
      const int n = min(det_count[selb], max_det);
      int seen = 0;
      for (int base = 0; base < n && seli < 0; base += 32) {
        const int i = base + lane;
        bool w = false;
        if (i < n) {
          const int c = (int)det[((int64_t)selb * max_det + i) * 6 + 5];
          w = c >= 0 && c < nc && ((class_mask[c >> 5] >> (c & 31)) & 1u);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, w);
        const unsigned me = __ballot_sync(0xffffffffu, w && seen + __popc(bal & ((1u << lane) - 1u)) == selk);
        if (me) seli = base + __ffs(me) - 1;
        seen += __popc(bal);
      }
    }
    if (lane == 0) { sm.sel[0] = seli >= 0 ? selb : -1; sm.sel[1] = seli; }
  }
  __syncthreads();
  const int b = sm.sel[0], i = sm.sel[1];
  if (b < 0) return;                       // g >= number of ROIs in this batch
  const float* row = det + ((int64_t)b * max_det + i) * 6;
  if (tid == 0) { roi_batch[g] = b; roi_det[g] = i; }
  roi_fast_body(frames, buf_hi, B, H, W, pitch, bstride, b, __float2int_rz(row[0]), __float2int_rz(row[1]),
                __float2int_rz(row[2]), __float2int_rz(row[3]), pad, dst + (int64_t)g * 3 * kS * kS, valid + g, sm, -1);
}

// Second launch of K5: the ROIs the first launch marked valid == 2 (crop area > kBigArea).  CTA (j, part)
// grid-strides over the deferred list (j-th marked slot, found by a block scan over valid[]) and produces
// kS/kBigParts output rows of it.  When nothing was deferred every CTA exits after the scan.
__global__ void __launch_bounds__(kThreads, 2) roi_big_kernel(const uint8_t* __restrict__ frames, const uint8_t* buf_hi,
                                                              int B, int H, int W, int64_t pitch, int64_t bstride,
                                                              const float* __restrict__ boxes,      // list form
                                                              const int* __restrict__ batch_idx,    // list form / roi_batch
                                                              const float* __restrict__ det,        // detection form
                                                              const int* __restrict__ roi_det, int max_det,
                                                              const int* __restrict__ count_ptr, int N, int pad,
                                                              float* __restrict__ dst, int* __restrict__ valid) {
  extern __shared__ __align__(16) uint8_t roi_smem[];
  RoiSmem& sm = *reinterpret_cast<RoiSmem*>(roi_smem);
  __shared__ int wsum[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = count_ptr ? min(*count_ptr, N) : N;
  {   // the common case: the fast path produced everything -- one pass over valid[], one barrier, done
    int any = 0;
    for (int g = tid; g < n; g += kThreads) any |= (valid[g] == 2);
    if (!__syncthreads_or(any)) return;
  }
  for (int j = blockIdx.x;; j += gridDim.x) {
    // locate the j-th deferred slot
    if (tid == 0) sm.sel[0] = -1;
    int carry = 0;
    for (int base = 0; base < n && carry <= j; base += kThreads) {
      const int g = base + tid;
      const int v = (g < n && valid[g] == 2) ? 1 : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, v);
      if (lane == 0) wsum[wid] = __popc(bal);
      __syncthreads();
      int wbase = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) { if (w < wid) wbase += wsum[w]; tot += wsum[w]; }
      if (v && carry + wbase + __popc(bal & ((1u << lane) - 1u)) == j) sm.sel[0] = g;
      carry += tot;
      __syncthreads();
    }
    __syncthreads();
    const int g = sm.sel[0];
    __syncthreads();
    if (g < 0) return;                      // fewer than j+1 deferred ROIs: done (uniform)
    int bi;
    const float* bx;
    if (det != nullptr) { bi = batch_idx[g]; bx = det + ((int64_t)bi * max_det + roi_det[g]) * 6; }
    else { bi = batch_idx[g]; bx = boxes + (int64_t)g * 4; }
    roi_body(frames, buf_hi, B, H, W, pitch, bstride, bi, __float2int_rz(bx[0]), __float2int_rz(bx[1]),
             __float2int_rz(bx[2]), __float2int_rz(bx[3]), pad, dst + (int64_t)g * 3 * kS * kS, valid + g, sm,
             blockIdx.y, kBigParts);
    __syncthreads();
  }
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
  if (tid == 0) *roi_count = carry;                 // not clamped: > roi_cap tells the host that ROIs were dropped
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

static size_t roi_smem_bytes() { return sizeof(RoiSmem); }
static size_t roi_fast_smem_bytes() { return sizeof(FastSmem); }

extern "C" int b200yolo_roi_crop_resize(const uint8_t* frames, int B, int H, int W, int64_t pitch,
                                        int64_t batch_stride, const float* boxes, const int* batch_idx,
                                        const int* roi_count, int N, int pad, int size, float* dst, int* valid,
                                        void* stream) {
  B200_REQUIRE(frames && boxes && batch_idx && dst && valid, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && N >= 0 && pad >= 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(pitch >= (int64_t)W * 3 && (B == 1 || batch_stride >= pitch * (int64_t)(H - 1) + (int64_t)W * 3), B200YOLO_ERR_SHAPE);
  B200_REQUIRE(size == kS, B200YOLO_ERR_UNSUPPORTED);
  if (N == 0) return B200YOLO_OK;
  B200_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, B200YOLO_ERR_ALIGN);
  const size_t smem = roi_smem_bytes(), fsmem = roi_fast_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(roi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
  if (e != cudaSuccess) return (int)e;
  const uint8_t* buf_hi = frames + (int64_t)(B - 1) * batch_stride + (int64_t)(H - 1) * pitch + (int64_t)W * 3;
  roi_kernel<<<N >= kSegMinRois ? 2 * N : N, kFT, fsmem, (cudaStream_t)stream>>>(frames, buf_hi, B, H, W, pitch, batch_stride,
                                                                                boxes, batch_idx, roi_count, pad, dst, valid, N);
  e = cudaFuncSetAttribute(roi_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  roi_big_kernel<<<dim3(kBigCtas, kBigParts), kThreads, smem, (cudaStream_t)stream>>>(
      frames, buf_hi, B, H, W, pitch, batch_stride, boxes, batch_idx, nullptr, nullptr, 0, roi_count, N, pad, dst, valid);
  return b200_launch_status();
}

extern "C" int b200yolo_roi_from_detections(const uint8_t* frames, int B, int H, int W, int64_t pitch,
                                            int64_t batch_stride, const float* det, const int* det_count,
                                            const int* roi_cnt, int max_det, const uint32_t* class_mask, int nc,
                                            int pad, int size, float* dst, int* roi_batch, int* roi_det, int* valid,
                                            int* roi_total, int roi_cap, void* stream) {
  B200_REQUIRE(frames && det && det_count && roi_cnt && class_mask && dst && roi_batch && roi_det && valid && roi_total,
               B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && H > 0 && W > 0 && max_det > 0 && nc > 0 && roi_cap > 0 && pad >= 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(pitch >= (int64_t)W * 3 && (B == 1 || batch_stride >= pitch * (int64_t)(H - 1) + (int64_t)W * 3), B200YOLO_ERR_SHAPE);
  B200_REQUIRE(size == kS, B200YOLO_ERR_UNSUPPORTED);
  const size_t smem = roi_smem_bytes(), fsmem = roi_fast_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(roi_det_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
  if (e != cudaSuccess) return (int)e;
  const uint8_t* buf_hi = frames + (int64_t)(B - 1) * batch_stride + (int64_t)(H - 1) * pitch + (int64_t)W * 3;
  roi_det_kernel<<<roi_cap, kFT, fsmem, (cudaStream_t)stream>>>(frames, buf_hi, B, H, W, pitch, batch_stride, det,
                                                                    det_count, roi_cnt, max_det, class_mask, nc, pad,
                                                                    dst, roi_batch, roi_det, valid, roi_total, roi_cap);
  e = cudaFuncSetAttribute(roi_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  roi_big_kernel<<<dim3(kBigCtas, kBigParts), kThreads, smem, (cudaStream_t)stream>>>(
      frames, buf_hi, B, H, W, pitch, batch_stride, nullptr, roi_batch, det, roi_det, max_det, roi_total, roi_cap, pad, dst,
      valid);
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
