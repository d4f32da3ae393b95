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
                                                            int cap) {
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
  if (tid == 0) cand_count[f] = offs[kMaxSlices];            // not clamped: the host can detect overflow past cap
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
                                                int max_det, const int* slice_xy, float* cand, int* cand_anchor,
                                                int* cand_count, int cap, void* stream) {
  B200_REQUIRE(det && det_count && slice_xy && cand && cand_anchor && cand_count, B200YOLO_ERR_NULL);
  B200_REQUIRE(n_frames > 0 && n_slices > 0 && max_det > 0 && cap > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE(n_slices <= kMaxSlices, B200YOLO_ERR_UNSUPPORTED);
  SliceOrigins so;
  so.n = n_slices;
  for (int i = 0; i < kMaxSlices; ++i) {
    so.x[i] = i < n_slices ? slice_xy[2 * i] : 0;
    so.y[i] = i < n_slices ? slice_xy[2 * i + 1] : 0;
  }
  gather_slices_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(det, det_count, max_det, so, cand, cand_anchor, cand_count,
                                                                    cap);
  return b200_launch_status();
}
