// N2: detection -> tracker hand-off.  The reference feeds every frame's detections to
// supervision's ByteTrack (detect.py:557 tracker.update_with_detections); its association step is
//     cost = 1 - box_iou_batch(track_boxes, det_boxes)          (supervision/tracker/byte_tracker/matching.py)
//     cost = 1 - (1 - cost) * det_scores                        (fuse_score)
// followed by a linear assignment on the host.  This kernel produces that cost matrix for a batch of frames /
// streams straight from the padded NMS output, so only the (tracks x detections) costs cross to the host.
// box_iou_batch in fp32, numpy op order (compile with -fmad=false):
//     area = (x2 - x1) * (y2 - y1);  inter = max(min(br) - max(tl), 0).prod();  iou = inter / (a_t + a_d - inter)
// supervision is not installed here: restated from its published code (oracle/assoc.py), parity unpinned.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) iou_cost_kernel(const float* __restrict__ tracks, const int* __restrict__ track_count,
                                                       const float* __restrict__ det, const int* __restrict__ det_count,
                                                       int T, int max_det, int fuse_score, float pad_cost,
                                                       float* __restrict__ cost) {
  const int b = blockIdx.y;
  const int nt = min(track_count[b], T), nd = min(det_count[b], max_det);
  const float* trow = tracks + (int64_t)b * T * 4;
  const float* drow = det + (int64_t)b * max_det * 6;
  float* crow = cost + (int64_t)b * T * max_det;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < T * max_det; e += gridDim.x * blockDim.x) {
    const int t = e / max_det, d = e - t * max_det;
    float c = pad_cost;
    if (t < nt && d < nd) {
      const float4 a = *reinterpret_cast<const float4*>(trow + t * 4);
      const float bx1 = drow[d * 6], by1 = drow[d * 6 + 1], bx2 = drow[d * 6 + 2], by2 = drow[d * 6 + 3];
      const float area_t = (a.z - a.x) * (a.w - a.y), area_d = (bx2 - bx1) * (by2 - by1);
      const float w = fmaxf(fminf(a.z, bx2) - fmaxf(a.x, bx1), 0.f), h = fmaxf(fminf(a.w, by2) - fmaxf(a.y, by1), 0.f);
      const float inter = w * h;
      const float iou = inter / (area_t + area_d - inter);
      c = 1.f - iou;
      if (fuse_score) c = 1.f - (1.f - c) * drow[d * 6 + 4];
    }
    crow[e] = c;
  }
}

}  // namespace

extern "C" int b200yolo_iou_cost_matrix(const float* tracks, const int* track_count, const float* det,
                                        const int* det_count, int B, int T, int max_det, int fuse_score,
                                        float pad_cost, float* cost, void* stream) {
  B200_REQUIRE(tracks && track_count && det && det_count && cost, B200YOLO_ERR_NULL);
  B200_REQUIRE(B > 0 && B <= 65535 && T > 0 && max_det > 0, B200YOLO_ERR_SHAPE);
  B200_REQUIRE((reinterpret_cast<uintptr_t>(tracks) & 15) == 0, B200YOLO_ERR_ALIGN);
  const int per = (T * max_det + 255) / 256;
  dim3 grid((unsigned)(per < 64 ? per : 64), (unsigned)B);
  iou_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tracks, track_count, det, det_count, T, max_det, fuse_score,
                                                           pad_cost, cost);
  return b200_launch_status();
}
