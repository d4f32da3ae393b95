// N2: the tracker the reference hands every frame's detections to -- supervision's ByteTrack
// (detect.py:22 `sv.ByteTrack()`, detect.py:557 `tracker.update_with_detections`).  Its per-frame arithmetic on the
// track states runs here, batched over tracks (one thread per track), on state arrays that stay resident on the device:
//
//   kalman_predict   STrack.multi_predict: 8-state constant-velocity filter (x, y, a, h, vx, vy, va, vh), fp64;
//                    mean <- F mean, P <- F P F^T + Q(h), with vh zeroed first for tracks that are not in the Tracked
//                    state (mode 1; mode 2 = no prediction: unconfirmed tracks); also emits each track's box (tlbr,
//                    fp32) -- the rows of the association cost matrix
//   kalman_update    KalmanFilter.update for matched (track, detection) pairs: measurement = xyah of the detection box,
//                    S = H P H^T + R(h), K = P H^T S^-1 through a 4x4 Cholesky solve, mean += K (z - H mean),
//                    P -= K S K^T
//   kalman_initiate  KalmanFilter.initiate for new tracks
//
// The association itself is b200yolo_iou_cost_matrix (assoc.cu) + a linear assignment on the host; the track
// lifecycle (activation, loss, removal, duplicate removal) is host bookkeeping (manual_yolo_b200/tracking.py).
// supervision (==0.26.1, requirements.txt:83) is not installed here: restated from the published ByteTrack /
// supervision algorithm, checked against the numpy restatement in oracle/bytetrack.py -- PARITY UNPINNED.
// Compiled with -fmad=false: every product and sum rounds separately, as the numpy expressions do.
#include "common.cuh"

namespace {

constexpr double kWp = 1.0 / 20.0;    // _std_weight_position
constexpr double kWv = 1.0 / 160.0;   // _std_weight_velocity

__device__ __forceinline__ double sq(double v) { return v * v; }

__global__ void kalman_predict_kernel(double* __restrict__ mean, double* __restrict__ cov, const int* __restrict__ slots,
                                      const int* __restrict__ zero_vh, int n, float* __restrict__ tlbr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int s = slots[t];
  double* m = mean + (int64_t)s * 8;
  double* P = cov + (int64_t)s * 64;
  const int mode = zero_vh[t];                                 // 0 predict, 1 zero vh then predict, 2 emit the box only
  if (mode != 2) {
    if (mode == 1) m[7] = 0.0;                                 // multi_predict: lost tracks do not keep growing
    const double h = m[3];
    const double q[8] = {sq(kWp * h), sq(kWp * h), sq(1e-2), sq(kWp * h), sq(kWv * h), sq(kWv * h), sq(1e-5), sq(kWv * h)};
    // mean <- F mean  (F = [[I, I], [0, I]])
    for (int i = 0; i < 4; ++i) m[i] = m[i] + m[4 + i];
    // P <- F P F^T + Q:  [[A + C + B + D, B + D], [C + D, D]]  (A, B, C, D the 4x4 blocks of P)
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        const double A = P[i * 8 + j], B = P[i * 8 + 4 + j], C = P[(4 + i) * 8 + j], D = P[(4 + i) * 8 + 4 + j];
        P[i * 8 + j] = (A + C) + (B + D);
        P[i * 8 + 4 + j] = B + D;
        P[(4 + i) * 8 + j] = C + D;
      }
    for (int i = 0; i < 8; ++i) P[i * 8 + i] += q[i];
  }
  if (tlbr) {                                                  // STrack.tlbr of the predicted state
    const double w = m[2] * m[3], x0 = m[0] - w / 2.0, y0 = m[1] - m[3] / 2.0;
    float* o = tlbr + (int64_t)t * 4;
    o[0] = (float)x0; o[1] = (float)y0; o[2] = (float)(x0 + w); o[3] = (float)(y0 + m[3]);
  }
}

__device__ __forceinline__ void xyah_of(const float* b, double* z) {     // STrack.tlbr_to_tlwh + tlwh_to_xyah
  const double w = (double)b[2] - (double)b[0], h = (double)b[3] - (double)b[1];
  z[0] = (double)b[0] + w / 2.0; z[1] = (double)b[1] + h / 2.0; z[2] = w / h; z[3] = h;
}

__global__ void kalman_update_kernel(double* __restrict__ mean, double* __restrict__ cov, const int* __restrict__ slots,
                                     const float* __restrict__ boxes, int box_stride, const int* __restrict__ box_idx,
                                     int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int s = slots[t];
  double* m = mean + (int64_t)s * 8;
  double* P = cov + (int64_t)s * 64;
  double z[4];
  xyah_of(boxes + (int64_t)box_idx[t] * box_stride, z);
  const double h = m[3];
  const double r[4] = {sq(kWp * h), sq(kWp * h), sq(1e-1), sq(kWp * h)};
  // S = H P H^T + R (upper-left 4x4 of P), Cholesky S = L L^T
  double L[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = P[i * 8 + j] + (i == j ? r[i] : 0.0);
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = i == j ? sqrt(v) : v / L[j][j];
    }
  // K^T = S^-1 (P H^T)^T: for every state row i solve S x = P[i, 0:4]
  double K[8][4];
  for (int i = 0; i < 8; ++i) {
    double y[4];
    for (int a = 0; a < 4; ++a) {
      double v = P[i * 8 + a];
      for (int k = 0; k < a; ++k) v -= L[a][k] * y[k];
      y[a] = v / L[a][a];
    }
    for (int a = 3; a >= 0; --a) {
      double v = y[a];
      for (int k = a + 1; k < 4; ++k) v -= L[k][a] * K[i][k];
      K[i][a] = v / L[a][a];
    }
  }
  double inn[4];
  for (int a = 0; a < 4; ++a) inn[a] = z[a] - m[a];
  for (int i = 0; i < 8; ++i) {
    double v = 0.0;
    for (int a = 0; a < 4; ++a) v += inn[a] * K[i][a];
    m[i] += v;
  }
  // P <- P - K S K^T, with S = H P H^T + R taken BEFORE P changes
  double S[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) S[a][b] = P[a * 8 + b] + (a == b ? r[a] : 0.0);
  double KS[8][4];
  for (int i = 0; i < 8; ++i)
    for (int b = 0; b < 4; ++b) {
      double v = 0.0;
      for (int a = 0; a < 4; ++a) v += K[i][a] * S[a][b];
      KS[i][b] = v;
    }
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) {
      double v = 0.0;
      for (int b = 0; b < 4; ++b) v += KS[i][b] * K[j][b];
      P[i * 8 + j] -= v;
    }
}

__global__ void kalman_initiate_kernel(double* __restrict__ mean, double* __restrict__ cov, const int* __restrict__ slots,
                                       const float* __restrict__ boxes, int box_stride, const int* __restrict__ box_idx,
                                       int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int s = slots[t];
  double* m = mean + (int64_t)s * 8;
  double* P = cov + (int64_t)s * 64;
  double z[4];
  xyah_of(boxes + (int64_t)box_idx[t] * box_stride, z);
  for (int i = 0; i < 4; ++i) { m[i] = z[i]; m[4 + i] = 0.0; }
  const double h = z[3];
  const double sd[8] = {2 * kWp * h, 2 * kWp * h, 1e-2, 2 * kWp * h, 10 * kWv * h, 10 * kWv * h, 1e-5, 10 * kWv * h};
  for (int i = 0; i < 64; ++i) P[i] = 0.0;
  for (int i = 0; i < 8; ++i) P[i * 8 + i] = sq(sd[i]);
}

}  // namespace

extern "C" int b200yolo_kalman_predict(double* mean, double* cov, const int* slots, const int* zero_vh, int n, float* tlbr,
                                       void* stream) {
  B200_REQUIRE(n >= 0, B200YOLO_ERR_SHAPE);
  if (n == 0) return B200YOLO_OK;
  B200_REQUIRE(mean && cov && slots && zero_vh, B200YOLO_ERR_NULL);
  kalman_predict_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(mean, cov, slots, zero_vh, n, tlbr);
  return b200_launch_status();
}

extern "C" int b200yolo_kalman_update(double* mean, double* cov, const int* slots, const float* boxes, int box_stride,
                                      const int* box_idx, int n, void* stream) {
  B200_REQUIRE(n >= 0 && box_stride >= 4, B200YOLO_ERR_SHAPE);
  if (n == 0) return B200YOLO_OK;
  B200_REQUIRE(mean && cov && slots && boxes && box_idx, B200YOLO_ERR_NULL);
  kalman_update_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(mean, cov, slots, boxes, box_stride, box_idx, n);
  return b200_launch_status();
}

extern "C" int b200yolo_kalman_initiate(double* mean, double* cov, const int* slots, const float* boxes, int box_stride,
                                        const int* box_idx, int n, void* stream) {
  B200_REQUIRE(n >= 0 && box_stride >= 4, B200YOLO_ERR_SHAPE);
  if (n == 0) return B200YOLO_OK;
  B200_REQUIRE(mean && cov && slots && boxes && box_idx, B200YOLO_ERR_NULL);
  kalman_initiate_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(mean, cov, slots, boxes, box_stride, box_idx, n);
  return b200_launch_status();
}
