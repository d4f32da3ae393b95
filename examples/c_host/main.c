/*
 * Plain-C host for libb200yolo.so: no Python, no torch -- device memory from cudaMalloc, one stream.
 *
 *   gcc -std=c99 -O2 -I include -I /usr/local/cuda/include examples/c_host/main.c \
 *       -L manual_yolo_b200 -lb200yolo -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/manual_yolo_b200 -o c_host
 *
 * Runs the path of the reference's `model(frame)` call (detect.py:541) on one synthetic 1600x900 BGR frame and a
 * synthetic Detect-head tensor with three planted objects: letterbox -> class filter -> fused post-processing
 * (DFL decode + sort + NMS + scale_boxes) -> ROI crops, then prints the detections.  Exit code 0 iff the planted
 * objects come back (one detection each, right class, box within a pixel) and the letterbox padding is 114/255.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200yolo.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
#define BK(x) do { int r_ = (x); if (r_ != B200YOLO_OK) { fprintf(stderr, "b200yolo: %s (%d) at %d\n", b200yolo_strerror(r_), r_, __LINE__); return 3; } } while (0)

enum { H = 900, W = 1600, IN = 640, NC = 64, A = 8400, NO = 64 + NC, CAP = 1024, MAXDET = 300, ROICAP = 8 };

int main(void) {
  /* letterbox geometry of LetterBox((640,640), auto=False) for 1600x900: scale 0.4 -> 640x360, pad 140/140 */
  const int new_w = 640, new_h = 360, top = 140, left = 0;
  unsigned char* h_frame = (unsigned char*)malloc((size_t)H * W * 3);
  for (size_t i = 0; i < (size_t)H * W * 3; ++i) h_frame[i] = (unsigned char)((i * 2654435761u) >> 24);
  float* h_head = (float*)malloc(sizeof(float) * NO * A);
  for (int c = 0; c < NO; ++c)
    for (int a = 0; a < A; ++a) h_head[(size_t)c * A + a] = c < 64 ? 0.f : -8.f;      /* flat DFL, no class fires */
  /* three objects on the stride-8 level (80x80 cells): anchor (ax, ay), distances l,t,r,b in cells, class, logit */
  const int obj[3][7] = {{20, 30, 2, 3, 4, 2, 6}, {50, 25, 5, 5, 5, 5, 11}, {60, 40, 1, 6, 3, 2, 37}};
  for (int o = 0; o < 3; ++o) {
    const int a = obj[o][1] * 80 + obj[o][0];
    for (int s = 0; s < 4; ++s)
      for (int k = 0; k < 16; ++k) h_head[(size_t)(s * 16 + k) * A + a] = k == obj[o][2 + s] ? 30.f : -30.f;   /* one-hot DFL */
    h_head[(size_t)(64 + obj[o][6]) * A + a] = 3.0f;                                    /* sigmoid(3) = 0.95 */
  }

  unsigned char* d_frame; float *d_head, *d_in, *d_cand, *d_det, *d_scale, *d_rois;
  int *d_canchor, *d_ccount, *d_danchor, *d_dcount, *d_roicnt, *d_rb, *d_rd, *d_valid, *d_total;
  unsigned int* d_mask;
  CK(cudaMalloc((void**)&d_frame, (size_t)H * W * 3)); CK(cudaMalloc((void**)&d_head, sizeof(float) * NO * A));
  CK(cudaMalloc((void**)&d_in, sizeof(float) * 3 * IN * IN)); CK(cudaMalloc((void**)&d_cand, sizeof(float) * CAP * 6));
  CK(cudaMalloc((void**)&d_canchor, sizeof(int) * CAP)); CK(cudaMalloc((void**)&d_ccount, sizeof(int)));
  CK(cudaMalloc((void**)&d_det, sizeof(float) * MAXDET * 6)); CK(cudaMalloc((void**)&d_danchor, sizeof(int) * MAXDET));
  CK(cudaMalloc((void**)&d_dcount, sizeof(int))); CK(cudaMalloc((void**)&d_scale, sizeof(float) * 5));
  CK(cudaMalloc((void**)&d_roicnt, sizeof(int))); CK(cudaMalloc((void**)&d_mask, sizeof(unsigned int) * 2));
  CK(cudaMalloc((void**)&d_rois, sizeof(float) * ROICAP * 3 * 64 * 64)); CK(cudaMalloc((void**)&d_rb, sizeof(int) * ROICAP));
  CK(cudaMalloc((void**)&d_rd, sizeof(int) * ROICAP)); CK(cudaMalloc((void**)&d_valid, sizeof(int) * ROICAP));
  CK(cudaMalloc((void**)&d_total, sizeof(int)));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  CK(cudaMemcpyAsync(d_frame, h_frame, (size_t)H * W * 3, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_head, h_head, sizeof(float) * NO * A, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(d_ccount, 0, sizeof(int), st));
  /* ops.scale_boxes parameters: gain, pad_x, pad_y, w0, h0 */
  const float scale[5] = {0.4f, 0.f, 140.f, (float)W, (float)H};
  CK(cudaMemcpyAsync(d_scale, scale, sizeof(scale), cudaMemcpyHostToDevice, st));
  /* the *_rank classes of the reference (roadmap1.v3i.yolov8/data.yaml): 6, 11, 16, 21, 26, 37, 43 */
  const unsigned int mask[2] = {(1u << 6) | (1u << 11) | (1u << 16) | (1u << 21) | (1u << 26), (1u << (37 - 32)) | (1u << (43 - 32))};
  CK(cudaMemcpyAsync(d_mask, mask, sizeof(mask), cudaMemcpyHostToDevice, st));

  BK(b200yolo_letterbox_u8_to_f32(d_frame, 1, H, W, (int64_t)W * 3, (int64_t)H * W * 3, d_in, IN, IN, new_w, new_h, top, left,
                                  114, 1, st));
  b200yolo_level lv[3] = {{d_head, (int64_t)NO * A, A, 80, 80, 8.f}, {d_head + 6400, (int64_t)NO * A, A, 40, 40, 16.f},
                          {d_head + 8000, (int64_t)NO * A, A, 20, 20, 32.f}};
  BK(b200yolo_class_filter(lv, 3, 1, NC, 0.25f, NULL, d_cand, d_canchor, d_ccount, CAP, st));
  BK(b200yolo_postprocess_small(lv, 3, d_cand, d_canchor, d_ccount, 1, CAP, 30000, 0.45, 7680.f, 0, MAXDET, d_scale, d_det,
                                d_danchor, d_dcount, d_mask, NC, d_roicnt, NULL, st));
  BK(b200yolo_roi_from_detections(d_frame, 1, H, W, (int64_t)W * 3, (int64_t)H * W * 3, d_det, d_dcount, d_roicnt, MAXDET, d_mask,
                                  NC, 6, 64, d_rois, d_rb, d_rd, d_valid, d_total, ROICAP, st));
  float det[MAXDET * 6], pad_px[4]; int n = 0, n_roi = 0;
  CK(cudaMemcpyAsync(&n, d_dcount, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(det, d_det, sizeof(det), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&n_roi, d_total, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pad_px, d_in, sizeof(pad_px), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));

  printf("libb200yolo %d: %d detections, %d ROIs, padding value %.9g\n", b200yolo_version(), n, n_roi, pad_px[0]);
  int ok = n == 3 && n_roi == 3 && fabsf(pad_px[0] - 114.f / 255.f) < 1e-7f;
  for (int i = 0; i < n; ++i) {
    printf("  [%7.2f %7.2f %7.2f %7.2f] conf %.4f class %d\n", det[i * 6], det[i * 6 + 1], det[i * 6 + 2], det[i * 6 + 3], det[i * 6 + 4],
           (int)det[i * 6 + 5]);
    int found = 0;
    for (int o = 0; o < 3; ++o) {
      /* letterboxed box of the planted object, then (x - pad) / gain */
      const float x1 = ((obj[o][0] + 0.5f - obj[o][2]) * 8.f) / 0.4f, y1 = ((obj[o][1] + 0.5f - obj[o][3]) * 8.f - 140.f) / 0.4f;
      const float x2 = ((obj[o][0] + 0.5f + obj[o][4]) * 8.f) / 0.4f, y2 = ((obj[o][1] + 0.5f + obj[o][5]) * 8.f - 140.f) / 0.4f;
      if ((int)det[i * 6 + 5] == obj[o][6] && fabsf(det[i * 6] - x1) < 1.f && fabsf(det[i * 6 + 1] - y1) < 1.f &&
          fabsf(det[i * 6 + 2] - x2) < 1.f && fabsf(det[i * 6 + 3] - y2) < 1.f)
        found = 1;
    }
    ok = ok && found;
  }
  printf(ok ? "OK\n" : "MISMATCH\n");
  return ok ? 0 : 1;
}
