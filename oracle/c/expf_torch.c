/* Oracle C restatement (TEST INFRASTRUCTURE ONLY) of the expf that torch's CPU sigmoid/softmax
 * kernels use: ATen Vectorized<float>::exp() == Sleef expf_u10, FMA build (the algorithm lives in the
 * third-party dependency torch==2.8.0 / sleef, requirements.txt:88 of the reference; absent from
 * /root/reference, restated from Sleef's published xexpf).  tests/test_oracle_sleef.py checks it
 * bit-for-bit against torch.sigmoid / torch.softmax on this host; the CUDA kernels' expf_torch()
 * (manual_yolo_b200/csrc/common.cuh) is the same sequence, which is what makes decoded boxes and
 * scores bit-identical to the oracle.
 * Build: gcc -O2 -mfma -ffp-contract=off -shared -fPIC expf_torch.c -o ../_build/libexpf_torch.so -lm */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float pow2if(int q) {
  uint32_t b = (uint32_t)(q + 0x7f) << 23;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

float oracle_expf_torch1(float d) {
  int q = (int)rintf(d * 1.442695040888963407359924681001892137426645954152985934135449406931f);
  float s = fmaf((float)q, -0.693145751953125f, d);
  s = fmaf((float)q, -1.428606765330187045e-06f, s);
  float u = 0.000198527617612853646278381f;
  u = fmaf(u, s, 0.00139304355252534151077271f);
  u = fmaf(u, s, 0.00833336077630519866943359f);
  u = fmaf(u, s, 0.0416664853692054748535156f);
  u = fmaf(u, s, 0.166666671633720397949219f);
  u = fmaf(u, s, 0.5f);
  u = 1.0f + fmaf(s * s, u, s);
  u = u * pow2if(q >> 1) * pow2if(q - (q >> 1));
  if (d < -104.0f) u = 0.0f;
  if (d > 100.0f) u = INFINITY;
  return u;
}

void oracle_expf_torch(const float* x, float* y, long n) {
  for (long i = 0; i < n; i++) y[i] = oracle_expf_torch1(x[i]);
}

void oracle_sigmoid_torch(const float* x, float* y, long n) {
  for (long i = 0; i < n; i++) y[i] = 1.0f / (1.0f + oracle_expf_torch1(-x[i]));
}
