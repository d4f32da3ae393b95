// Shared device/host helpers for libb200yolo (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "b200yolo.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200yolo targets sm_100a (B200) only"
#endif

#define B200_NUM_SMS 148

#define B200_REQUIRE(cond, code) \
  do {                           \
    if (!(cond)) return (code);  \
  } while (0)

// Checked build (-DB200_CHECKS, libb200yolo_checked.so; tools/checked.sh): device-side assertions on every
// shared-memory ring / strip / key-array index and every global store index the kernels compute.  A failed check
// prints its location and traps (the launch fails with a CUDA error).  compute-sanitizer is closed on the build pool;
// this is the stand-in: the whole GPU test suite is run against the checked library.  No-ops in the product build.
#ifdef B200_CHECKS
#include <cstdio>
#define B200_CHECK(cond)                                                                          \
  do {                                                                                            \
    if (!(cond)) {                                                                                \
      printf("B200_CHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__,   \
             (int)blockIdx.x, (int)threadIdx.x);                                                  \
      __trap();                                                                                   \
    }                                                                                             \
  } while (0)
#else
#define B200_CHECK(cond) do { } while (0)
#endif

static inline int b200_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? B200YOLO_OK : (int)e;
}

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait suspends the thread (up to the time hint, in ns) until the phase completes: a waiting warp does not burn
// issue slots polling.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Wait for the phase with the given parity.  A transaction that never completes would otherwise hang the GPU until
// the driver's watchdog: after ~2 s of failed waits the kernel traps (the launch fails loudly with a CUDA error).
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  if (mbar_try_wait(bar_addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar_addr, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// address-based forms (32-bit shared-window addresses kept in registers by the caller)
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_addr(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar_addr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(bar_addr)
               : "memory");
}

// ---- streaming global access ----------------------------------------------------------------
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// u8 -> v/255 as torch's true fp32 division computes it (ToTensor / `im /= 255`): reciprocal
// multiply + one Newton residual step gives the correctly rounded quotient for every v in [0,255]
// (checked exhaustively by tests/test_gpu_letterbox.py::test_div255_exhaustive).
__device__ __forceinline__ float u8_div255(int v) {
  const float r = 0.003921568859368563f;  // fl32(1/255)
  float x = (float)v;
  float q = __fmul_rn(x, r);
  float e = __fmaf_rn(-q, 255.0f, x);
  return __fmaf_rn(e, r, q);
}

// expf exactly as torch's CPU kernels compute it: ATen Vectorized<float>::exp() is Sleef's
// expf_u10 (FMA build).  Restating its range reduction + degree-6 polynomial with the same
// constants and fused multiply-adds makes sigmoid / softmax results BIT-IDENTICAL to the oracle
// (oracle/head.py runs torch CPU ops; checked bitwise by tests/test_oracle_sleef.py on the host
// and tests/test_gpu_decode.py on the device).  Accuracy: 1.0 ulp.
__device__ __forceinline__ float pow2i(int q) { return __int_as_float((q + 0x7f) << 23); }
__device__ __forceinline__ float expf_torch(float d) {
  const int q = __float2int_rn(__fmul_rn(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
  const float qf = (float)q;
  float s = __fmaf_rn(qf, -0.693145751953125f, d);
  s = __fmaf_rn(qf, -1.428606765330187045e-06f, s);
  float u = 0.000198527617612853646278381f;
  u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
  u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
  u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
  u = __fmaf_rn(u, s, 0.166666671633720397949219f);
  u = __fmaf_rn(u, s, 0.5f);
  u = __fadd_rn(1.0f, __fmaf_rn(__fmul_rn(s, s), u, s));
  u = __fmul_rn(__fmul_rn(u, pow2i(q >> 1)), pow2i(q - (q >> 1)));
  if (d < -104.0f) u = 0.0f;
  if (d > 100.0f) u = __int_as_float(0x7f800000);
  return u;
}
// expf_torch for -80 <= d <= 0 (softmax arguments x - max): same bits, fewer instructions.  The polynomial
// value u lies in [0.70, 1.42]; Sleef scales it by 2^q in two exact steps (ldexp2kf).  For d >= -80, q >= -116:
// the result is a normal number, both products are exact, and the scaling is a single exponent add; the range
// checks of the general form cannot fire.  NaN propagates (q = 0).  Checked against expf_torch for every float
// in [-80, 0] by b200yolo_selftest_math (tests/test_gpu_decode.py).
__device__ __forceinline__ float expf_torch_m80_0(float d) {
  const int q = __float2int_rn(__fmul_rn(d, 1.442695040888963407359924681001892137426645954152985934135449406931f));
  const float qf = (float)q;
  float s = __fmaf_rn(qf, -0.693145751953125f, d);
  s = __fmaf_rn(qf, -1.428606765330187045e-06f, s);
  float u = 0.000198527617612853646278381f;
  u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
  u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
  u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
  u = __fmaf_rn(u, s, 0.166666671633720397949219f);
  u = __fmaf_rn(u, s, 0.5f);
  u = __fadd_rn(1.0f, __fmaf_rn(__fmul_rn(s, s), u, s));
  return __int_as_float(__float_as_int(u) + (q << 23));
}

// a / b correctly rounded, given r = RN(1/b) (__frcp_rn): q = RN(a*r) is within an ulp of a/b, the residual
// e = a - q*b is exact in one fma, and RN(q + e*r) is the correctly rounded quotient (Markstein's theorem; needs
// normal a, b, q -- the caller routes tiny numerators to __fdiv_rn).  Checked against __fdiv_rn over ~4e9 pairs of
// the softmax domain by b200yolo_selftest_math.
__device__ __forceinline__ float div_by_rcp(float a, float b, float r) {
  const float q = __fmul_rn(a, r);
  const float e = __fmaf_rn(-q, b, a);
  return __fmaf_rn(e, r, q);
}

// torch CPU sigmoid: (1 + exp(-x)).reciprocal() with a true division
__device__ __forceinline__ float sigmoid_torch(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_torch(-x)));
}

}  // namespace b200
