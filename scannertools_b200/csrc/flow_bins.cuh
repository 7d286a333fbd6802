// FlowHistogram bin arithmetic shared by the stand-alone FlowHistogram kernel (hist.cu) and the
// fused epilogue of the last Farneback iteration (farneback.cu).
// Reproduces cv::cartToPolar(angleInDegrees) + cv::calcHist(64 bins) bit for bit
// (scannertools/old/cpp_ops/flow_histogram_kernel_cpu.cpp:33-49).
#pragma once
#include "stb_rt.h"

namespace stb {

__device__ __forceinline__ void flow_bins(float x, float y, int& bm, int& ba) {
  const float m = __fsqrt_rn(__fmaf_rn(x, x, __fmul_rn(y, y)));
  // magnitude: floor(double(m) * 1.0); floor of a float is exact in float
  bm = (m < 64.0f) ? __float2int_rd(m) : -1;  // NaN compares false -> dropped, as cvFloor garbage is
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = fabsf(x), ay = fabsf(y);
  const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
  float a;
  if (ax >= ay) {
    const float c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    const float c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(c2, p7, p5), c2, p3), c2, p1), c);
  } else {
    const float c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    const float c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.0f, __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(c2, p7, p5), c2, p3), c2, p1), c));
  }
  if (x < 0.0f) a = __fsub_rn(180.0f, a);
  if (y < 0.0f) a = __fsub_rn(360.0f, a);
  const double t = __dmul_rn((double)a, 64.0 / 360.0);
  const int ia = __double2int_rd(t);
  ba = (ia >= 0 && ia < 64) ? ia : -1;
}


#ifdef STB_CPU_EMU
#define STB_NOINLINE __attribute__((noinline))
__device__ __forceinline__ float flow_sqrt_approx(float s) { return sqrtf(s); }
__device__ __forceinline__ float flow_rcp_approx(float s) { return 1.0f / s; }
#else
#define STB_NOINLINE __noinline__
// MUFU.SQRT / MUFU.RCP, flush-to-zero: one instruction each (rsqrtf() / __fdividef() carry denormal scaling)
__device__ __forceinline__ float flow_sqrt_approx(float s) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s)); return r; }
__device__ __forceinline__ float flow_rcp_approx(float s) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s)); return r; }
#endif
// out of line, result by value (bm | ba << 16, 0xffff = dropped): reference parameters of a non-inlined
// function would live in local memory on the hot path too
static __device__ STB_NOINLINE unsigned flow_bins_exact_slow(float x, float y) {
  int bm, ba;
  flow_bins(x, y, bm, ba);
  return ((unsigned)bm & 0xffffu) | ((unsigned)ba << 16);
}

// Same bins as flow_bins, cheaper on average: the bins are first located with approximate (MUFU)
// square root / reciprocal arithmetic, no double precision and no float -> int conversions; only values
// that land within a guard band of a bin edge (4e-6 * (m + 1) for the magnitude against an approximation
// error <= 2^-22 relative; 8e-5 bins = 4.5e-4 degrees for the angle against <= ~1.2e-5 bins: reciprocal
// 2^-22 relative on the atan argument, one ulp of 360 in the quadrant folding, the float instead of double
// bin scaling) -- and anything non-finite, tiny or huge -- take the exact IEEE path above.  A value outside the
// guard band cannot change bin, so the histogram is bit-identical to the exact path's (tests compare the
// two on edge-heavy fields).  A bin is valid iff (unsigned)bin < 64.
//   rint / floor without FRND / F2I (quarter-rate pipe): t1 = v + 1.5 * 2^23 holds rint(v) in its low
//   22 mantissa bits for 0 <= v < 2^22; outside the guard band floor(v) = rint(v) - (v < rint(v)).
__device__ __forceinline__ void flow_bins_fast(float x, float y, int& bm, int& ba) {
  const float kMagic = 12582912.0f;
  const float s = __fmaf_rn(x, x, __fmul_rn(y, y));
  // s > 1e-15: max(|x|, |y|) > 2.2e-8, so OpenCV's "+ DBL_EPSILON" in the atan denominator is below half an ulp
  bool exact = !(s > 1e-15f && s < 1e30f);
  const float m = fminf(flow_sqrt_approx(s), 100.0f);
  const float m1 = __fadd_rn(m, kMagic);
  const float dm = __fsub_rn(m, __fsub_rn(m1, kMagic));            // m - rint(m), in [-0.5, 0.5]
  exact |= fabsf(dm) <= __fmaf_rn(m, 4e-6f, 4e-6f);
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = fabsf(x), ay = fabsf(y);
  const float c = fminf(ax, ay) * flow_rcp_approx(fmaxf(ax, ay));
  const float c2 = c * c;
  float a = fmaf(fmaf(fmaf(c2, p7, p5), c2, p3), c2, p1) * c;
  if (ax < ay) a = 90.0f - a;
  if (x < 0.0f) a = 180.0f - a;
  if (y < 0.0f) a = 360.0f - a;
  const float t = a * (64.0f / 360.0f);                            // in [0, 64]
  const float t1 = __fadd_rn(t, kMagic);
  const float dt = __fsub_rn(t, __fsub_rn(t1, kMagic));
  exact |= fabsf(dt) <= 8e-5f;
  if (__builtin_expect(exact, 0)) {
    const unsigned r = flow_bins_exact_slow(x, y);   // rare (< 0.1 % of pixels on generic content)
    bm = (int)(r & 0xffffu);
    ba = (int)(r >> 16);
    return;
  }
  bm = (__float_as_int(m1) & 0x3fffff) + (__float_as_int(dm) >> 31);
  ba = (__float_as_int(t1) & 0x3fffff) + (__float_as_int(dt) >> 31);
}

}  // namespace stb
