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
#else
#define STB_NOINLINE __noinline__
#endif
static __device__ STB_NOINLINE void flow_bins_exact_slow(float x, float y, int& bm, int& ba) { flow_bins(x, y, bm, ba); }

// Same result as flow_bins, cheaper on average: the bins are first located with approximate
// (MUFU) reciprocal / rsqrt arithmetic and no double precision; only values that land within
// a guard band of a bin edge (10x / 7x wider than the approximation error bound: 4e-6 relative
// for the magnitude, 1.5e-4 bins = 8e-4 degrees for the angle) -- and anything non-finite, zero or huge -- take the
// exact IEEE path above.  A value outside the guard band cannot change bin, so the histogram is
// bit-identical to the exact path's (tests compare the two on edge-heavy fields).
__device__ __forceinline__ void flow_bins_fast(float x, float y, int& bm, int& ba) {
  const float s = __fmaf_rn(x, x, __fmul_rn(y, y));
  const float m = s * rsqrtf(s);
  bool exact = !(s > 1e-30f && s < 1e30f);
  exact |= fabsf(m - rintf(m)) <= 4e-6f * fmaxf(m, 1.0f) && m < 65.0f;   // approximation error <= ~4e-7 relative
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  const float ax = fabsf(x), ay = fabsf(y);
  const float mn = fminf(ax, ay), mx = fmaxf(ax, ay);
  const float c = __fdividef(mn, mx + 2.2204460492503131e-16f);
  const float c2 = c * c;
  float a = fmaf(fmaf(fmaf(c2, p7, p5), c2, p3), c2, p1) * c;
  if (ax < ay) a = 90.0f - a;
  if (x < 0.0f) a = 180.0f - a;
  if (y < 0.0f) a = 360.0f - a;
  const float t = a * (64.0f / 360.0f);
  exact |= fabsf(t - rintf(t)) <= 1.5e-4f;                               // approximation error <= ~2e-5 bins
  if (__builtin_expect(exact, 0)) {
    flow_bins_exact_slow(x, y, bm, ba);   // rare (< 0.1 % of pixels): kept out of line so the hot path stays small
    return;
  }
  bm = (m < 64.0f) ? (int)m : -1;
  ba = (t < 64.0f) ? (int)t : -1;
}

}  // namespace stb
