// Resize op: cv::resize(..., INTER_LINEAR) on 8-bit frames, bit-exact with
// OpenCV's fixed-point path (scannertools_cpp/imgproc/resize_kernel.cpp:69-71 on CPU,
// cvc::resize :75-79 on GPU).  It is the step in front of OpticalFlow in the shipped
// flow-histogram pipeline (scannertools/old/histograms.py:64-68, 426x240).
//
// One thread per output pixel (all channels).  11-bit coefficients cvRound(f * 2048), horizontal
// taps in int, vertical (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  The x
// index/fraction are clamped when the coefficient is formed, y keeps the unclamped fraction and
// clips the two row indices -- exactly OpenCV's asymmetry (it changes results on up-scaled
// border rows).  Exact 2x down-scaling takes OpenCV's INTER_AREA fast path (a+b+c+d+2)>>2.
#include "stb_rt.h"

namespace stb {

struct PtrAddrPairU8 {
  PtrBatch<const uint8_t> src;
  PtrBatch<uint8_t> dst;
};

__device__ __forceinline__ int cv_round_f(float v) { return __float2int_rn(v); }

template <int CN>
__global__ void __launch_bounds__(256)
resize_linear_u8_kernel(PtrBatch<const uint8_t> srcs, PtrBatch<uint8_t> dsts, int sw, int sh, int dw, int dh,
                        double scale_x, double scale_y, int area2x) {
  const int dx = blockIdx.x * 32 + (threadIdx.x & 31);
  const int dy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (dx >= dw || dy >= dh) return;
  const uint8_t* src = srcs.p[blockIdx.z];
  uint8_t* dst = dsts.p[blockIdx.z] + ((size_t)dy * dw + dx) * CN;
  if (area2x) {
    const uint8_t* p = src + ((size_t)(2 * dy) * sw + 2 * dx) * CN;
    const size_t rs = (size_t)sw * CN;
#pragma unroll
    for (int c = 0; c < CN; ++c) dst[c] = (uint8_t)((p[c] + p[CN + c] + p[rs + c] + p[rs + CN + c] + 2) >> 2);
    return;
  }
  float fx = (float)((dx + 0.5) * scale_x - 0.5);
  int sx = __float2int_rd(fx);
  fx -= (float)sx;
  if (sx < 0) { fx = 0.f; sx = 0; }
  if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
  const int sx1 = sx + 1 < sw ? sx + 1 : sx;
  const int a0 = cv_round_f(__fmul_rn(1.f - fx, 2048.f)), a1 = cv_round_f(__fmul_rn(fx, 2048.f));
  float fy = (float)((dy + 0.5) * scale_y - 0.5);
  const int sy = __float2int_rd(fy);
  fy -= (float)sy;
  const int b0 = cv_round_f(__fmul_rn(1.f - fy, 2048.f)), b1 = cv_round_f(__fmul_rn(fy, 2048.f));
  const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
  const uint8_t* r0 = src + (size_t)y0 * sw * CN;
  const uint8_t* r1 = src + (size_t)y1 * sw * CN;
#pragma unroll
  for (int c = 0; c < CN; ++c) {
    const int h0 = r0[sx * CN + c] * a0 + r0[sx1 * CN + c] * a1;
    const int h1 = r1[sx * CN + c] * a0 + r1[sx1 * CN + c] * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    dst[c] = (uint8_t)min(max(v, 0), 255);
  }
}

}  // namespace stb

using namespace stb;

extern "C" {

int stb_resize_target(int frame_w, int frame_h, int width, int height, int min_flag, int preserve_aspect, int* out_w,
                      int* out_h) {
  // resize_kernel.cpp:43-61, verbatim semantics (integer division included)
  if (frame_w <= 0 || frame_h <= 0 || !out_w || !out_h) { set_error("stb_resize_target: invalid argument"); return STB_ERR_INVALID; }
  int tw = width, th = height;
  if (preserve_aspect) {
    if (tw == 0) tw = frame_w * th / frame_h;
    else th = frame_h * tw / frame_w;
  }
  if (min_flag) {
    if (frame_w <= tw && frame_h <= th) { tw = frame_w; th = frame_h; }
  }
  *out_w = tw; *out_h = th;
  return STB_OK;
}

int stb_resize_bilinear_u8(const uint8_t* const* d_src, int n, int src_w, int src_h, int channels, uint8_t* const* d_dst,
                           int dst_w, int dst_h, stb_stream_t stream) {
  if (n == 0) return STB_OK;
  if (!d_src || !d_dst || n < 0 || src_w <= 0 || src_h <= 0 || dst_w <= 0 || dst_h <= 0) {
    set_error("stb_resize_bilinear_u8: invalid argument (n=%d, %dx%d -> %dx%d)", n, src_w, src_h, dst_w, dst_h);
    return STB_ERR_INVALID;
  }
  if (channels != 1 && channels != 3 && channels != 4) {
    set_error("stb_resize_bilinear_u8: %d channels not supported (1, 3, 4)", channels);
    return STB_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const double scale_x = 1. / ((double)dst_w / src_w), scale_y = 1. / ((double)dst_h / src_h);
  const int area2x = (src_w == 2 * dst_w && src_h == 2 * dst_h) ? 1 : 0;
  for (int base = 0; base < n; base += kMaxPtrBatch) {
    const int m = n - base < kMaxPtrBatch ? n - base : kMaxPtrBatch;
    PtrBatch<const uint8_t> a;
    PtrBatch<uint8_t> b;
    for (int i = 0; i < kMaxPtrBatch; ++i) { a.p[i] = nullptr; b.p[i] = nullptr; }
    for (int i = 0; i < m; ++i) {
      if (!d_src[base + i] || !d_dst[base + i]) { set_error("stb_resize_bilinear_u8: NULL frame %d", base + i); return STB_ERR_INVALID; }
      a.p[i] = d_src[base + i];
      b.p[i] = d_dst[base + i];
    }
    const dim3 grid(ceil_div(dst_w, 32), ceil_div(dst_h, 8), m);
    if (channels == 3) stb_launch(resize_linear_u8_kernel<3>, grid, dim3(256), 0, s, a, b, src_w, src_h, dst_w, dst_h, scale_x, scale_y, area2x);
    else if (channels == 1) stb_launch(resize_linear_u8_kernel<1>, grid, dim3(256), 0, s, a, b, src_w, src_h, dst_w, dst_h, scale_x, scale_y, area2x);
    else stb_launch(resize_linear_u8_kernel<4>, grid, dim3(256), 0, s, a, b, src_w, src_h, dst_w, dst_h, scale_x, scale_y, area2x);
    STB_CHECK_LAUNCH("resize_linear_u8_kernel");
  }
  return STB_OK;
}

}  // extern "C"
