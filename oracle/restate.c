/* TEST INFRASTRUCTURE ONLY -- not part of the product, never linked into it.
 *
 * OpenCV-free plain-C restatement of the arithmetic behind the reference's per-frame
 * analysis ops.  The reference files are thin wrappers (cited per function, relative to
 * /root/reference/scannertools/); the arithmetic itself lives in OpenCV (un-vendored,
 * "opencv >= 3.4.0", scannertools/README.md:10), whose published algorithm
 * (modules/video/src/optflowgf.cpp, imgproc color/smooth/resize/histogram, core
 * mathfuncs) is restated here as documented in SURVEY.md Appendix A/B.
 *
 * Pinned by tests/test_oracle.py against goldens generated with cv2 4.13.0
 * (tests/golden/make_golden.py) and, when cv2 is importable, against cv2 live.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle_restate.so
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- cvRound: round half to even (SSE cvtsd2si) ------------------------------------- */
static int cv_round(double v) { return (int)nearbyint(v); }
static int cv_floor_f(float v) { int i = (int)v; return i - (v < (float)i); }

/* ---- gray: cv::cvtColor(COLOR_BGR2GRAY) applied to the RGB frame ----------------------
 * optical_flow_kernel_cpu.cpp:38-39; SURVEY Appendix B: 15-bit fixed point, c0 = first byte. */
ORC_API void orc_gray(const uint8_t* rgb, int n_px, uint8_t* gray) {
  for (int i = 0; i < n_px; ++i) {
    int c0 = rgb[3 * i], c1 = rgb[3 * i + 1], c2 = rgb[3 * i + 2];
    gray[i] = (uint8_t)((c0 * 3735 + c1 * 19235 + c2 * 9798 + (1 << 14)) >> 15);
  }
}

/* ---- Histogram: histogram_kernel_cpu.cpp:16-46 --------------------------------------- */
ORC_API void orc_hist_rgb16(const uint8_t* frame, int w, int h, int32_t* out /*[3][16]*/) {
  memset(out, 0, 48 * sizeof(int32_t));
  size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; ++i)
    for (int j = 0; j < 3; ++j) out[j * 16 + (frame[3 * i + j] >> 4)]++;
}

/* ---- shot scores: shot_detection.py:14-18 (Chebyshev per channel, summed) ------------- */
ORC_API void orc_shot_scores(const int32_t* hists /*[n][48]*/, int n, int32_t* S) {
  if (n > 0) S[0] = 0;
  for (int i = 1; i < n; ++i) {
    int32_t s = 0;
    for (int j = 0; j < 3; ++j) {
      int32_t m = 0;
      for (int b = 0; b < 16; ++b) {
        int32_t d = hists[i * 48 + j * 16 + b] - hists[(i - 1) * 48 + j * 16 + b];
        if (d < 0) d = -d;
        if (d > m) m = d;
      }
      s += m;
    }
    S[i] = s;
  }
}

/* ---- FrameDifference: intended semantics of frame_difference_kernel_cpu.cpp:51-61 ----- */
ORC_API void orc_frame_diff(const uint8_t* prev, const uint8_t* cur, uint8_t* out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)(cur[i] - prev[i]);
}

/* ---- FlowHistogram: flow_histogram_kernel_cpu.cpp:17-56 ------------------------------ */
static float fast_atan2_deg(float y, float x) {
  /* OpenCV core fastAtan32f polynomial */
  const float scale = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
              p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = fmaf(fmaf(fmaf(c2, p7, p5), c2, p3), c2, p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - fmaf(fmaf(fmaf(c2, p7, p5), c2, p3), c2, p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

ORC_API void orc_polar(const float* flow, int n_px, float* mag, float* deg) {
  for (int i = 0; i < n_px; ++i) {
    float x = flow[2 * i], y = flow[2 * i + 1];
    mag[i] = sqrtf(fmaf(x, x, y * y));
    deg[i] = fast_atan2_deg(y, x);
  }
}

ORC_API void orc_flow_hist(const float* flow, int w, int h, int32_t* out /*[2][64]*/) {
  memset(out, 0, 128 * sizeof(int32_t));
  size_t n = (size_t)w * h;
  const double a_mag = 64.0 / (64.0 - 0.0), a_deg = 64.0 / (360.0 - 0.0);
  for (size_t i = 0; i < n; ++i) {
    float x = flow[2 * i], y = flow[2 * i + 1];
    float m = sqrtf(fmaf(x, x, y * y));
    float d = fast_atan2_deg(y, x);
    int im = (int)floor((double)m * a_mag);
    int id = (int)floor((double)d * a_deg);
    if (im >= 0 && im < 64) out[im]++;
    if (id >= 0 && id < 64) out[64 + id]++;
  }
}

/* =======================================================================================
 * Farneback: cv::FarnebackOpticalFlow(3, 0.5, false, 15, 3, 5, 1.2, 0)->calc(prev, next)
 * optical_flow_kernel_cpu.cpp:15-16,41; algorithm as in SURVEY Appendix A.
 * ======================================================================================= */

static int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * (n - 1) - i;
  }
  return i;
}

static void gaussian_taps(int ksize, double sigma, float* taps) {
  if (ksize == 3 && sigma <= 0) { taps[0] = 0.25f; taps[1] = 0.5f; taps[2] = 0.25f; return; }
  if (sigma <= 0) sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8;
  double sum = 0, t[64];
  for (int i = 0; i < ksize; ++i) {
    double x = i - (ksize - 1) * 0.5;
    t[i] = exp(-0.5 / (sigma * sigma) * x * x);
    sum += t[i];
  }
  for (int i = 0; i < ksize; ++i) taps[i] = (float)(t[i] / sum);
}

/* separable Gaussian, float32, BORDER_REFLECT_101, rows then columns */
static void gaussian_blur(const float* src, float* dst, int w, int h, int ksize, double sigma) {
  float taps[64];
  gaussian_taps(ksize, sigma, taps);
  int r = ksize / 2;
  float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float s = taps[r] * src[(size_t)y * w + x];
      for (int k = 1; k <= r; ++k)
        s += taps[r + k] * (src[(size_t)y * w + reflect101(x - k, w)] + src[(size_t)y * w + reflect101(x + k, w)]);
      tmp[(size_t)y * w + x] = s;
    }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      float s = taps[r] * tmp[(size_t)y * w + x];
      for (int k = 1; k <= r; ++k)
        s += taps[r + k] * (tmp[(size_t)reflect101(y - k, h) * w + x] + tmp[(size_t)reflect101(y + k, h) * w + x]);
      dst[(size_t)y * w + x] = s;
    }
  free(tmp);
}

/* cv::resize INTER_LINEAR for CV_32FC<cn>: horizontal lerp then vertical lerp, float32 */
static void resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh, int cn) {
  if (sw == dw && sh == dh) { memcpy(dst, src, sizeof(float) * (size_t)sw * sh * cn); return; }
  double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  int* xofs = (int*)malloc(sizeof(int) * dw);
  float* xa = (float*)malloc(sizeof(float) * dw);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor_f(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx; xa[dx] = fx;
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor_f(fy);
    fy -= sy;
    if (sy < 0) { fy = 0; sy = 0; }
    if (sy >= sh - 1) { fy = 0; sy = sh - 1; }
    int sy1 = sy + 1 < sh ? sy + 1 : sy;
    const float* r0 = src + (size_t)sy * sw * cn;
    const float* r1 = src + (size_t)sy1 * sw * cn;
    for (int dx = 0; dx < dw; ++dx) {
      int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sx;
      float a1 = xa[dx], a0 = 1.f - a1;
      for (int c = 0; c < cn; ++c) {
        float h0 = r0[sx * cn + c] * a0 + r0[sx1 * cn + c] * a1;
        float h1 = r1[sx * cn + c] * a0 + r1[sx1 * cn + c] * a1;
        dst[((size_t)dy * dw + dx) * cn + c] = h0 * (1.f - fy) + h1 * fy;
      }
    }
  }
  free(xofs); free(xa);
}

typedef struct { float g[11], xg[11], xxg[11]; double ig11, ig03, ig33, ig55; } poly_consts;

/* 6x6 inverse by Gauss-Jordan in double (G is SPD and tiny; matches Cholesky to ~1e-16) */
static void invert6(double A[6][6], double inv[6][6]) {
  double M[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { M[i][j] = A[i][j]; M[i][6 + j] = (i == j); }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    if (p != c) for (int j = 0; j < 12; ++j) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; }
    double d = 1.0 / M[c][c];
    for (int j = 0; j < 12; ++j) M[c][j] *= d;
    for (int r = 0; r < 6; ++r) if (r != c) {
      double f = M[r][c];
      if (f != 0) for (int j = 0; j < 12; ++j) M[r][j] -= f * M[c][j];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) inv[i][j] = M[i][6 + j];
}

static void prepare_gaussian(int n, double sigma, poly_consts* pc) {
  /* FarnebackPrepareGaussian; arrays indexed by |x| (g symmetric, xg antisymmetric) */
  float gf[32];
  double s = 0;
  if (sigma < FLT_EPSILON) sigma = n * 0.3;
  for (int x = -n; x <= n; ++x) { gf[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += gf[x + n]; }
  s = 1. / s;
  for (int x = -n; x <= n; ++x) gf[x + n] = (float)(gf[x + n] * s);
  for (int x = 0; x <= n; ++x) {
    pc->g[x] = gf[x + n];
    pc->xg[x] = (float)(x * gf[x + n]);
    pc->xxg[x] = (float)(x * x * gf[x + n]);
  }
  double G[6][6], inv[6][6];
  memset(G, 0, sizeof(G));
  for (int y = -n; y <= n; ++y)
    for (int x = -n; x <= n; ++x) {
      /* OpenCV evaluates g[y]*g[x]*x*x... in float (float*int -> float), then adds to double */
      float gg = gf[y + n] * gf[x + n];
      G[0][0] += gg;
      G[1][1] += gg * x * x;
      G[3][3] += gg * x * x * x * x;
      G[5][5] += gg * x * x * y * y;
    }
  G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
  G[4][4] = G[3][3];
  G[3][4] = G[4][3] = G[5][5];
  invert6(G, inv);
  pc->ig11 = inv[1][1]; pc->ig03 = inv[0][3]; pc->ig33 = inv[3][3]; pc->ig55 = inv[5][5];
}

ORC_API void orc_poly_consts(int n, double sigma, float* g, float* xg, float* xxg, double* ig /*[4]*/) {
  poly_consts pc;
  prepare_gaussian(n, sigma, &pc);
  for (int i = 0; i <= n; ++i) { g[i] = pc.g[i]; xg[i] = pc.xg[i]; xxg[i] = pc.xxg[i]; }
  ig[0] = pc.ig11; ig[1] = pc.ig03; ig[2] = pc.ig33; ig[3] = pc.ig55;
}

/* FarnebackPolyExp: vertical float32, horizontal double accumulators, replicate borders.
 * dst: h x w x 5 interleaved (as OpenCV) */
static void poly_exp(const float* src, float* dst, int w, int h, int n, const poly_consts* pc) {
  float* rowbuf = (float*)malloc(sizeof(float) * (size_t)(w + 2 * n) * 3);
  float* row = rowbuf + n * 3;
  for (int y = 0; y < h; ++y) {
    const float* s0 = src + (size_t)y * w;
    for (int x = 0; x < w; ++x) { row[x * 3] = s0[x] * pc->g[0]; row[x * 3 + 1] = row[x * 3 + 2] = 0.f; }
    for (int k = 1; k <= n; ++k) {
      float g0 = pc->g[k], g1 = pc->xg[k], g2 = pc->xxg[k];
      const float* a = src + (size_t)(y - k > 0 ? y - k : 0) * w;
      const float* b = src + (size_t)(y + k < h - 1 ? y + k : h - 1) * w;
      for (int x = 0; x < w; ++x) {
        float p = a[x] + b[x];
        row[x * 3] = row[x * 3] + g0 * p;
        row[x * 3 + 1] = row[x * 3 + 1] + g1 * (b[x] - a[x]);
        row[x * 3 + 2] = row[x * 3 + 2] + g2 * p;
      }
    }
    for (int x = 0; x < n * 3; ++x) {
      row[-1 - x] = row[2 - (x % 3)];
      row[w * 3 + x] = row[(w - 1) * 3 + (x % 3)];
    }
    float* d = dst + (size_t)y * w * 5;
    for (int x = 0; x < w; ++x) {
      float g0 = pc->g[0];
      double b1 = row[x * 3] * g0, b2 = 0, b3 = row[x * 3 + 1] * g0, b4 = 0, b5 = row[x * 3 + 2] * g0, b6 = 0;
      for (int k = 1; k <= n; ++k) {
        double tg = row[(x + k) * 3] + row[(x - k) * 3];
        g0 = pc->g[k];
        b1 += tg * g0;
        b4 += tg * pc->xxg[k];
        b2 += (row[(x + k) * 3] - row[(x - k) * 3]) * pc->xg[k];
        b3 += (row[(x + k) * 3 + 1] + row[(x - k) * 3 + 1]) * g0;
        b6 += (row[(x + k) * 3 + 1] - row[(x - k) * 3 + 1]) * pc->xg[k];
        b5 += (row[(x + k) * 3 + 2] + row[(x - k) * 3 + 2]) * g0;
      }
      d[x * 5 + 1] = (float)(b2 * pc->ig11);
      d[x * 5] = (float)(b3 * pc->ig11);
      d[x * 5 + 3] = (float)(b1 * pc->ig03 + b4 * pc->ig33);
      d[x * 5 + 2] = (float)(b1 * pc->ig03 + b5 * pc->ig33);
      d[x * 5 + 4] = (float)(b6 * pc->ig55);
    }
  }
  free(rowbuf);
}

/* FarnebackUpdateMatrices over all rows */
static void update_matrices(const float* R0, const float* R1, const float* flow, float* M, int w, int h) {
  static const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
  const int BORDER = 5;
  size_t step1 = (size_t)w * 5;
  for (int y = 0; y < h; ++y) {
    const float* r0 = R0 + (size_t)y * w * 5;
    const float* fl = flow + (size_t)y * w * 2;
    float* m = M + (size_t)y * w * 5;
    for (int x = 0; x < w; ++x) {
      float dx = fl[x * 2], dy = fl[x * 2 + 1];
      float fx = x + dx, fy = y + dy;
      int x1 = cv_floor_f(fx), y1 = cv_floor_f(fy);
      float r2, r3, r4, r5, r6;
      fx -= x1; fy -= y1;
      if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float* p = R1 + (size_t)y1 * step1 + (size_t)x1 * 5;
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * p[0] + a01 * p[5] + a10 * p[step1] + a11 * p[step1 + 5];
        r3 = a00 * p[1] + a01 * p[6] + a10 * p[step1 + 1] + a11 * p[step1 + 6];
        r4 = a00 * p[2] + a01 * p[7] + a10 * p[step1 + 2] + a11 * p[step1 + 7];
        r5 = a00 * p[3] + a01 * p[8] + a10 * p[step1 + 3] + a11 * p[step1 + 8];
        r6 = a00 * p[4] + a01 * p[9] + a10 * p[step1 + 4] + a11 * p[step1 + 9];
        r4 = (r0[x * 5 + 2] + r4) * 0.5f;
        r5 = (r0[x * 5 + 3] + r5) * 0.5f;
        r6 = (r0[x * 5 + 4] + r6) * 0.25f;
      } else {
        r2 = r3 = 0.f;
        r4 = r0[x * 5 + 2];
        r5 = r0[x * 5 + 3];
        r6 = r0[x * 5 + 4] * 0.5f;
      }
      r2 = (r0[x * 5] - r2) * 0.5f;
      r3 = (r0[x * 5 + 1] - r3) * 0.5f;
      r2 += r4 * dy + r6 * dx;
      r3 += r6 * dy + r5 * dx;
      if ((unsigned)(x - BORDER) >= (unsigned)(w - BORDER * 2) || (unsigned)(y - BORDER) >= (unsigned)(h - BORDER * 2)) {
        float scale = (x < BORDER ? border[x] : 1.f) * (x >= w - BORDER ? border[w - x - 1] : 1.f) *
                      (y < BORDER ? border[y] : 1.f) * (y >= h - BORDER ? border[h - y - 1] : 1.f);
        r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
      }
      m[x * 5] = r4 * r4 + r6 * r6;
      m[x * 5 + 1] = (r4 + r5) * r6;
      m[x * 5 + 2] = r5 * r5 + r6 * r6;
      m[x * 5 + 3] = r4 * r2 + r6 * r3;
      m[x * 5 + 4] = r6 * r2 + r5 * r3;
    }
  }
}

/* FarnebackUpdateFlow_Blur: 15x15 box (replicate) in double, 2x2 solve; "blur all with the
 * old M, then update all" (the lagging stripes in OpenCV are semantically this). */
static void update_flow_blur(const float* R0, const float* R1, float* flow, float* M, int w, int h,
                             int block, int update) {
  int m = block / 2;
  double scale = 1. / (block * block);
  double* vs = (double*)malloc(sizeof(double) * (size_t)w * 5);
  for (int y = 0; y < h; ++y) {
    for (int i = 0; i < w * 5; ++i) vs[i] = 0;
    for (int dy = -m; dy <= m; ++dy) {
      int yy = y + dy; yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
      const float* r = M + (size_t)yy * w * 5;
      for (int i = 0; i < w * 5; ++i) vs[i] += r[i];
    }
    for (int x = 0; x < w; ++x) {
      double s[5] = {0, 0, 0, 0, 0};
      for (int dx = -m; dx <= m; ++dx) {
        int xx = x + dx; xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx);
        for (int c = 0; c < 5; ++c) s[c] += vs[xx * 5 + c];
      }
      double g11 = s[0] * scale, g12 = s[1] * scale, g22 = s[2] * scale, h1 = s[3] * scale, h2 = s[4] * scale;
      double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
      flow[((size_t)y * w + x) * 2] = (float)((g11 * h2 - g12 * h1) * idet);
      flow[((size_t)y * w + x) * 2 + 1] = (float)((g22 * h1 - g12 * h2) * idet);
    }
  }
  free(vs);
  if (update) update_matrices(R0, R1, flow, M, w, h);
}

/* FarnebackUpdateFlow_GaussianBlur (flags & OPTFLOW_FARNEBACK_GAUSSIAN): separable Gaussian window
 * sigma = 0.3 * (block/2), float taps normalised in double, float accumulation in OpenCV's order
 * (centre first, then symmetric pairs), replicate border; same 2x2 solve in double. */
static void update_flow_gaussian(const float* R0, const float* R1, float* flow, float* M, int w, int h,
                                 int block, int update) {
  int m = block / 2;
  double sigma = m * 0.3, s = 1;
  float kernel[64];
  kernel[0] = (float)s;
  for (int i = 1; i <= m; ++i) {
    float t = (float)exp(-i * i / (2 * sigma * sigma));
    kernel[i] = t;
    s += t * 2;
  }
  s = 1. / s;
  for (int i = 0; i <= m; ++i) kernel[i] = (float)(kernel[i] * s);
  float* vs = (float*)malloc(sizeof(float) * (size_t)w * 5);
  for (int y = 0; y < h; ++y) {
    const float* rc = M + (size_t)y * w * 5;
    for (int i = 0; i < w * 5; ++i) vs[i] = rc[i] * kernel[0];
    for (int d = 1; d <= m; ++d) {
      int ya = y + d > h - 1 ? h - 1 : y + d, yb = y - d < 0 ? 0 : y - d;
      const float* ra = M + (size_t)ya * w * 5;
      const float* rb = M + (size_t)yb * w * 5;
      for (int i = 0; i < w * 5; ++i) vs[i] += (ra[i] + rb[i]) * kernel[d];
    }
    for (int x = 0; x < w; ++x) {
      float hs[5];
      for (int c = 0; c < 5; ++c) hs[c] = vs[x * 5 + c] * kernel[0];
      for (int d = 1; d <= m; ++d) {
        int xa = x - d < 0 ? 0 : x - d, xb = x + d > w - 1 ? w - 1 : x + d;
        for (int c = 0; c < 5; ++c) hs[c] += kernel[d] * (vs[xa * 5 + c] + vs[xb * 5 + c]);
      }
      double g11 = hs[0], g12 = hs[1], g22 = hs[2], h1 = hs[3], h2 = hs[4];
      double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
      flow[((size_t)y * w + x) * 2] = (float)((g11 * h2 - g12 * h1) * idet);
      flow[((size_t)y * w + x) * 2 + 1] = (float)((g22 * h1 - g12 * h2) * idet);
    }
  }
  free(vs);
  if (update) update_matrices(R0, R1, flow, M, w, h);
}

typedef struct {
  int levels;       /* number of scales actually processed (<= 4 for numLevels = 3) */
  int w[8], h[8];   /* per scale k */
} orc_pyr_info;

static int pyramid_levels(int W, int H, int num_levels, double pyr_scale, orc_pyr_info* info) {
  int k; double scale = 1;
  for (k = 0; k < num_levels; ++k) {
    scale *= pyr_scale;
    if (W * scale < 32 || H * scale < 32) break;
  }
  int levels = k;
  for (k = 0; k <= levels; ++k) {
    double sc = 1;
    for (int i = 0; i < k; ++i) sc *= pyr_scale;
    info->w[k] = cv_round(W * sc);
    info->h[k] = cv_round(H * sc);
  }
  info->levels = levels;
  return levels;
}

ORC_API int orc_pyramid_info(int W, int H, int num_levels, double pyr_scale, int* ws, int* hs) {
  orc_pyr_info info;
  int l = pyramid_levels(W, H, num_levels, pyr_scale, &info);
  for (int k = 0; k <= l; ++k) { ws[k] = info.w[k]; hs[k] = info.h[k]; }
  return l;
}

/* Stage dumps for stage-by-stage diffing of the CUDA kernels.  dump_level < 0: none.
 * When dump_level == k the buffers (if non-NULL) receive level k's I (prev image),
 * R0, R1 (h x w x 5 interleaved) and M after the initial UpdateMatrices. */
typedef struct {
  int level;
  float* I0; float* I1; float* R0; float* R1; float* M0; float* flow_out;
} orc_dump;

#define ORC_FARNEBACK_GAUSSIAN 256 /* cv::OPTFLOW_FARNEBACK_GAUSSIAN */

ORC_API void orc_farneback_flags(const uint8_t* gray0, const uint8_t* gray1, int W, int H, float* flow_out,
                                 int num_levels, double pyr_scale, int winsize, int iters, int poly_n,
                                 double poly_sigma, int flags, orc_dump* dump) {
  orc_pyr_info info;
  int levels = pyramid_levels(W, H, num_levels, pyr_scale, &info);
  poly_consts pc;
  prepare_gaussian(poly_n, poly_sigma, &pc);
  const uint8_t* img[2] = {gray0, gray1};
  size_t N = (size_t)W * H;
  float* fimg = (float*)malloc(sizeof(float) * N);
  float* blur = (float*)malloc(sizeof(float) * N);
  float* prev_flow = NULL; int pw = 0, ph = 0;
  for (int k = levels; k >= 0; --k) {
    double scale = 1;
    for (int i = 0; i < k; ++i) scale *= pyr_scale;
    double sigma = (1. / scale - 1) * 0.5;
    int smooth = cv_round(sigma * 5) | 1;
    if (smooth < 3) smooth = 3;
    int w = info.w[k], h = info.h[k];
    size_t n = (size_t)w * h;
    float* flow = (float*)malloc(sizeof(float) * n * 2);
    if (!prev_flow) memset(flow, 0, sizeof(float) * n * 2);
    else {
      resize_linear(prev_flow, pw, ph, flow, w, h, 2);
      float mul = (float)(1. / pyr_scale);
      for (size_t i = 0; i < n * 2; ++i) flow[i] *= mul;
    }
    float* R[2]; float* I = (float*)malloc(sizeof(float) * n);
    for (int i = 0; i < 2; ++i) {
      for (size_t p = 0; p < N; ++p) fimg[p] = (float)img[i][p];
      gaussian_blur(fimg, blur, W, H, smooth, sigma);
      resize_linear(blur, W, H, I, w, h, 1);
      R[i] = (float*)malloc(sizeof(float) * n * 5);
      poly_exp(I, R[i], w, h, poly_n, &pc);
      if (dump && dump->level == k) {
        float* dI = i == 0 ? dump->I0 : dump->I1;
        if (dI) memcpy(dI, I, sizeof(float) * n);
        float* dR = i == 0 ? dump->R0 : dump->R1;
        if (dR) memcpy(dR, R[i], sizeof(float) * n * 5);
      }
    }
    float* M = (float*)malloc(sizeof(float) * n * 5);
    update_matrices(R[0], R[1], flow, M, w, h);
    if (dump && dump->level == k && dump->M0) memcpy(dump->M0, M, sizeof(float) * n * 5);
    for (int i = 0; i < iters; ++i) {
      if (flags & ORC_FARNEBACK_GAUSSIAN) update_flow_gaussian(R[0], R[1], flow, M, w, h, winsize, i < iters - 1);
      else update_flow_blur(R[0], R[1], flow, M, w, h, winsize, i < iters - 1);
    }
    if (dump && dump->level == k && dump->flow_out) memcpy(dump->flow_out, flow, sizeof(float) * n * 2);
    free(M); free(R[0]); free(R[1]); free(I);
    free(prev_flow);
    prev_flow = flow; pw = w; ph = h;
  }
  memcpy(flow_out, prev_flow, sizeof(float) * N * 2);
  free(prev_flow); free(fimg); free(blur);
}

ORC_API void orc_farneback_ex(const uint8_t* gray0, const uint8_t* gray1, int W, int H, float* flow_out,
                              int num_levels, double pyr_scale, int winsize, int iters, int poly_n,
                              double poly_sigma, orc_dump* dump) {
  orc_farneback_flags(gray0, gray1, W, H, flow_out, num_levels, pyr_scale, winsize, iters, poly_n, poly_sigma, 0, dump);
}

/* the reference's fixed parameters: optical_flow_kernel_cpu.cpp:16 */
ORC_API void orc_farneback(const uint8_t* gray0, const uint8_t* gray1, int W, int H, float* flow_out) {
  orc_farneback_ex(gray0, gray1, W, H, flow_out, 3, 0.5, 15, 3, 5, 1.2, NULL);
}

/* OpticalFlow op end to end: optical_flow_kernel_cpu.cpp:27-43 */
ORC_API void orc_optical_flow_rgb(const uint8_t* rgb0, const uint8_t* rgb1, int W, int H, float* flow_out) {
  size_t N = (size_t)W * H;
  uint8_t* g0 = (uint8_t*)malloc(N); uint8_t* g1 = (uint8_t*)malloc(N);
  orc_gray(rgb0, (int)N, g0); orc_gray(rgb1, (int)N, g1);
  orc_farneback(g0, g1, W, H, flow_out);
  free(g0); free(g1);
}

/* ---- Resize (next row): cv::resize INTER_LINEAR on 8-bit frames ---------
 * scannertools_cpp/imgproc/resize_kernel.cpp:69-71.  OpenCV's 8U path is fixed point:
 * 11-bit coefficients (cvRound(f * 2048)), horizontal pass to int, vertical pass
 * (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  x source index/fraction are
 * clamped when the coefficients are built; y keeps the unclamped fraction and clips the two row
 * indices.  Exact 2x down-scaling in both directions takes the INTER_AREA fast path
 * (a + b + c + d + 2) >> 2.  Verified bit-exact against cv2 4.13 (tests/test_oracle.py). */
ORC_API void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh) {
  if (sw == 2 * dw && sh == 2 * dh) {
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          const uint8_t* p = src + ((size_t)(2 * y) * sw + 2 * x) * cn + c;
          dst[((size_t)y * dw + x) * cn + c] = (uint8_t)((p[0] + p[cn] + p[(size_t)sw * cn] + p[(size_t)sw * cn + cn] + 2) >> 2);
        }
    return;
  }
  double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
  int* xofs = (int*)malloc(sizeof(int) * dw);
  int* xa = (int*)malloc(sizeof(int) * dw * 2);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = cv_floor_f(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    xa[dx * 2] = cv_round((double)((1.f - fx) * 2048.f));
    xa[dx * 2 + 1] = cv_round((double)(fx * 2048.f));
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = cv_floor_f(fy);
    fy -= sy;
    int b0 = cv_round((double)((1.f - fy) * 2048.f)), b1 = cv_round((double)(fy * 2048.f));
    int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
    int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
    const uint8_t* r0 = src + (size_t)y0 * sw * cn;
    const uint8_t* r1 = src + (size_t)y1 * sw * cn;
    for (int dx = 0; dx < dw; ++dx) {
      int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sx;
      int a0 = xa[dx * 2], a1 = xa[dx * 2 + 1];
      for (int c = 0; c < cn; ++c) {
        int h0 = r0[sx * cn + c] * a0 + r0[sx1 * cn + c] * a1;
        int h1 = r1[sx * cn + c] * a0 + r1[sx1 * cn + c] * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        dst[((size_t)dy * dw + dx) * cn + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
    }
  }
  free(xofs); free(xa);
}

/* ---- Resize with the other interpolation names the reference maps (resize_kernel.cpp:9-20):
 * INTER_NEAREST and INTER_AREA on 8-bit frames, as OpenCV 4.x executes them (imgproc/src/resize.cpp).
 *   NEAREST: sx = min(floor(dx * (1 / inv_scale_x)), sw - 1), same for y.
 *   AREA, both scales >= 1: integer factors -> block sums, (a+b+c+d+2)>>2 for 2x2, otherwise
 *         saturate(round(sum * (1.f / area))); else the general table of (index, float weight)
 *         per axis (computeResizeAreaTab) with float accumulation in table order:
 *         buf = sum_x S*alpha; sum = beta0*buf0, then sum += beta*buf.
 *   AREA with an up-scaled axis: INTER_LINEAR arithmetic with the "area" coefficient
 *         fx = (dx+1) - (sx+1)*inv_scale, sx = floor(dx*scale).
 * Verified bit-exact against cv2 4.13 (tests/test_oracle.py). */
enum { ORC_INTER_LINEAR = 0, ORC_INTER_NEAREST = 1, ORC_INTER_AREA = 2, ORC_INTER_CUBIC = 3, ORC_INTER_LANCZOS4 = 4 };

/* ---- INTER_CUBIC / INTER_LANCZOS4 on 8-bit frames (resize_kernel.cpp:13,15), OpenCV's own generic
 * path (imgproc/src/resize.cpp resizeGeneric_, HResizeCubic / HResizeLanczos4 <uchar,int,short>):
 *   per destination index: f = (float)((d + 0.5) * scale - 0.5), s = floor(f), f -= s, K float taps
 *   (interpolateCubic, A = -0.75 / interpolateLanczos4), shorts saturate_cast<short>(tap * 2048);
 *   tap k reads source index clamp(s - (K/2 - 1) + k); horizontal pass in int; vertical pass:
 *   cubic -- VResizeCubicVec_32s8u for the first floor(W*cn/8)*8 elements of a row (128-bit universal
 *   intrinsics: float, beta * 2^-22, S3*b3 + S2*b2 + S1*b1 + S0*b0 accumulated in that order with
 *   separately rounded mul / add, round half even), (sum + 2^21) >> 22 in int for the scalar tail;
 *   Lanczos4 -- int only.  Verified bit-exact against cv2 4.13 with cv2.ipp.setUseIPP(False): this
 *   wheel otherwise dispatches 8-bit INTER_CUBIC to IPP, whose output differs from OpenCV's own code
 *   by one grey level on ~5 % of pixels (tests/test_oracle.py bounds that too). */
static uint8_t sat_u8_f(float v) { long r = lrintf(v); return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r)); }

static void orc_cubic_coeffs(float x, float* c) {
  const float A = -0.75f;
  c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
  c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
  c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
  c[3] = 1.f - c[0] - c[1] - c[2];
}

static void orc_lanczos4_coeffs(float x, float* c) {
  static const double s45 = 0.70710678118654752440;
  static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  if (x < 1.1920928955078125e-07f) { for (int i = 0; i < 8; ++i) c[i] = 0.f; c[3] = 1.f; return; }
  float sum = 0.f;
  double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = sin(y0), c0 = cos(y0);
  for (int i = 0; i < 8; ++i) {
    double y = -(x + 3 - i) * 3.1415926535897932384626433832795 * 0.25;
    c[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    sum += c[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; ++i) c[i] *= sum;
}

static void orc_taps(int dn, int sn, int K, int cubic, int* ofs, short* al) {
  double scale = 1. / ((double)dn / sn);
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s0 = (int)floor(f);
    f -= s0;
    float c[8];
    if (cubic) orc_cubic_coeffs(f, c); else orc_lanczos4_coeffs(f, c);
    ofs[d] = s0 - (K / 2 - 1);
    for (int k = 0; k < K; ++k) {
      long r = lrintf(c[k] * 2048.f);
      al[d * K + k] = (short)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
    }
  }
}

static void resize_taps_u8(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh, int cubic) {
  const int K = cubic ? 4 : 8, W = dw * cn;
  int* xofs = (int*)malloc(sizeof(int) * (size_t)dw); short* xa = (short*)malloc(sizeof(short) * (size_t)dw * K);
  int* yofs = (int*)malloc(sizeof(int) * (size_t)dh); short* ya = (short*)malloc(sizeof(short) * (size_t)dh * K);
  orc_taps(dw, sw, K, cubic, xofs, xa);
  orc_taps(dh, sh, K, cubic, yofs, ya);
  int* rows = (int*)malloc(sizeof(int) * (size_t)W * K);   /* the K horizontally resampled source rows of one output row */
  const int nvec = cubic ? (W / 8) * 8 : 0;
  for (int dy = 0; dy < dh; ++dy) {
    for (int j = 0; j < K; ++j) {
      int sy = yofs[dy] + j; sy = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
      const uint8_t* S = src + (size_t)sy * sw * cn;
      for (int dx = 0; dx < dw; ++dx)
        for (int c = 0; c < cn; ++c) {
          int h = 0;
          for (int k = 0; k < K; ++k) {
            int sx = xofs[dx] + k; sx = sx < 0 ? 0 : (sx > sw - 1 ? sw - 1 : sx);
            h += S[sx * cn + c] * xa[dx * K + k];
          }
          rows[(size_t)j * W + dx * cn + c] = h;
        }
    }
    const short* b = ya + (size_t)dy * K;
    uint8_t* D = dst + (size_t)dy * W;
    int x = 0;
    if (cubic) {
      const float sc = 1.f / (2048.f * 2048.f);
      const float b0 = b[0] * sc, b1 = b[1] * sc, b2 = b[2] * sc, b3 = b[3] * sc;
      for (; x < nvec; ++x) {
        volatile float t3 = (float)rows[3 * (size_t)W + x] * b3;
        volatile float t2 = (float)rows[2 * (size_t)W + x] * b2;
        volatile float t1 = (float)rows[1 * (size_t)W + x] * b1;
        volatile float t0 = (float)rows[x] * b0;
        volatile float acc = t2 + t3;
        acc = t1 + acc;
        acc = t0 + acc;
        D[x] = sat_u8_f(acc);
      }
    }
    for (; x < W; ++x) {
      int t = 0;
      for (int j = 0; j < K; ++j) t += rows[(size_t)j * W + x] * b[j];
      int v = (t + (1 << 21)) >> 22;
      D[x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
  }
  free(rows); free(xofs); free(xa); free(yofs); free(ya);
}

typedef struct { int di, si; float alpha; } area_tab_t;

static int area_tab(int ssize, int dsize, double scale, area_tab_t* tab) {
  int k = 0;
  for (int dx = 0; dx < dsize; ++dx) {
    double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cell); }
    for (int sx = sx1; sx < sx2; ++sx) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cell); }
    if (fsx2 - sx2 > 1e-3) {
      double a = fsx2 - sx2; if (a > 1.) a = 1.; if (a > cell) a = cell;
      tab[k].di = dx; tab[k].si = sx2; tab[k++].alpha = (float)(a / cell);
    }
  }
  return k;
}


static void resize_linear_area_mode(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh) {
  double inv_x = (double)dw / sw, inv_y = (double)dh / sh, scale_x = 1. / inv_x, scale_y = 1. / inv_y;
  for (int dy = 0; dy < dh; ++dy) {
    int sy = (int)floor(dy * scale_y);
    float fy = (float)((dy + 1) - (sy + 1) * inv_y);
    fy = fy <= 0 ? 0.f : fy - floorf(fy);
    int b0 = cv_round((double)((1.f - fy) * 2048.f)), b1 = cv_round((double)(fy * 2048.f));
    int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
    int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
    const uint8_t* r0 = src + (size_t)y0 * sw * cn;
    const uint8_t* r1 = src + (size_t)y1 * sw * cn;
    for (int dx = 0; dx < dw; ++dx) {
      int sx = (int)floor(dx * scale_x);
      float fx = (float)((dx + 1) - (sx + 1) * inv_x);
      fx = fx <= 0 ? 0.f : fx - floorf(fx);
      if (sx < 0) { fx = 0; sx = 0; }
      if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
      int sx1 = sx + 1 < sw ? sx + 1 : sx;
      int a0 = cv_round((double)((1.f - fx) * 2048.f)), a1 = cv_round((double)(fx * 2048.f));
      for (int c = 0; c < cn; ++c) {
        int h0 = r0[sx * cn + c] * a0 + r0[sx1 * cn + c] * a1;
        int h1 = r1[sx * cn + c] * a0 + r1[sx1 * cn + c] * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        dst[((size_t)dy * dw + dx) * cn + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
    }
  }
}

ORC_API int orc_resize_u8(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh, int interp) {
  if (interp == ORC_INTER_LINEAR) { orc_resize_linear_u8(src, sw, sh, cn, dst, dw, dh); return 0; }
  double inv_x = (double)dw / sw, inv_y = (double)dh / sh, scale_x = 1. / inv_x, scale_y = 1. / inv_y;
  if (interp == ORC_INTER_NEAREST) {
    for (int y = 0; y < dh; ++y) {
      int sy = (int)floor(y * scale_y); if (sy > sh - 1) sy = sh - 1;
      for (int x = 0; x < dw; ++x) {
        int sx = (int)floor(x * scale_x); if (sx > sw - 1) sx = sw - 1;
        memcpy(dst + ((size_t)y * dw + x) * cn, src + ((size_t)sy * sw + sx) * cn, (size_t)cn);
      }
    }
    return 0;
  }
  if (interp == ORC_INTER_CUBIC || interp == ORC_INTER_LANCZOS4) {
    resize_taps_u8(src, sw, sh, cn, dst, dw, dh, interp == ORC_INTER_CUBIC);
    return 0;
  }
  if (interp != ORC_INTER_AREA) return -1;
  if (!(scale_x >= 1 && scale_y >= 1)) { resize_linear_area_mode(src, sw, sh, cn, dst, dw, dh); return 0; }
  int isx = cv_round(scale_x), isy = cv_round(scale_y);
  if (fabs(scale_x - isx) < 2.220446049250313e-16 && fabs(scale_y - isy) < 2.220446049250313e-16) {
    float scale = 1.f / (isx * isy);
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          int sum = 0;
          for (int j = 0; j < isy; ++j)
            for (int i = 0; i < isx; ++i) sum += src[((size_t)(y * isy + j) * sw + (x * isx + i)) * cn + c];
          dst[((size_t)y * dw + x) * cn + c] = (isx == 2 && isy == 2) ? (uint8_t)((sum + 2) >> 2) : sat_u8_f(sum * scale);
        }
    return 0;
  }
  area_tab_t* xt = (area_tab_t*)malloc(sizeof(area_tab_t) * ((size_t)sw * 2 + dw * 2));
  area_tab_t* yt = (area_tab_t*)malloc(sizeof(area_tab_t) * ((size_t)sh * 2 + dh * 2));
  int nx = area_tab(sw, dw, scale_x, xt), ny = area_tab(sh, dh, scale_y, yt);
  float* buf = (float*)malloc(sizeof(float) * (size_t)dw * cn);
  float* sum = (float*)malloc(sizeof(float) * (size_t)dw * cn);
  int prev = -1;
  for (int j = 0; j < ny; ++j) {
    const uint8_t* S = src + (size_t)yt[j].si * sw * cn;
    for (int i = 0; i < dw * cn; ++i) buf[i] = 0.f;
    for (int k = 0; k < nx; ++k)
      for (int c = 0; c < cn; ++c) buf[xt[k].di * cn + c] = buf[xt[k].di * cn + c] + S[xt[k].si * cn + c] * xt[k].alpha;
    if (yt[j].di != prev) {
      if (prev >= 0) for (int i = 0; i < dw * cn; ++i) dst[(size_t)prev * dw * cn + i] = sat_u8_f(sum[i]);
      for (int i = 0; i < dw * cn; ++i) sum[i] = yt[j].alpha * buf[i];
      prev = yt[j].di;
    } else {
      for (int i = 0; i < dw * cn; ++i) sum[i] = sum[i] + yt[j].alpha * buf[i];
    }
  }
  if (prev >= 0) for (int i = 0; i < dw * cn; ++i) dst[(size_t)prev * dw * cn + i] = sat_u8_f(sum[i]);
  free(xt); free(yt); free(buf); free(sum);
  return 0;
}

/* ---- ConvertColor (next row): cv::cvtColor RGB2HSV on 8-bit frames -------------
 * old/cpp_ops/imgproc.cpp:41.  OpenCV's integer path (hsv_shift = 12, H range 180). */
ORC_API void orc_rgb2hsv_u8(const uint8_t* rgb, size_t n_px, uint8_t* hsv) {
  static int sdiv[256], hdiv[256], init = 0;
  if (!init) {
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; ++i) {
      sdiv[i] = cv_round((255 << 12) / (1. * i));
      hdiv[i] = cv_round((180 << 12) / (6. * i));
    }
    init = 1;
  }
  for (size_t i = 0; i < n_px; ++i) {
    int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    int v = b > g ? b : g; if (r > v) v = r;
    int vmin = b < g ? b : g; if (r < vmin) vmin = r;
    int diff = v - vmin;
    int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    int s = (diff * sdiv[v] + (1 << 11)) >> 12;
    int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * hdiv[diff] + (1 << 11)) >> 12;
    h += h < 0 ? 180 : 0;
    hsv[3 * i] = (uint8_t)(h < 0 ? 0 : (h > 255 ? 255 : h));
    hsv[3 * i + 1] = (uint8_t)s;
    hsv[3 * i + 2] = (uint8_t)v;
  }
}
