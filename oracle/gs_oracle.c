/*
 * gs_oracle.c -- CPU restatement of GSASR's 2-D Gaussian rasteriser.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product path (gsasr_b200/) never does.
 *
 * What it restates (paths relative to the GSASR tree):
 *   gso_forward   utils/gs_cuda_dmax/gs.cu:24-62   (forward, dmax window)
 *                 utils/gs_cuda/gs.cu:26-58        (forward without window == dmax = +inf)
 *   gso_backward  utils/gs_cuda_dmax/gs.cu:97-162  (backward, dmax window)
 *                 utils/gs_cuda/gs.cu:98-176       (backward without window)
 *   brute-force formula + mask cross-check: utils/gs_cuda_dmax/check.py:13-29
 *
 * Faithful where it decides WHAT is summed, exact where it decides HOW MUCH:
 *   - pixel coordinates follow the reference to the bit: (float)(2.0*i/(n-1) - 1.0) in double,
 *     rounded once to float (gs.cu:39,46);
 *   - d_x, d_y are fp32 subtractions and the window test is the reference's
 *     `d > dmax || d < -dmax -> skip` on those fp32 values (gs.cu:40-50,124-132);
 *   - mode 0 ("exact"): everything after d_x, d_y is evaluated in double and accumulated in
 *     double -- the value the reference approximates, free of its fp32 rounding and of the
 *     run-to-run noise of its atomics;
 *   - mode 1 ("fp32"): the per-pixel value is computed with the reference's fp32 operation
 *     order (compiled with -ffp-contract=off; nvcc's FMA contraction is not reproduced) and
 *     accumulated in double.
 * Loop order differs from the reference (rows are split across OpenMP threads; the per-axis
 * inclusion flags are evaluated once per Gaussian instead of once per pixel) but the set of
 * (Gaussian, pixel) pairs and every per-pair value are the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GSO_API __attribute__((visibility("default")))

static float pix_coord(int i, int n) { return (float)(2.0 * i / (n - 1) - 1.0); }

static int in_window(float coord, float ctr, float dmax) {
  volatile float d = coord - ctr; /* fp32 subtraction, as on the GPU */
  return !(d > dmax || d < -dmax);
}

/* First/last index passing the window test, by linear scan (no monotonicity assumed for the
 * bounds search: we scan every index; *contig reports whether the passing set is contiguous). */
static void axis_range(const float* tab, int n, float ctr, float dmax, int* lo, int* hi, int* contig) {
  int l = n, h = -1, cnt = 0;
  for (int i = 0; i < n; ++i) {
    if (in_window(tab[i], ctr, dmax)) {
      if (i < l) l = i;
      h = i;
      ++cnt;
    }
  }
  *lo = l;
  *hi = h;
  *contig = (h < l) || (cnt == h - l + 1);
}

GSO_API int gso_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

GSO_API void gso_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* Per-Gaussian inclusion ranges; ranges[4*g..] = x0,x1,y0,y1 (inclusive; empty if x1<x0).
 * Returns 0 if every inclusion set was contiguous (it always is), 1 otherwise. */
GSO_API int gso_ranges(const float* coords, int s, int h, int w, float dmax, int* ranges) {
  float* px = (float*)malloc(sizeof(float) * (size_t)w);
  float* py = (float*)malloc(sizeof(float) * (size_t)h);
  for (int i = 0; i < w; ++i) px[i] = pix_coord(i, w);
  for (int i = 0; i < h; ++i) py[i] = pix_coord(i, h);
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int g = 0; g < s; ++g) {
    int c1, c2;
    axis_range(px, w, coords[2 * g + 0], dmax, &ranges[4 * g + 0], &ranges[4 * g + 1], &c1);
    axis_range(py, h, coords[2 * g + 1], dmax, &ranges[4 * g + 2], &ranges[4 * g + 3], &c2);
    bad |= !(c1 && c2);
  }
  free(px);
  free(py);
  return bad;
}

static double pair_value(float sx, float sy, float rho, float dx, float dy, int mode) {
  if (mode == 1) {
    /* gs.cu:33-36,52-56 in fp32 operation order */
    float w1 = (float)(-0.5 / (1 - rho * rho));
    float w2 = (float)(1.0 / sx / sx);
    float w4 = (float)(1.0 / sy / sy);
    float w3 = (float)(2 * rho / sx / sy);
    float v = w2 * dx * dx;
    v -= w3 * dx * dy;
    v += w4 * dy * dy;
    v *= w1;
    return (double)expf(v);
  }
  const double r = rho, a = sx, b = sy, x = dx, y = dy;
  const double q = x * x / (a * a) - 2.0 * r * x * y / (a * b) + y * y / (b * b);
  return exp(-0.5 * q / (1.0 - r * r));
}

/* img (h,w,3) double, ACCUMULATED into (the reference accumulates into rendered_img).
 * count (h,w) int32 or NULL: number of Gaussians whose window includes the pixel. */
GSO_API void gso_forward(const float* sigmas, const float* coords, const float* colors, double* img,
                         int32_t* count, int s, int h, int w, float dmax, int mode) {
  int* ranges = (int*)malloc(sizeof(int) * 4 * (size_t)(s > 0 ? s : 1));
  gso_ranges(coords, s, h, w, dmax, ranges);
  float* px = (float*)malloc(sizeof(float) * (size_t)w);
  float* py = (float*)malloc(sizeof(float) * (size_t)h);
  for (int i = 0; i < w; ++i) px[i] = pix_coord(i, w);
  for (int i = 0; i < h; ++i) py[i] = pix_coord(i, h);
  /* rows are owned by threads: no two threads ever touch the same pixel */
#pragma omp parallel
  {
#ifdef _OPENMP
    const int nt = omp_get_num_threads(), t = omp_get_thread_num();
#else
    const int nt = 1, t = 0;
#endif
    const int r0 = (int)((long long)h * t / nt), r1 = (int)((long long)h * (t + 1) / nt) - 1;
    for (int g = 0; g < s; ++g) {
      const int x0 = ranges[4 * g], x1 = ranges[4 * g + 1];
      int y0 = ranges[4 * g + 2], y1 = ranges[4 * g + 3];
      if (x1 < x0 || y1 < y0) continue;
      if (y0 < r0) y0 = r0;
      if (y1 > r1) y1 = r1;
      const float sx = sigmas[3 * g], sy = sigmas[3 * g + 1], rho = sigmas[3 * g + 2];
      const float cx = coords[2 * g], cy = coords[2 * g + 1];
      const double cr = colors[3 * g], cg = colors[3 * g + 1], cb = colors[3 * g + 2];
      for (int hi = y0; hi <= y1; ++hi) {
        volatile float dyv = py[hi] - cy;
        const float dy = dyv;
        for (int wi = x0; wi <= x1; ++wi) {
          volatile float dxv = px[wi] - cx;
          const float dx = dxv;
          const double v = pair_value(sx, sy, rho, dx, dy, mode);
          double* o = img + ((size_t)hi * w + wi) * 3;
          o[0] += v * cr;
          o[1] += v * cg;
          o[2] += v * cb;
          if (count) count[(size_t)hi * w + wi] += 1;
        }
      }
    }
  }
  free(ranges);
  free(px);
  free(py);
}

/* Same sum restricted to the pixel rectangle rows [cy0, cy0+ch) x columns [cx0, cx0+cw) of the h x w image:
 * crop (ch,cw,3) double, ACCUMULATED into.  Every Gaussian whose window reaches the rectangle contributes
 * exactly the pairs gso_forward would have summed there (used by the full-size parity tests, where the
 * whole image is out of the CPU's reach). */
GSO_API void gso_forward_crop(const float* sigmas, const float* coords, const float* colors, double* crop,
                              int s, int h, int w, float dmax, int mode, int cy0, int cx0, int ch, int cw) {
  int* ranges = (int*)malloc(sizeof(int) * 4 * (size_t)(s > 0 ? s : 1));
  gso_ranges(coords, s, h, w, dmax, ranges);
  float* px = (float*)malloc(sizeof(float) * (size_t)w);
  float* py = (float*)malloc(sizeof(float) * (size_t)h);
  for (int i = 0; i < w; ++i) px[i] = pix_coord(i, w);
  for (int i = 0; i < h; ++i) py[i] = pix_coord(i, h);
#pragma omp parallel
  {
#ifdef _OPENMP
    const int nt = omp_get_num_threads(), t = omp_get_thread_num();
#else
    const int nt = 1, t = 0;
#endif
    const int r0 = cy0 + (int)((long long)ch * t / nt), r1 = cy0 + (int)((long long)ch * (t + 1) / nt) - 1;
    for (int g = 0; g < s; ++g) {
      int x0 = ranges[4 * g], x1 = ranges[4 * g + 1];
      int y0 = ranges[4 * g + 2], y1 = ranges[4 * g + 3];
      if (x1 < x0 || y1 < y0) continue;
      if (y0 < r0) y0 = r0;
      if (y1 > r1) y1 = r1;
      if (x0 < cx0) x0 = cx0;
      if (x1 > cx0 + cw - 1) x1 = cx0 + cw - 1;
      if (x1 < x0 || y1 < y0) continue;
      const float sx = sigmas[3 * g], sy = sigmas[3 * g + 1], rho = sigmas[3 * g + 2];
      const float cx = coords[2 * g], cy = coords[2 * g + 1];
      const double cr = colors[3 * g], cg = colors[3 * g + 1], cb = colors[3 * g + 2];
      for (int hi = y0; hi <= y1; ++hi) {
        volatile float dyv = py[hi] - cy;
        const float dy = dyv;
        for (int wi = x0; wi <= x1; ++wi) {
          volatile float dxv = px[wi] - cx;
          const float dx = dxv;
          const double v = pair_value(sx, sy, rho, dx, dy, mode);
          double* o = crop + ((size_t)(hi - cy0) * cw + (wi - cx0)) * 3;
          o[0] += v * cr;
          o[1] += v * cg;
          o[2] += v * cb;
        }
      }
    }
  }
  free(ranges);
  free(px);
  free(py);
}

/* grads (h,w,3) float; outputs double, ACCUMULATED into: g_sigmas (s,3), g_coords (s,2),
 * g_colors (s,3).  Formulas: gs.cu:134-159. */
GSO_API void gso_backward(const float* sigmas, const float* coords, const float* colors,
                          const float* grads, double* g_sigmas, double* g_coords, double* g_colors,
                          int s, int h, int w, float dmax) {
  int* ranges = (int*)malloc(sizeof(int) * 4 * (size_t)(s > 0 ? s : 1));
  gso_ranges(coords, s, h, w, dmax, ranges);
  float* px = (float*)malloc(sizeof(float) * (size_t)w);
  float* py = (float*)malloc(sizeof(float) * (size_t)h);
  for (int i = 0; i < w; ++i) px[i] = pix_coord(i, w);
  for (int i = 0; i < h; ++i) py[i] = pix_coord(i, h);
#pragma omp parallel for schedule(dynamic, 64)
  for (int g = 0; g < s; ++g) {
    const int x0 = ranges[4 * g], x1 = ranges[4 * g + 1];
    const int y0 = ranges[4 * g + 2], y1 = ranges[4 * g + 3];
    if (x1 < x0 || y1 < y0) continue;
    const double sx = sigmas[3 * g], sy = sigmas[3 * g + 1], rho = sigmas[3 * g + 2];
    const float cx = coords[2 * g], cy = coords[2 * g + 1];
    const double col[3] = {colors[3 * g], colors[3 * g + 1], colors[3 * g + 2]};
    const double w1 = -0.5 / (1.0 - rho * rho);
    const double w2 = 1.0 / (sx * sx), w3 = 1.0 / (sx * sy), w4 = 1.0 / (sy * sy);
    const double od_sx = 1.0 / sx, od_sy = 1.0 / sy;
    double gsx = 0, gsy = 0, gr = 0, gx = 0, gy = 0, gc[3] = {0, 0, 0};
    for (int hi = y0; hi <= y1; ++hi) {
      volatile float dyv = py[hi] - cy;
      const double dy = dyv;
      for (int wi = x0; wi <= x1; ++wi) {
        volatile float dxv = px[wi] - cx;
        const double dx = dxv;
        const double d = w2 * dx * dx - 2 * rho * w3 * dx * dy + w4 * dy * dy;
        const double v = exp(w1 * d);
        const double v2w1 = v * 2 * w1;
        const double to_x = v2w1 * (-w2 * dx + rho * w3 * dy);
        const double to_y = v2w1 * (-w4 * dy + rho * w3 * dx);
        const double to_sx = v2w1 * od_sx * (w3 * rho * dx * dy - w2 * dx * dx);
        const double to_sy = v2w1 * od_sy * (w3 * rho * dx * dy - w4 * dy * dy);
        const double to_r = -v2w1 * (2 * w1 * rho * d + w3 * dx * dy);
        const float* gp = grads + ((size_t)hi * w + wi) * 3;
        for (int c = 0; c < 3; ++c) {
          const double gptc = gp[c];
          const double gpt = gptc * col[c];
          gc[c] += v * gptc;
          gx += gpt * to_x;
          gy += gpt * to_y;
          gsx += gpt * to_sx;
          gsy += gpt * to_sy;
          gr += gpt * to_r;
        }
      }
    }
    g_sigmas[3 * g + 0] += gsx;
    g_sigmas[3 * g + 1] += gsy;
    g_sigmas[3 * g + 2] += gr;
    g_coords[2 * g + 0] += gx;
    g_coords[2 * g + 1] += gy;
    for (int c = 0; c < 3; ++c) g_colors[3 * g + c] += gc[c];
  }
  free(ranges);
  free(px);
  free(py);
}
