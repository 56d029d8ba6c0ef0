// gsr_common.cuh -- host/device math shared by every kernel of the rasteriser.
//
// Everything that decides WHICH pixels a Gaussian touches lives here as
// __host__ __device__ functions so that the CPU test hooks (gsr_hostcheck.cu) can
// run the very same code against the oracle without a GPU.
//
// Reference semantics restated here (GSASR tree):
//   pixel coordinate rule + inclusive dmax test   utils/gs_cuda_dmax/gs.cu:39-50, 124-132
//   conic coefficients                            utils/gs_cuda_dmax/gs.cu:33-36, 106-111
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define GSR_HD __host__ __device__ __forceinline__

// ---- geometry of the decomposition ---------------------------------------------------
// (overridable at compile time for tuning sweeps: -DGSR_CFG_TILE_W=.. etc.)
#ifndef GSR_CFG_TILE_W
#define GSR_CFG_TILE_W 32
#endif
#ifndef GSR_CFG_TILE_H
#define GSR_CFG_TILE_H 16
#endif
#ifndef GSR_CFG_BIN
#define GSR_CFG_BIN 8
#endif
#ifndef GSR_CFG_LARGE_PX
#define GSR_CFG_LARGE_PX 96
#endif
constexpr int GSR_TILE_W = GSR_CFG_TILE_W;  // forward CTA tile, pixels
constexpr int GSR_TILE_H = GSR_CFG_TILE_H;
constexpr int GSR_BIN = GSR_CFG_BIN;        // home-bin side (Gaussians are counting-sorted by bin)
constexpr int GSR_REGION = 8;               // one warp of the forward kernel owns an 8x8 region
constexpr int GSR_NRX = GSR_TILE_W / GSR_REGION, GSR_NRY = GSR_TILE_H / GSR_REGION;
constexpr int GSR_LARGE_PX = GSR_CFG_LARGE_PX;  // half-extent above which a Gaussian goes to the "large" list
static_assert(GSR_NRX * GSR_NRY <= 16, "region mask is 16 bits");
constexpr int GSR_MAX_DIM = 32767;    // bbox corners are stored as int16
// Region buckets of the forward fast path: one warp of gsr_forward_region_kernel rasterises a
// GSR_RGW x GSR_RGH pixel region (a 2x2 pixel block per lane); four lanes -- a 4x4-pixel CELL -- share one
// bit of the 8-bit cell mask every bucket entry carries (bit = cell row * 4 + cell column).
#ifndef GSR_CFG_MASK_PER_BAND
#define GSR_CFG_MASK_PER_BAND 0
#endif
constexpr int GSR_RGW = 16, GSR_RGH = 8, GSR_CELL = 4;
constexpr int GSR_CELLS_X = GSR_RGW / GSR_CELL, GSR_CELLS_Y = GSR_RGH / GSR_CELL;
static_assert(GSR_CELLS_X * GSR_CELLS_Y == 8 && GSR_CELLS_Y == 2, "entry = index | cell mask (8 bit) << 23 | binds << 31");
constexpr int GSR_ENT_MASK_SHIFT = 23;
constexpr uint32_t GSR_ENT_INDEX = (1u << GSR_ENT_MASK_SHIFT) - 1u;
constexpr int GSR_BUCKET_MAX_S = 1 << GSR_ENT_MASK_SHIFT;  // more Gaussians per call than this take the home-bin path
constexpr float GSR_LOG2E = 1.4426950408889634f;
constexpr float GSR_CULL_PAD_PX = 0.02f;  // slack on every culling bound (pixels)

// Sorted per-Gaussian record consumed by the raster kernels (32 B, two float4 loads).
//   E(dx,dy) = a*dx*dx + b*dx*dy + c*dy*dy   (log2 units, dx/dy in normalised [-1,1] units)
//   value    = exp2(E);  contribution = value * (r,g,bl)
struct __align__(16) GsrRec {
  float x, y, a, b;
  float c, r, g, bl;
};

// Cull box: inclusive integer pixel ranges packed as 4 x 15 bit in a uint2.  Bit 15 of the
// first half-word carries the "window binds" flag: the dmax window cuts the k-sigma box on at
// least one side, so pixels outside the box must be excluded EXACTLY (they may carry large
// values) -- the predicated path of the kernels.
GSR_HD uint2 gsr_box_pack(int x0, int x1, int y0, int y1, bool binds) {
  // x0,x1,y0,y1 in [0,32767]; the flag goes to the sign bit of the first half-word.
  uint32_t lo = (uint32_t)(x0 & 0x7fff) | (binds ? 0x8000u : 0u) | ((uint32_t)(x1 & 0x7fff) << 16);
  uint32_t hi = (uint32_t)(y0 & 0x7fff) | ((uint32_t)(y1 & 0x7fff) << 16);
  return make_uint2(lo, hi);
}
GSR_HD void gsr_box_unpack(uint2 p, int& x0, int& x1, int& y0, int& y1, bool& binds) {
  x0 = (int)(p.x & 0x7fffu);
  binds = (p.x & 0x8000u) != 0;
  x1 = (int)((p.x >> 16) & 0x7fffu);
  y0 = (int)(p.y & 0x7fffu);
  y1 = (int)((p.y >> 16) & 0x7fffu);
}

// ---- the reference's pixel coordinate rule (gs.cu:39,46): double arithmetic, one rounding
// to float on assignment.
GSR_HD float gsr_pix_coord(int i, int n) { return (float)(2.0 * i / (n - 1) - 1.0); }

// The reference's inclusion predicate for one axis (gs.cu:40-43,47-50): fp32 subtraction,
// skip iff d > dmax || d < -dmax (so NaN d is NOT skipped; we never get here with NaN centres).
// `tab` is the table of gsr_pix_coord values of the axis (device: filled by gsr_table_kernel);
// NULL evaluates the rule directly.
// A band view (rows [off, off+cnt) of the n-pixel axis) has a table of its own rows only: tab[i - off].
GSR_HD bool gsr_in_window(int i, int n, float ctr, float dmax, const float* tab = nullptr, int off = 0) {
  float d = (tab ? tab[i - off] : gsr_pix_coord(i, n)) - ctr;
  return !(d > dmax || d < -dmax);
}

// Inclusive index range [lo,hi] of the pixels of an n-pixel axis that pass gsr_in_window.
// The predicate is monotone in i, so the set is contiguous; we estimate both ends in double
// and repair them with the exact predicate (the estimate is off by at most one pixel).
// Empty range <=> lo > hi.
GSR_HD void gsr_window_range(int n, float ctr, float dmax, int& lo, int& hi,
                             const float* tab = nullptr, int off = 0, int cnt = -1) {
  // [off, off + cnt): the part of the axis that is rendered (a row band; the whole axis by default)
  if (cnt < 0) cnt = n;
  const int end = off + cnt;
  if (dmax != dmax || dmax >= 3.0e38f) {  // NaN never skips; +inf never skips
    lo = off;
    hi = end - 1;
    return;
  }
  if (dmax < 0.0f) {  // every finite d fails one of the two comparisons
    lo = 1;
    hi = 0;
    return;
  }
  const double s = 0.5 * (double)(n - 1);
  double flo = ((double)ctr - (double)dmax + 1.0) * s;
  double fhi = ((double)ctr + (double)dmax + 1.0) * s;
  flo = fmin(fmax(flo, (double)off - 2.0), (double)end + 1.0);
  fhi = fmin(fmax(fhi, (double)off - 2.0), (double)end + 1.0);
  int l = (int)ceil(flo) - 1;
  int h = (int)floor(fhi) + 1;
  l = l < off ? off : (l > end ? end : l);
  h = h > end - 1 ? end - 1 : (h < off - 1 ? off - 1 : h);
  for (int t = 0; t < 4 && l < end && !gsr_in_window(l, n, ctr, dmax, tab, off); ++t) ++l;
  if (l < end && !gsr_in_window(l, n, ctr, dmax, tab, off)) l = end;
  for (int t = 0; t < 4 && h >= off && !gsr_in_window(h, n, ctr, dmax, tab, off); ++t) --h;
  if (h >= off && !gsr_in_window(h, n, ctr, dmax, tab, off)) h = off - 1;
  lo = l;
  hi = h;
}

// ---- per-Gaussian set-up ------------------------------------------------------------------
struct GsrSetup {
  bool live;        // false: skipped entirely (invalid parameters or empty cull box)
  bool large;       // cull box half-extent above GSR_LARGE_PX: goes to the global "large" list
  bool binds;       // dmax window cuts the k-sigma box (exact predicate needed)
  int x0, x1, y0, y1;   // cull box, inclusive, clipped to the image
  int bin_x, bin_y;     // home bin (clamped into the image)
  int ext_x, ext_y;     // ceil of the distance from the centre to the far box edge (pixels)
};

GSR_HD bool gsr_finite(float v) { return fabsf(v) < 3.0e38f; }  // false for NaN

// Cull box = exact dmax window  INTERSECT  [c - (k*sigma_px + pad), c + (k*sigma_px + pad)].
// The image may be a ROW BAND of a taller one: rows [row0, row0 + h) of an hf-row image (hf = 0: the
// image is whole).  Pixel coordinates, windows and boxes are then those of the full image, cut to the
// band; band edges behave like image edges.  Everything returned is in band-local rows.
GSR_HD GsrSetup gsr_setup(float sx, float sy, float rho, float x, float y, float cr, float cg,
                          float cb, int h, int w, float dmax, float ksigma,
                          const float* px_tab = nullptr, const float* py_tab = nullptr, int hf = 0,
                          int row0 = 0) {
  if (hf <= 0) {
    hf = h;
    row0 = 0;
  }
  const int rend = row0 + h - 1;  // last row of the band, full-image coordinates
  GsrSetup o;
  o.live = false;
  o.large = false;
  o.binds = false;
  o.x0 = o.y0 = 1;
  o.x1 = o.y1 = 0;
  o.bin_x = o.bin_y = 0;
  o.ext_x = o.ext_y = 0;
  // (bitwise &: one predicate chain instead of eight branches)
  if (!(gsr_finite(sx) & gsr_finite(sy) & gsr_finite(rho) & gsr_finite(x) & gsr_finite(y) & gsr_finite(cr) &
        gsr_finite(cg) & gsr_finite(cb) & (sx != 0.0f) & (sy != 0.0f) & (fabsf(rho) < 1.0f)))
    return o;

  // Fast path (single precision): the k-sigma box lies at least three pixels inside the dmax window on every
  // side, so the window cannot bind and the box is a pure truncation bound -- pixels it drops carry less than
  // exp(-k^2/2) whichever way a boundary pixel falls (fp32 rounding of the centre, <= 1e-3 px at 32767, is
  // covered by the pad).  Everything else (window binds or nearly does, huge values) takes the double path.
  {
    const float hxf = 0.5f * (float)(w - 1), hyf = 0.5f * (float)(hf - 1);
    const float cxf = (x + 1.0f) * hxf, cyf = (y + 1.0f) * hyf;
    const float exf = ksigma * fabsf(sx) * hxf + (GSR_CULL_PAD_PX + 2.0e-6f * fabsf(cxf));
    const float eyf = ksigma * fabsf(sy) * hyf + (GSR_CULL_PAD_PX + 2.0e-6f * fabsf(cyf));
    const bool nowin = dmax != dmax || dmax >= 3.0e38f;
    const bool inside = nowin || (dmax >= 0.0f && exf + 3.0f < dmax * hxf && eyf + 3.0f < dmax * hyf);
    if (inside && fabsf(cxf) < 1.0e8f && fabsf(cyf) < 1.0e8f && exf < 1.0e8f && eyf < 1.0e8f) {
      const int kx0 = (int)ceilf(cxf - exf), kx1 = (int)floorf(cxf + exf);
      const int ky0 = (int)ceilf(cyf - eyf), ky1 = (int)floorf(cyf + eyf);
      o.x0 = kx0 > 0 ? kx0 : 0;
      o.x1 = kx1 < w - 1 ? kx1 : w - 1;
      o.y0 = (ky0 > row0 ? ky0 : row0) - row0;  // band-local
      o.y1 = (ky1 < rend ? ky1 : rend) - row0;
      if (o.x0 > o.x1 || o.y0 > o.y1) return o;
      o.live = true;
      const float cyl = cyf - (float)row0;
      const int nbx = (w + GSR_BIN - 1) / GSR_BIN, nby = (h + GSR_BIN - 1) / GSR_BIN;
      const float bx = floorf(cxf * (1.0f / GSR_BIN)), by = floorf(cyl * (1.0f / GSR_BIN));
      o.bin_x = (int)fminf(fmaxf(bx, 0.0f), (float)(nbx - 1));
      o.bin_y = (int)fminf(fmaxf(by, 0.0f), (float)(nby - 1));
      const float ccx = fminf(fmaxf(cxf, 0.0f), (float)(w - 1)), ccy = fminf(fmaxf(cyl, 0.0f), (float)(h - 1));
      const float dxm = fmaxf(ccx - (float)o.x0, (float)o.x1 - ccx);
      const float dym = fmaxf(ccy - (float)o.y0, (float)o.y1 - ccy);
      o.ext_x = (int)ceilf(fmaxf(dxm, 0.0f));
      o.ext_y = (int)ceilf(fmaxf(dym, 0.0f));
      o.large = (o.ext_x > GSR_LARGE_PX) || (o.ext_y > GSR_LARGE_PX);
      return o;
    }
  }
  const double hx = 0.5 * (double)(w - 1), hy = 0.5 * (double)(hf - 1);
  const double cx = ((double)x + 1.0) * hx, cy = ((double)y + 1.0) * hy;
  const double ex = (double)ksigma * fabs((double)sx) * hx + (double)GSR_CULL_PAD_PX;
  const double ey = (double)ksigma * fabs((double)sy) * hy + (double)GSR_CULL_PAD_PX;
  const double lim = 1.0e9;
  int kx0 = (int)ceil(fmin(fmax(cx - ex, -lim), lim));
  int kx1 = (int)floor(fmin(fmax(cx + ex, -lim), lim));
  int ky0 = (int)ceil(fmin(fmax(cy - ey, -lim), lim));
  int ky1 = (int)floor(fmin(fmax(cy + ey, -lim), lim));

  // Exact dmax window.  When the k-sigma box lies at least two pixels inside the window estimate
  // (the estimate is within one pixel of the truth) the window cannot bind and the exact repair
  // with the reference's predicate is skipped: the window is then only known to contain the box.
  int wx0, wx1, wy0, wy1;
  const double dmx = (dmax != dmax || dmax >= 3.0e38f) ? 1.0e9 : (double)dmax;
  if (dmx >= 0.0 && kx0 - 2.0 > (cx - dmx * hx) && kx1 + 2.0 < (cx + dmx * hx)) {
    wx0 = kx0 < 0 ? 0 : kx0;
    wx1 = kx1 > w - 1 ? w - 1 : kx1;
  } else {
    gsr_window_range(w, x, dmax, wx0, wx1, px_tab);
  }
  if (dmx >= 0.0 && ky0 - 2.0 > (cy - dmx * hy) && ky1 + 2.0 < (cy + dmx * hy)) {
    wy0 = ky0 < row0 ? row0 : ky0;
    wy1 = ky1 > rend ? rend : ky1;
  } else {
    gsr_window_range(hf, y, dmax, wy0, wy1, py_tab, row0, h);
  }
  if (wx0 > wx1 || wy0 > wy1) return o;
  // "binds": on some side the window is tighter than both the k-sigma box and the image edge.
  o.binds = (wx0 > (kx0 > 0 ? kx0 : 0)) || (wx1 < (kx1 < w - 1 ? kx1 : w - 1)) ||
            (wy0 > (ky0 > row0 ? ky0 : row0)) || (wy1 < (ky1 < rend ? ky1 : rend));
  o.x0 = wx0 > kx0 ? wx0 : kx0;
  o.x1 = wx1 < kx1 ? wx1 : kx1;
  o.y0 = (wy0 > ky0 ? wy0 : ky0) - row0;  // band-local from here on
  o.y1 = (wy1 < ky1 ? wy1 : ky1) - row0;
  if (o.x0 > o.x1 || o.y0 > o.y1) return o;
  o.live = true;

  const double cyl = cy - (double)row0;
  double bx = floor(cx / GSR_BIN), by = floor(cyl / GSR_BIN);
  const int nbx = (w + GSR_BIN - 1) / GSR_BIN, nby = (h + GSR_BIN - 1) / GSR_BIN;
  o.bin_x = (int)fmin(fmax(bx, 0.0), (double)(nbx - 1));
  o.bin_y = (int)fmin(fmax(by, 0.0), (double)(nby - 1));
  // Distance from the (clamped-into-image) centre to the far edges of the box: a tile that
  // overlaps the box lies within this distance of the home bin along each axis.
  const double ccx = fmin(fmax(cx, 0.0), (double)(w - 1)), ccy = fmin(fmax(cyl, 0.0), (double)(h - 1));
  double dxm = fmax(ccx - (double)o.x0, (double)o.x1 - ccx);
  double dym = fmax(ccy - (double)o.y0, (double)o.y1 - ccy);
  o.ext_x = (int)ceil(fmax(dxm, 0.0));
  o.ext_y = (int)ceil(fmax(dym, 0.0));
  o.large = (o.ext_x > GSR_LARGE_PX) || (o.ext_y > GSR_LARGE_PX);
  return o;
}

// Conic in log2 units (double set-up, rounded once to float):
//   E = log2(e) * w1 * (w2 dx^2 - 2 rho w3 dx dy + w4 dy^2),  w1 = -0.5/(1-rho^2), w2 = 1/sx^2 ...
GSR_HD GsrRec gsr_make_rec(float sx, float sy, float rho, float x, float y, float cr, float cg,
                           float cb) {
  // 1 - rho^2 cancels for |rho| -> 1 (the head emits up to 0.999999): formed in double, rounded once; the
  // rest is single precision (relative error of a, b, c <= 4e-7, i.e. <= 2e-7 absolute on a value).
  const float omr = (float)(1.0 - (double)rho * (double)rho);
  const float w1 = -0.5f * GSR_LOG2E / omr;
  const float isx = 1.0f / sx, isy = 1.0f / sy;
  GsrRec o;
  o.x = x;
  o.y = y;
  o.a = w1 * isx * isx;
  o.b = -2.0f * rho * w1 * isx * isy;
  o.c = w1 * isy * isy;
  o.r = cr;
  o.g = cg;
  o.bl = cb;
  return o;
}

// ---- region mask ----------------------------------------------------------------------------
// For a GSR_TILE_W x GSR_TILE_H tile at pixel origin (tx0,ty0), which of its GSR_NRX x GSR_NRY
// warp regions does the ellipse {E >= ecut}, clipped to the cull box, touch?  Bit
// (ry*GSR_NRX + rx).  Conservative (never misses a pixel with E >= ecut inside the
// box); pixels it drops carry exp2(E) < exp2(ecut).
//
// Per band of REGION rows: the ellipse's x-interval at row dy is centred on m(dy) = -b/(2a)*dy
// with half-width sqrt((ecut - c' dy^2)/a), c' = c - b^2/(4a); over a band we take the hull of
// the two end-row centres widened by the largest half-width in the band.
struct GsrEllipse {
  float cx, cy;     // centre, pixels
  float inv_a;      // 1 / (conic a in pixel units), < 0
  float kappa;      // ridge slope dx/dy
  float cp;         // c - b^2/(4a) in pixel units, <= 0
};

GSR_HD GsrEllipse gsr_ellipse(const GsrRec& g, int h, int w, int hf = 0, int row0 = 0) {
  const float hxs = 0.5f * (float)(w - 1), hys = 0.5f * (float)((hf > 0 ? hf : h) - 1);
  const float gx = 1.0f / hxs, gy = 1.0f / hys;  // normalised units per pixel
  GsrEllipse e;
  e.cx = (g.x + 1.0f) * hxs;
  e.cy = (g.y + 1.0f) * hys - (float)(hf > 0 ? row0 : 0);  // band-local rows
  const float a = g.a * gx * gx, b = g.b * gx * gy, c = g.c * gy * gy;  // conic in pixel units
  e.inv_a = 1.0f / a;
  e.kappa = -0.5f * b * e.inv_a;
  e.cp = c + 0.5f * e.kappa * b;
  // a degenerate conic (overflow, a = 0) must not cull: an ellipse as wide as the box, checked ONCE here so
  // that gsr_band_xrange needs no NaN handling
  if (!(gsr_finite(e.inv_a) & gsr_finite(e.kappa) & gsr_finite(e.cp) & gsr_finite(e.cx) & gsr_finite(e.cy) &
        (e.inv_a < 0.0f) & !(e.cp > 0.0f))) {
    e.inv_a = -3.0e37f;
    e.kappa = 0.0f;
    e.cp = 0.0f;
    e.cx = gsr_finite(e.cx) ? e.cx : 0.0f;
    e.cy = gsr_finite(e.cy) ? e.cy : 0.0f;
  }
  return e;
}

GSR_HD float gsr_sqrt_fast(float v) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));   // 1 MUFU; its 2^-22 relative error is inside the pad
  return r;
#else
  return sqrtf(v);
#endif
}

// Conservative pixel x-range [xl, xh] of {E >= ecut} over the rows ya..yb (inclusive, already
// clipped to the cull box), clipped to [cx0, cx1].  Returns false if the band is empty.
GSR_HD bool gsr_band_xrange(const GsrEllipse& e, float ecut, int ya, int yb, int cx0, int cx1,
                            int& xl, int& xh) {
  const float da = (float)ya - e.cy, db = (float)yb - e.cy;
  // row offset in the band closest to the centre row, shrunk by the pad so that rounding of cy
  // can only widen the band
  float t = da > 0.0f ? da : (db < 0.0f ? db : 0.0f);
  t = t > 0.0f ? fmaxf(t - GSR_CULL_PAD_PX, 0.0f) : fminf(t + GSR_CULL_PAD_PX, 0.0f);
  const float w2 = (ecut - e.cp * t * t) * e.inv_a;
  if (w2 < 0.0f) return false;  // band entirely outside the ellipse
  // fp32 slack on cx and on the approximate square root
  const float hw = gsr_sqrt_fast(w2) * 1.000001f + (GSR_CULL_PAD_PX + 1.0e-6f * fabsf(e.cx));
  const float ma = e.cx + e.kappa * da, mb = e.cx + e.kappa * db;
  const float lo = fmaxf(fminf(ma, mb) - hw, (float)cx0);
  const float hi = fminf(fmaxf(ma, mb) + hw, (float)cx1);
  xl = (int)ceilf(lo);
  xh = (int)floorf(hi);
  return xl <= xh;
}

GSR_HD uint32_t gsr_region_mask(const GsrRec& g, int bx0, int bx1, int by0, int by1, int tx0,
                                int ty0, int h, int w, float ecut, int hf = 0, int row0 = 0, int bhs = 0) {
  // uniform batch (bhs > 0): the Gaussian lives in the sample whose block of bhs rows holds its box
  GsrEllipse e = gsr_ellipse(g, bhs > 0 ? bhs : h, w, hf, row0);
  if (bhs > 0) e.cy += (float)((by0 / bhs) * bhs);
  uint32_t mask = 0;
  const int cx0 = bx0 > tx0 ? bx0 : tx0;
  const int cx1 = bx1 < tx0 + GSR_TILE_W - 1 ? bx1 : tx0 + GSR_TILE_W - 1;
  if (cx0 > cx1) return 0;
#pragma unroll
  for (int ry = 0; ry < GSR_NRY; ++ry) {
    int ya = ty0 + ry * GSR_REGION, yb = ya + GSR_REGION - 1;
    ya = ya > by0 ? ya : by0;
    yb = yb < by1 ? yb : by1;
    if (ya > yb) continue;
    int xl, xh;
    if (!gsr_band_xrange(e, ecut, ya, yb, cx0, cx1, xl, xh)) continue;
    const int r0 = (xl - tx0) / GSR_REGION, r1 = (xh - tx0) / GSR_REGION;
    const uint32_t bits = ((2u << r1) - 1u) & ~((1u << r0) - 1u);
    mask |= bits << (ry * GSR_NRX);
  }
  return mask;
}

// ---- cell masks of the region buckets -----------------------------------------------------------
// Cell-column ranges (4-pixel cells, inclusive; empty: cl > ch) of the two cell rows of region band `band`
// (rows band*8 .. band*8+7) that the ellipse {E >= ecut} touches inside the cull box [x0,x1] x [y0,y1].
// Returns false if the band is empty.  Conservative like gsr_band_xrange.
GSR_HD bool gsr_band_cells(const GsrEllipse& e, float ecut, int band, int x0, int x1, int y0, int y1,
                           int cl[2], int ch[2]) {
#if GSR_CFG_MASK_PER_BAND
  // one x-range for the whole band (the hull over its rows), given to every cell row the box reaches: the
  // vertical culling stays exact (the box's y-range is the ellipse's), the horizontal one is that of 8 rows
  int ya = band * GSR_RGH, yb = ya + GSR_RGH - 1, xl, xh;
  ya = ya > y0 ? ya : y0;
  yb = yb < y1 ? yb : y1;
  cl[0] = cl[1] = 1;
  ch[0] = ch[1] = 0;
  if (ya > yb || !gsr_band_xrange(e, ecut, ya, yb, x0, x1, xl, xh)) return false;
#pragma unroll
  for (int r = 0; r < GSR_CELLS_Y; ++r) {
    const int ra = band * GSR_RGH + r * GSR_CELL, rb_ = ra + GSR_CELL - 1;
    if (ra <= yb && rb_ >= ya) {
      cl[r] = xl / GSR_CELL;
      ch[r] = xh / GSR_CELL;
    }
  }
  return true;
#else
  bool any = false;
#pragma unroll
  for (int r = 0; r < GSR_CELLS_Y; ++r) {
    int ya = band * GSR_RGH + r * GSR_CELL, yb = ya + GSR_CELL - 1, xl, xh;
    ya = ya > y0 ? ya : y0;
    yb = yb < y1 ? yb : y1;
    cl[r] = 1;
    ch[r] = 0;
    if (ya <= yb && gsr_band_xrange(e, ecut, ya, yb, x0, x1, xl, xh)) {
      cl[r] = (int)((unsigned)xl / GSR_CELL);  // 0 <= xl <= xh: unsigned division is a shift
      ch[r] = (int)((unsigned)xh / GSR_CELL);
      any = true;
    }
  }
  return any;
#endif
}
// Region-column range [c0, c1] of a band whose cell ranges are (cl, ch) (at least one non-empty).
GSR_HD void gsr_band_columns(const int cl[2], const int ch[2], int& c0, int& c1) {
  const bool e0 = cl[0] <= ch[0], e1 = cl[1] <= ch[1];
  const int lo = e0 ? (e1 ? (cl[0] < cl[1] ? cl[0] : cl[1]) : cl[0]) : cl[1];
  const int hi = e0 ? (e1 ? (ch[0] > ch[1] ? ch[0] : ch[1]) : ch[0]) : ch[1];
  c0 = lo / GSR_CELLS_X;
  c1 = hi / GSR_CELLS_X;
}
// 8-bit cell mask of region column c.
GSR_HD uint32_t gsr_cell_mask(const int cl[2], const int ch[2], int c) {
  uint32_t m = 0;
#pragma unroll
  for (int r = 0; r < GSR_CELLS_Y; ++r) {
    int lo = cl[r] - c * GSR_CELLS_X, hi = ch[r] - c * GSR_CELLS_X;
    lo = lo > 0 ? lo : 0;
    hi = hi < GSR_CELLS_X - 1 ? hi : GSR_CELLS_X - 1;
    if (lo <= hi) m |= (((2u << hi) - 1u) & ~((1u << lo) - 1u)) << (r * GSR_CELLS_X);
  }
  return m;
}

// Row bitmaps: the same cell ranges as bit sets relative to cell column `cbase` (bit j = cell cbase + j), valid
// when the cull box spans at most 32 cells from cbase.  The cell mask of region column c is then two shifts away.
GSR_HD bool gsr_band_rowbits(const GsrEllipse& e, float ecut, int band, int x0, int x1, int y0, int y1, int cbase,
                             uint32_t rb[2]) {
  int cl[2], ch[2];
  if (!gsr_band_cells(e, ecut, band, x0, x1, y0, y1, cl, ch)) return false;
#pragma unroll
  for (int r = 0; r < GSR_CELLS_Y; ++r) {
    const int lo = cl[r] - cbase, hi = ch[r] - cbase;
    rb[r] = lo <= hi ? (((2u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
  }
  return true;
}
GSR_HD uint32_t gsr_rowbits_mask(const uint32_t rb[2], int c, int cbase) {
  const unsigned sh = (unsigned)(c * GSR_CELLS_X - cbase);
  return ((rb[0] >> sh) & 0xfu) | (((rb[1] >> sh) & 0xfu) << GSR_CELLS_X);
}

// Default and exact k-sigma handling shared by host API and kernels.
GSR_HD float gsr_effective_ksigma(float ksigma) {
  if (!(ksigma == ksigma) || ksigma <= 0.0f) return 5.0f;       // GSR_DEFAULT_KSIGMA
  if (ksigma > 13.25f) return 13.25f;                            // GSR_EXACT_KSIGMA
  return ksigma;
}
GSR_HD float gsr_ecut(float ksigma_eff) { return -0.5f * ksigma_eff * ksigma_eff * GSR_LOG2E; }
