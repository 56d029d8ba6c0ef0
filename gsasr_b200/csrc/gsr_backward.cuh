// gsr_backward.cuh -- backward raster kernel (sm_100a).
//
// Replaces _gs_render_backward_cuda (utils/gs_cuda_dmax/gs.cu:85-165; utils/gs_cuda/gs.cu:82-178).
// The reference walks ALL h*w pixels per Gaussian with 24 global read-modify-writes per
// in-window pixel.  Here the work is Gaussian-centric but tile-staged:
//
//   * one CTA per 32x32 "super tile" (2x2 home bins).  It stages the gradient image of the tile
//     plus a halo (the reach of the Gaussians, capped at GSR_BWD_HALO) into shared memory as
//     three planes whose row stride is 8 mod 32 words, so that an 8x4 pixel patch is read
//     without bank conflicts;
//   * the Gaussians homed in the tile are dealt to the 16 warps; a warp sweeps one Gaussian's
//     cull box in 8x4 patches (lane = pixel), accumulates eight sums in registers, reduces them
//     with shuffles and lane 0 issues ONE plain read-modify-write per output value:
//     no atomics, deterministic;
//   * Gaussians whose box leaves the staged window, and the "large" list (extra CTAs at the end
//     of the grid), read the gradient image through L1/L2 instead.
//
// Gradient algebra (gs.cu:139-159) refactored into moments: with u = v * sum_c g_c col_c,
//   Sx = sum u dx, Sy = sum u dy, Sxx = sum u dx^2, Sxy = sum u dx dy, Syy = sum u dy^2
//   d/dx   = 2 w1 (-w2 Sx + rho w3 Sy)            d/dy   = 2 w1 (-w4 Sy + rho w3 Sx)
//   d/dsx  = 2 w1 / sx (rho w3 Sxy - w2 Sxx)      d/dsy  = 2 w1 / sy (rho w3 Sxy - w4 Syy)
//   d/drho = -2 w1 (2 w1 rho (w2 Sxx - 2 rho w3 Sxy + w4 Syy) + w3 Sxy)
//   d/dcol_c = sum v g_c
#pragma once
#include "gsr_forward_ws.cuh"

constexpr int GSR_BWD_THREADS = 512;
constexpr int GSR_BWD_WARPS = GSR_BWD_THREADS / 32;
constexpr int GSR_BWD_TILE = 32;                       // super tile side (multiple of GSR_BIN)
#ifndef GSR_CFG_BWD_BATCH
#define GSR_CFG_BWD_BATCH 16
#endif
#ifndef GSR_CFG_BWD_HALO
#define GSR_CFG_BWD_HALO 16
#endif
constexpr int GSR_BWD_HALO = GSR_CFG_BWD_HALO;         // staged halo cap, pixels (5 sigma at x4 is <= 16.7)
constexpr int GSR_BWD_RW = GSR_BWD_TILE + 2 * GSR_BWD_HALO;   // staged window side (64)
constexpr int GSR_BWD_RS = GSR_BWD_RW + 8;            // window row stride in pixels (one float4 each)
constexpr int GSR_BWD_PAD_X = 8, GSR_BWD_PAD_Y = 3;    // a sweep may run up to 7 columns / 3 rows past a box
constexpr int GSR_BWD_WIN = (GSR_BWD_RW + GSR_BWD_PAD_Y) * GSR_BWD_RS;   // float4 pixels of the staged window
constexpr int GSR_BWD_GCAP = 512;                      // Gaussians of a tile indexed per pass
constexpr int GSR_BWD_LARGE_CHUNK = 64;                // large-list Gaussians per extra CTA
static_assert(GSR_BWD_RS >= GSR_BWD_RW + GSR_BWD_PAD_X, "padded patch reads");
static_assert(GSR_BWD_TILE % GSR_BIN == 0, "super tile is made of whole bins");

constexpr int GSR_BWD_BATCH = GSR_CFG_BWD_BATCH;
#ifndef GSR_CFG_BWD_ELLIPSE_MIN_W
#define GSR_CFG_BWD_ELLIPSE_MIN_W 25
#endif
constexpr int GSR_BWD_ELLIPSE_MIN_W = GSR_CFG_BWD_ELLIPSE_MIN_W;  // Gaussians a warp reduces before one lane-parallel chain rule

struct GsrBwdSmem {
  float4 win[GSR_BWD_WIN];                     // {g_r, g_g, g_b, -}: one LDS.128 per pixel, a quarter-warp reads 128 contiguous bytes
  float px[GSR_BWD_RW + GSR_BWD_PAD_X];
  float py[GSR_BWD_RW + GSR_BWD_PAD_Y + 1];
  float4 zero4;                                // what lanes past a box read instead of a neighbour's gradients
  int gidx[GSR_BWD_GCAP];                      // sorted index of the tile's Gaussians (current pass)
  float tot[GSR_BWD_WARPS][GSR_BWD_BATCH][8];  // reduced sums of the current batch
  int tot_gi[GSR_BWD_WARPS][GSR_BWD_BATCH];    // sorted index of the Gaussians of the batch
};

struct GsrBwdArgs {
  const GsrRec* rec;
  const uint2* box;
  const int* ids;
  const int* bin_off;
  const int* stats;
  const float* px_tab;
  const float* py_tab;
  const float* grads;
  const float* sigmas;
  float* g_sigmas;
  float* g_coords;
  float* g_colors;
  int h, w, nbx, nby, nb;
  int tiles_x, tiles_y;
  uint32_t flags;
  const GsrBDesc* bdesc;  // padded batch (ragged != 0): records and moments are in canvas coordinates
  int bn, ragged;
  int hf, row0, bhs;      // row-band view / rows per sample of a stacked batch (see gsr_setup, gsr_region_mask)
  const int* guard;       // run only if *guard == want (nullptr: always)
  int want;
};

// The k-sigma ellipse of a sorted record in the pixel units of the image the kernel sweeps (band-local rows; the
// sample's block of rows of a stacked batch), and the cut-off of the set-up that made the boxes.
__device__ __forceinline__ GsrEllipse gsr_bwd_ellipse(const GsrBwdArgs& p, const float4& a0, const float4& a1, int by0) {
  GsrRec r;
  r.x = a0.x, r.y = a0.y, r.a = a0.z, r.b = a0.w, r.c = a1.x, r.r = r.g = r.bl = 0.f;
  GsrEllipse e = gsr_ellipse(r, p.bhs > 0 ? p.bhs : p.h, p.w, p.hf, p.row0);
  if (p.bhs > 0) e.cy += (float)((by0 / p.bhs) * p.bhs);
  return e;
}

__device__ __forceinline__ float gsr_warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

struct GsrBwdAcc {
  gsr_f2 crg;   // {sum v g_r, sum v g_g}
  float cb;     // sum v g_b
  gsr_f2 s1;    // {Sx, Sy}
  gsr_f2 s2;    // {Sxx, Sxy}
  float syy;
};
__device__ __forceinline__ GsrBwdAcc gsr_bwd_acc_zero() {
  GsrBwdAcc a;
  a.crg = a.s1 = a.s2 = gsr_pk(0.f, 0.f);
  a.cb = a.syy = 0.f;
  return a;
}

// One pixel's contribution (packed FP32x2 where two sums share a factor).  v must already be 0
// for lanes outside the cull box.
__device__ __forceinline__ void gsr_bwd_accum(GsrBwdAcc& a, float v, float g0, float g1, float g2,
                                              float dx, float dy, const float4& a1) {
  a.crg = gsr_fma2(gsr_pk(v, v), gsr_pk(g0, g1), a.crg);
  a.cb = fmaf(v, g2, a.cb);
  const float G = fmaf(g0, a1.y, fmaf(g1, a1.z, g2 * a1.w));
  const float u = v * G;
  const gsr_f2 d2 = gsr_pk(dx, dy);
  a.s1 = gsr_fma2(gsr_pk(u, u), d2, a.s1);
  const float ux = u * dx, uy = u * dy;
  a.s2 = gsr_fma2(gsr_pk(ux, ux), d2, a.s2);
  a.syy = fmaf(uy, dy, a.syy);
}

// Sweep of one Gaussian's cull box with gradients read from the staged planes (explicit 32-bit
// shared addresses, plane offsets as immediates).
template <int OFF>
__device__ __forceinline__ float gsr_lds32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}

__device__ __forceinline__ void gsr_bwd_sweep_smem_box(GsrBwdAcc& acc, uint32_t win_s, uint32_t px_s,
                                                   uint32_t py_s, const float4& a0, const float4& a1,
                                                   int bx0, int bx1, int by0, int by1, int sx0,
                                                   int sy0, int lx, int ly, float xstep8, uint32_t zero_s) {
  // Lanes past the box read zero padding (see the staging code) and contribute v = 0.
  const int nx = (bx1 - bx0 + 8) >> 3;            // 8-pixel steps per row
  const int x0 = bx0 - sx0 + lx, xlast = bx1 - sx0;
  // dx of this lane's first column; the following columns are 8 pixels further each.  (The forward's
  // bit-exact pixel table is not needed here: the inclusion set comes from the integer box, and the
  // gradient tolerance is 1e-3 relative.)
  const float dx_first = gsr_lds32<0>(px_s + x0 * 4) - a0.x;
  for (int y = by0 + ly; y - ly <= by1; y += 4) {
    const int yc = y - sy0;
    const bool yok = y <= by1;
    const float dy = gsr_lds32<0>(py_s + yc * 4) - a0.y;
    const float t1 = a0.w * dy;
    const float t0 = a1.x * dy * dy;
    uint32_t ad = win_s + (uint32_t)(yc * GSR_BWD_RS + x0) * 16u;
    float dx = dx_first;
    int xi = x0;
    for (int k = 0; k < nx; ++k, ad += 128, xi += 8, dx += xstep8) {
      const float e = fmaf(dx, fmaf(a0.z, dx, t1), t0);
      // lanes past the box read a zero pixel instead of a neighbour's real gradients: an inf / NaN in dL/dimg
      // outside this Gaussian's box cannot reach its sums as 0 * inf
      const bool ok = yok && xi <= xlast;
      const float4 g = gsr_lds128(ok ? ad : zero_s);
      const float v = ok ? gsr_ex2(e) : 0.f;
      gsr_bwd_accum(acc, v, g.x, g.y, g.z, dx, dy, a1);
    }
  }
}

// Wide boxes (x8 fields: 5 sigma = 33 px): every 4-row strip of the box is swept only over the columns the
// Gaussian's k-sigma ellipse reaches in it (gsr_band_xrange, warp-uniform) instead of the box's full width;
// pixels it skips carry less than exp(-k^2/2), the same truncation the forward applies.  For the 17-pixel boxes
// of a x4 field the per-strip range costs more than the patches it saves (measured: HL 1.84 ms vs 1.43 ms), so
// boxes narrower than GSR_BWD_ELLIPSE_MIN_W take the rectangular sweep above.
__device__ __forceinline__ void gsr_bwd_sweep_smem(GsrBwdAcc& acc, uint32_t win_s, uint32_t px_s,
                                                   uint32_t py_s, const float4& a0, const float4& a1,
                                                   int bx0, int bx1, int by0, int by1, int sx0,
                                                   int sy0, int lx, int ly, float xstep8, const GsrEllipse* el,
                                                   float ecut, uint32_t zero_s) {
  // Lanes past the box read zero padding (see the staging code) and contribute v = 0.
  for (int yb = by0; yb <= by1; yb += 4) {
    int xa = bx0, xb = bx1;
    if (el) {
      const int ye = min(yb + 3, by1);
      if (!gsr_band_xrange(*el, ecut, yb, ye, bx0, bx1, xa, xb)) continue;
    }
    const int y = yb + ly;
    const int nx = (xb - xa + 8) >> 3;            // 8-pixel steps per row
    const int x0 = xa - sx0 + lx, xlast = xb - sx0;
    // dx of this lane's first column; the following columns are 8 pixels further each.  (The forward's
    // bit-exact pixel table is not needed here: the inclusion set comes from the integer box, and the
    // gradient tolerance is 1e-3 relative.)
    const int yc = y - sy0;
    const bool yok = y <= by1;
    const float dy = gsr_lds32<0>(py_s + yc * 4) - a0.y;
    const float t1 = a0.w * dy;
    const float t0 = a1.x * dy * dy;
    uint32_t ad = win_s + (uint32_t)(yc * GSR_BWD_RS + x0) * 16u;
    float dx = gsr_lds32<0>(px_s + x0 * 4) - a0.x;
    int xi = x0;
    for (int k = 0; k < nx; ++k, ad += 128, xi += 8, dx += xstep8) {
      const float e = fmaf(dx, fmaf(a0.z, dx, t1), t0);
      const bool ok = yok && xi <= xlast;
      const float4 g = gsr_lds128(ok ? ad : zero_s);  // (see gsr_bwd_sweep_smem_box)
      const float v = ok ? gsr_ex2(e) : 0.f;
      gsr_bwd_accum(acc, v, g.x, g.y, g.z, dx, dy, a1);
    }
  }
}

// Same sweep with gradients read from global memory (through L1/L2).
__device__ __forceinline__ void gsr_bwd_sweep_gmem(GsrBwdAcc& acc, const GsrBwdArgs& p,
                                                   const float4& a0, const float4& a1, int bx0,
                                                   int bx1, int by0, int by1, int lx, int ly, const GsrEllipse* el,
                                                   float ecut) {
  const bool chw = (p.flags & 2u) != 0;
  const size_t plane = (size_t)p.h * p.w;
  for (int yb = by0; yb <= by1; yb += 4) {
    int xa = bx0, xe = bx1;
    if (el) {
      const int ye = min(yb + 3, by1);
      if (!gsr_band_xrange(*el, ecut, yb, ye, bx0, bx1, xa, xe)) continue;
    }
    const int y = yb + ly;
    const int yc = min(y, by1);
    const bool yok = y <= by1;
    const float dy = __ldg(p.py_tab + yc) - a0.y;
    const float t1 = a0.w * dy;
    const float t0 = a1.x * dy * dy;
    for (int xb = xa; xb <= xe; xb += 8) {
      const int x = xb + lx;
      const int xc = min(x, xe);
      const bool ok = yok && x <= xe;
      const float dx = __ldg(p.px_tab + xc) - a0.x;
      const float e = fmaf(dx, fmaf(a0.z, dx, t1), t0);
      float g0, g1, g2;
      if (chw) {
        const float* gp = p.grads + (size_t)yc * p.w + xc;
        g0 = __ldg(gp);
        g1 = __ldg(gp + plane);
        g2 = __ldg(gp + 2 * plane);
      } else {
        const float* gp = p.grads + ((size_t)yc * p.w + xc) * 3;
        g0 = __ldg(gp);
        g1 = __ldg(gp + 1);
        g2 = __ldg(gp + 2);
      }
      const float v = ok ? gsr_ex2(e) : 0.f;
      gsr_bwd_accum(acc, v, g0, g1, g2, dx, dy, a1);
    }
  }
}

// Warp-reduce the eight sums (recursive halving: 9 shuffles instead of 40, then 8 to collect) and
// let lane 0 apply the chain rule and accumulate the outputs.  The chain rule is written in
// terms of the raster conic (a, b, c) = log2(e) * w1 * (w2, -2 rho w3, w4):
//   d/dx   = -(2 a Sx + b Sy) / L           d/dy   = -(2 c Sy + b Sx) / L
//   d/dsx  = -(b Sxy + 2 a Sxx) / (L sx)    d/dsy  = -(b Sxy + 2 c Syy) / (L sy)
//   d/drho = (2 rho Q / L + Sxy / (sx sy)) / (1 - rho^2),   Q = a Sxx + b Sxy + c Syy
// (Q cancels strongly for |rho| -> 1 and is formed in double.)
// Recursive-halving reduction of the eight sums: afterwards total k lives in the lanes whose
// (bit4, bit3, bit2) = (k>>2&1, k>>1&1, k&1); lanes 0,4,...,28 store one total each.
__device__ __forceinline__ void gsr_bwd_reduce_store(const GsrBwdAcc& acc, int lane, float* dst8) {
  float v[8];
  gsr_upk(acc.crg, v[0], v[1]);
  v[2] = acc.cb;
  gsr_upk(acc.s1, v[3], v[4]);
  gsr_upk(acc.s2, v[5], v[6]);
  v[7] = acc.syy;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float w4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  float w2[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b3 ? w4[i] : w4[i + 2], keep = b3 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float x = (b2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? w2[0] : w2[1], 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  if ((lane & 3) == 0) dst8[(b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0)] = x;
}

// Chain rule for ONE Gaussian (executed lane-parallel: one lane per Gaussian of a batch) and the
// single read-modify-write per output value.
__device__ __forceinline__ void gsr_bwd_chain(const float* t, const GsrBwdArgs& p, int gi) {
  const float4* rp = reinterpret_cast<const float4*>(p.rec + gi);
  const float4 a0 = __ldg(rp), a1 = __ldg(rp + 1);
  const int id = __ldg(p.ids + gi);
  const float sgx = __ldg(p.sigmas + 3 * (size_t)id + 0);
  const float sgy = __ldg(p.sigmas + 3 * (size_t)id + 1);
  const float rho = __ldg(p.sigmas + 3 * (size_t)id + 2);
  const float Sx = t[3], Sy = t[4], Sxx = t[5], Syy = t[7];
  float Sxy = t[6];
  const float a = a0.z, b = a0.w, c = a1.x;
  const float iL = 0.6931471805599453f;  // 1 / log2(e)
  float gx = -(2.0f * a * Sx + b * Sy) * iL;
  float gy = -(2.0f * c * Sy + b * Sx) * iL;
  float sxy_own = Sxy;  // sum u dx dy in the sample's OWN units (the explicit Sxy term of d/drho)
  if (p.ragged) {
    // padded batch: d_own = (ax, ay) * d_canvas.  The products conic x moment are invariant (so are the
    // sigma gradients and Q); the centre gradients and the bare moment are not.
    const GsrBDesc d = p.bdesc[id / p.bn];
    gx = (float)((double)gx / d.ax);
    gy = (float)((double)gy / d.ay);
    sxy_own = (float)((double)Sxy * d.ax * d.ay);
  }
  const float gsx = -(b * Sxy + 2.0f * a * Sxx) * iL / sgx;
  const float gsy = -(b * Sxy + 2.0f * c * Syy) * iL / sgy;
  const double Q = (double)a * Sxx + (double)b * Sxy + (double)c * Syy;
  const float grho = (float)((2.0 * (double)rho * Q * (double)iL + (double)sxy_own / ((double)sgx * sgy)) /
                             (1.0 - (double)rho * rho));
  float* os = p.g_sigmas + 3 * (size_t)id;
  float* oc = p.g_coords + 2 * (size_t)id;
  float* ok = p.g_colors + 3 * (size_t)id;
  os[0] += gsx;
  os[1] += gsy;
  os[2] += grho;
  oc[0] += gx;
  oc[1] += gy;
  ok[0] += t[0];
  ok[1] += t[1];
  ok[2] += t[2];
}

// One block of work: a 64x64 tile's Gaussians (bid < tiles) or a chunk of the "large" list.
__device__ __forceinline__ void gsr_backward_block(const GsrBwdArgs& p, GsrBwdSmem& sm, const int bid) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lx = lane & 7, ly = lane >> 3;
  const int ntiles = p.tiles_x * p.tiles_y;
  const float ecut = gsr_ecut(__int_as_float(__ldg(p.stats + GSR_STAT_KSIGMA)));  // the k of the set-up that made the boxes

  if (bid >= ntiles) {
    // ---- "large" list: no staging, one chunk per CTA ----
    const int l0 = __ldg(p.bin_off + p.nb), l1 = __ldg(p.bin_off + p.nb + 1);
    const int c0 = l0 + (bid - ntiles) * GSR_BWD_LARGE_CHUNK;
    const int c1 = min(c0 + GSR_BWD_LARGE_CHUNK, l1);
    for (int g0 = c0 + warp; g0 < c1; g0 += GSR_BWD_WARPS * GSR_BWD_BATCH) {
      int nb = 0;
      for (int j = 0; j < GSR_BWD_BATCH; ++j) {
        const int gi = g0 + j * GSR_BWD_WARPS;
        if (gi >= c1) break;
        const float4* rp = reinterpret_cast<const float4*>(p.rec + gi);
        const float4 a0 = __ldg(rp), a1 = __ldg(rp + 1);
        int bx0, bx1, by0, by1;
        bool binds;
        gsr_box_unpack(__ldg(p.box + gi), bx0, bx1, by0, by1, binds);
        GsrBwdAcc acc = gsr_bwd_acc_zero();
        const GsrEllipse el = gsr_bwd_ellipse(p, a0, a1, by0);
        gsr_bwd_sweep_gmem(acc, p, a0, a1, bx0, bx1, by0, by1, lx, ly, &el, ecut);
        gsr_bwd_reduce_store(acc, lane, sm.tot[warp][j]);
        if (lane == 0) sm.tot_gi[warp][j] = gi;
        ++nb;
      }
      __syncwarp();
      if (lane < nb) gsr_bwd_chain(sm.tot[warp][lane], p, sm.tot_gi[warp][lane]);
      __syncwarp();
    }
    return;
  }

  const int tx = bid % p.tiles_x, ty = bid / p.tiles_x;
  const int tx0 = tx * GSR_BWD_TILE, ty0 = ty * GSR_BWD_TILE;
  // Gaussians homed in this tile: one contiguous run per bin row.
  constexpr int BPT = GSR_BWD_TILE / GSR_BIN;
  const int bx_lo = tx * BPT, bx_hi = min(bx_lo + BPT, p.nbx);
  int run_s[BPT], run_n[BPT], ntot = 0;
#pragma unroll
  for (int r = 0; r < BPT; ++r) {
    const int by = ty * BPT + r;
    run_s[r] = 0;
    run_n[r] = 0;
    if (by < p.nby) {
      run_s[r] = __ldg(p.bin_off + by * p.nbx + bx_lo);
      run_n[r] = __ldg(p.bin_off + by * p.nbx + bx_hi) - run_s[r];
    }
    ntot += run_n[r];
  }
  if (ntot == 0) return;

  // ---- stage the gradient window ----
  const int hx = min(__ldg(p.stats + 0), GSR_BWD_HALO), hy = min(__ldg(p.stats + 1), GSR_BWD_HALO);
  const int sx0 = max(tx0 - hx, 0), sx1 = min(tx0 + GSR_BWD_TILE - 1 + hx, p.w - 1);
  const int sy0 = max(ty0 - hy, 0), sy1 = min(ty0 + GSR_BWD_TILE - 1 + hy, p.h - 1);
  const int rw = sx1 - sx0 + 1, rh = sy1 - sy0 + 1;
  // Coordinate tables and the zero padding a sweep may read past a box (7 columns / 3 rows): finite
  // values there, multiplied by v = 0.
  if (tid < rw + GSR_BWD_PAD_X) sm.px[tid] = tid < rw ? __ldg(p.px_tab + sx0 + tid) : 0.f;
  if (tid >= 128 && tid - 128 < rh + GSR_BWD_PAD_Y)
    sm.py[tid - 128] = tid - 128 < rh ? __ldg(p.py_tab + sy0 + tid - 128) : 0.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < (rh + GSR_BWD_PAD_Y) * GSR_BWD_PAD_X; e += GSR_BWD_THREADS)  // right margin
    sm.win[(e / GSR_BWD_PAD_X) * GSR_BWD_RS + rw + (e % GSR_BWD_PAD_X)] = zero4;
  for (int e = tid; e < GSR_BWD_PAD_Y * rw; e += GSR_BWD_THREADS)  // bottom margin
    sm.win[(rh + e / rw) * GSR_BWD_RS + (e % rw)] = zero4;
  float* const winf = reinterpret_cast<float*>(sm.win);
  if (p.flags & 2u) {  // CHW source: plane by plane
    const size_t gplane = (size_t)p.h * p.w;
    for (int r = warp; r < 3 * rh; r += GSR_BWD_WARPS) {
      const int c = r / rh, yi = r - c * rh;
      const float* src = p.grads + c * gplane + (size_t)(sy0 + yi) * p.w + sx0;
      float* dst = winf + (size_t)yi * GSR_BWD_RS * 4 + c;
      for (int e = lane; e < rw; e += 32) dst[e * 4] = __ldg(src + e);
    }
  } else {  // HWC source: a row of the window is 3*rw contiguous floats
    for (int yi = warp; yi < rh; yi += GSR_BWD_WARPS) {
      const float* src = p.grads + ((size_t)(sy0 + yi) * p.w + sx0) * 3;
      float* dst = winf + (size_t)yi * GSR_BWD_RS * 4;
      for (int e = lane; e < 3 * rw; e += 32) {
        const int x = e / 3, c = e - 3 * x;
        dst[x * 4 + c] = __ldg(src + e);
      }
    }
  }

  const uint32_t win_s = gsr_smem_addr(sm.win), px_s = gsr_smem_addr(sm.px), py_s = gsr_smem_addr(sm.py);
  const uint32_t zero_s = gsr_smem_addr(&sm.zero4);
  if (tid == 0) sm.zero4 = zero4;
  const float xstep8 = 16.0f / (float)(p.w - 1);  // eight pixels in normalised units
  for (int pass0 = 0; pass0 < ntot; pass0 += GSR_BWD_GCAP) {
    // ---- index the Gaussians of this pass: position in the tile -> sorted index (one lookup each,
    // instead of one per warp per Gaussian)
    const int npass = min(ntot - pass0, GSR_BWD_GCAP);
    __syncthreads();  // the previous pass is done with gidx
    for (int t = tid; t < npass; t += GSR_BWD_THREADS) {
      int gi = 0, kk = pass0 + t;
#pragma unroll
      for (int r = 0; r < BPT; ++r) {
        if (kk >= 0 && kk < run_n[r]) gi = run_s[r] + kk;
        kk = kk < run_n[r] ? -1 : kk - run_n[r];
      }
      sm.gidx[t] = gi;
    }
    __syncthreads();  // also orders the staging above before the sweeps

    // ---- one Gaussian per warp at a time, chain rule once per batch ----
    for (int k0 = warp; k0 < npass; k0 += GSR_BWD_WARPS * GSR_BWD_BATCH) {
      int nb = 0;
      for (int j = 0; j < GSR_BWD_BATCH; ++j) {
        const int k = k0 + j * GSR_BWD_WARPS;
        if (k >= npass) break;
        const int gi = sm.gidx[k];
        const float4* rp = reinterpret_cast<const float4*>(p.rec + gi);
        const float4 a0 = __ldg(rp), a1 = __ldg(rp + 1);
        int bx0, bx1, by0, by1;
        bool binds;
        gsr_box_unpack(__ldg(p.box + gi), bx0, bx1, by0, by1, binds);
        GsrBwdAcc acc = gsr_bwd_acc_zero();
        const bool wide = bx1 - bx0 + 1 >= GSR_BWD_ELLIPSE_MIN_W;  // warp-uniform
        const bool staged = bx0 >= sx0 && bx1 <= sx1 && by0 >= sy0 && by1 <= sy1;
        if (staged && !wide) {
          gsr_bwd_sweep_smem_box(acc, win_s, px_s, py_s, a0, a1, bx0, bx1, by0, by1, sx0, sy0, lx, ly, xstep8, zero_s);
        } else {
          const GsrEllipse el = gsr_bwd_ellipse(p, a0, a1, by0);
          if (staged)
            gsr_bwd_sweep_smem(acc, win_s, px_s, py_s, a0, a1, bx0, bx1, by0, by1, sx0, sy0, lx, ly, xstep8, &el, ecut, zero_s);
          else
            gsr_bwd_sweep_gmem(acc, p, a0, a1, bx0, bx1, by0, by1, lx, ly, wide ? &el : nullptr, ecut);
        }
        gsr_bwd_reduce_store(acc, lane, sm.tot[warp][j]);
        if (lane == 0) sm.tot_gi[warp][j] = gi;
        ++nb;
      }
      __syncwarp();
      if (lane < nb) gsr_bwd_chain(sm.tot[warp][lane], p, sm.tot_gi[warp][lane]);
      __syncwarp();
    }
  }
}

// The grid is the number of blocks of work when the launch is unconditional; a guarded launch (the fallback behind
// the region backward) uses a small grid that strides over them: 40,960 CTAs that only read the guard and leave
// took 169 us at the headline shape.
__global__ void __launch_bounds__(GSR_BWD_THREADS, 2) gsr_backward_kernel(GsrBwdArgs p, int nblocks) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  extern __shared__ __align__(16) unsigned char gsr_smem_raw[];
  GsrBwdSmem& sm = *reinterpret_cast<GsrBwdSmem*>(gsr_smem_raw);
  for (int bid = blockIdx.x; bid < nblocks; bid += gridDim.x) {
    gsr_backward_block(p, sm, bid);
    __syncthreads();  // shared memory is reused by the next block of work
  }
}
