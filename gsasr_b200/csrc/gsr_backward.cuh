// gsr_backward.cuh -- backward raster kernel (sm_100a).
//
// Replaces _gs_render_backward_cuda (utils/gs_cuda_dmax/gs.cu:85-165; utils/gs_cuda/gs.cu:82-178).
// The reference walks ALL h*w pixels per Gaussian with 24 global read-modify-writes per
// in-window pixel.  Here one warp owns one Gaussian (taken in home-bin order, so neighbouring
// warps read neighbouring parts of the gradient image), sweeps only the Gaussian's cull box in
// 8x4 pixel patches, accumulates eight sums in registers, reduces them with warp shuffles and
// issues ONE plain read-modify-write per output value: no atomics, deterministic.
//
// Gradient algebra (gs.cu:139-159) refactored into moments: with u = v * sum_c g_c col_c,
//   Sx = sum u dx, Sy = sum u dy, Sxx = sum u dx^2, Sxy = sum u dx dy, Syy = sum u dy^2
//   d/dx   = 2 w1 (-w2 Sx + rho w3 Sy)            d/dy   = 2 w1 (-w4 Sy + rho w3 Sx)
//   d/dsx  = 2 w1 / sx (rho w3 Sxy - w2 Sxx)      d/dsy  = 2 w1 / sy (rho w3 Sxy - w4 Syy)
//   d/drho = -2 w1 (2 w1 rho (w2 Sxx - 2 rho w3 Sxy + w4 Syy) + w3 Sxy)
//   d/dcol_c = sum v g_c
#pragma once
#include "gsr_forward.cuh"

constexpr int GSR_BWD_THREADS = 256;
constexpr int GSR_BWD_WARPS = GSR_BWD_THREADS / 32;

struct GsrBwdArgs {
  const GsrRec* rec;
  const uint2* box;
  const int* ids;
  const int* bin_off;
  const float* px_tab;
  const float* py_tab;
  const float* grads;
  const float* sigmas;
  float* g_sigmas;
  float* g_coords;
  float* g_colors;
  int h, w, nb;
  uint32_t flags;
};

__device__ __forceinline__ float gsr_warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__global__ void __launch_bounds__(GSR_BWD_THREADS) gsr_backward_kernel(GsrBwdArgs p) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * GSR_BWD_WARPS + (threadIdx.x >> 5);
  const int n_live = __ldg(p.bin_off + p.nb + 1);
  if (gw >= n_live) return;

  const float4* rp = reinterpret_cast<const float4*>(p.rec + gw);
  const float4 a0 = __ldg(rp), a1 = __ldg(rp + 1);
  int bx0, bx1, by0, by1;
  bool binds;
  gsr_box_unpack(__ldg(p.box + gw), bx0, bx1, by0, by1, binds);

  const bool chw = (p.flags & 2u) != 0;
  const size_t plane = (size_t)p.h * p.w;
  const int lx = lane & 7, ly = lane >> 3;
  float cr = 0.f, cg = 0.f, cb = 0.f, sx_ = 0.f, sy_ = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f;

  for (int yb = by0; yb <= by1; yb += 4) {
    const int y = yb + ly;
    const bool yok = y <= by1;
    const float py = __ldg(p.py_tab + min(y, p.h - 1));
    const float dy = py - a0.y;
    const float t1 = a0.w * dy;
    const float t0 = a1.x * dy * dy;
    for (int xb = bx0; xb <= bx1; xb += 8) {
      const int x = xb + lx;
      const bool ok = yok && x <= bx1;
      const float px = __ldg(p.px_tab + min(x, p.w - 1));
      const float dx = px - a0.x;
      const float e = fmaf(dx, fmaf(a0.z, dx, t1), t0);
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
      if (ok) {
        if (chw) {
          const float* gp = p.grads + (size_t)y * p.w + x;
          g0 = __ldg(gp);
          g1 = __ldg(gp + plane);
          g2 = __ldg(gp + 2 * plane);
        } else {
          const float* gp = p.grads + ((size_t)y * p.w + x) * 3;
          g0 = __ldg(gp);
          g1 = __ldg(gp + 1);
          g2 = __ldg(gp + 2);
        }
      }
      const float v = ok ? gsr_ex2(e) : 0.f;
      cr = fmaf(v, g0, cr);
      cg = fmaf(v, g1, cg);
      cb = fmaf(v, g2, cb);
      const float G = fmaf(g0, a1.y, fmaf(g1, a1.z, g2 * a1.w));
      const float u = v * G;
      const float ux = u * dx, uy = u * dy;
      sx_ += ux;
      sy_ += uy;
      sxx = fmaf(ux, dx, sxx);
      sxy = fmaf(ux, dy, sxy);
      syy = fmaf(uy, dy, syy);
    }
  }
  cr = gsr_warp_sum(cr);
  cg = gsr_warp_sum(cg);
  cb = gsr_warp_sum(cb);
  sx_ = gsr_warp_sum(sx_);
  sy_ = gsr_warp_sum(sy_);
  sxx = gsr_warp_sum(sxx);
  sxy = gsr_warp_sum(sxy);
  syy = gsr_warp_sum(syy);

  if (lane == 0) {
    const int id = __ldg(p.ids + gw);
    const double sgx = (double)__ldg(p.sigmas + 3 * (size_t)id + 0);
    const double sgy = (double)__ldg(p.sigmas + 3 * (size_t)id + 1);
    const double rho = (double)__ldg(p.sigmas + 3 * (size_t)id + 2);
    const double w1 = -0.5 / (1.0 - rho * rho);
    const double w2 = 1.0 / (sgx * sgx), w3 = 1.0 / (sgx * sgy), w4 = 1.0 / (sgy * sgy);
    const double Sx = sx_, Sy = sy_, Sxx = sxx, Sxy = sxy, Syy = syy;
    const double gx = 2.0 * w1 * (-w2 * Sx + rho * w3 * Sy);
    const double gy = 2.0 * w1 * (-w4 * Sy + rho * w3 * Sx);
    const double gsx = 2.0 * w1 / sgx * (rho * w3 * Sxy - w2 * Sxx);
    const double gsy = 2.0 * w1 / sgy * (rho * w3 * Sxy - w4 * Syy);
    const double D = w2 * Sxx - 2.0 * rho * w3 * Sxy + w4 * Syy;
    const double grho = -2.0 * w1 * (2.0 * w1 * rho * D + w3 * Sxy);
    float* os = p.g_sigmas + 3 * (size_t)id;
    float* oc = p.g_coords + 2 * (size_t)id;
    float* ok = p.g_colors + 3 * (size_t)id;
    os[0] += (float)gsx;
    os[1] += (float)gsy;
    os[2] += (float)grho;
    oc[0] += (float)gx;
    oc[1] += (float)gy;
    ok[0] += cr;
    ok[1] += cg;
    ok[2] += cb;
  }
}
