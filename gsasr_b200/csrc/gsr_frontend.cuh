// gsr_frontend.cuh -- fused front end of the render path: raw head output -> raster inputs.
//
// Restates utils/gaussian_splatting.py:174-180 (activations) and :121-123 (unit / coordinate
// mapping of rendering_cuda_dmax) as one elementwise kernel, and its chain rule as another.
// The arithmetic follows what PyTorch executes on the INFERENCE path (inference_paper.py:113-131:
// sr_size and scale_modify are CPU tensors, so CUDA tensor / CPU-scalar divisions run as a
// multiplication by the fp32 reciprocal); every operation is individually rounded (no FMA
// contraction) so the mapped parameters agree with the unfused path to the last bit or one ulp.
#pragma once
#include "gsr_prepass.cuh"

// raw (s,9) -> sigmas (s,3), coords (s,2), colors (s,3): gsr_map_one (gsr_prepass.cuh) as a kernel of its own.
// The forward applies the same function inside its set-up kernel; this stand-alone form serves calls with more
// Gaussians than the region buckets index.
__global__ void __launch_bounds__(256)
gsr_map_kernel(const float* __restrict__ raw, float* __restrict__ sigmas,
               float* __restrict__ coords, float* __restrict__ colors, int s, int h, int w,
               float step, const GsrBDesc* __restrict__ bdesc = nullptr, int bn = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s) return;
  if (bdesc) {  // padded batch: every sample its own size and step
    const GsrBDesc d = bdesc[i / bn];
    h = d.h;
    w = d.w;
    step = d.step;
  }
  const GsrMapped m = gsr_map_one(raw + 9 * (size_t)i, h, w, step);
  sigmas[3 * (size_t)i + 0] = m.sx;
  sigmas[3 * (size_t)i + 1] = m.sy;
  sigmas[3 * (size_t)i + 2] = m.rho;
  coords[2 * (size_t)i + 0] = m.x;
  coords[2 * (size_t)i + 1] = m.y;
  colors[3 * (size_t)i + 0] = m.cr;
  colors[3 * (size_t)i + 1] = m.cg;
  colors[3 * (size_t)i + 2] = m.cb;
}

// Chain rule of gsr_map_kernel: (d/dsigmas, d/dcoords, d/dcolors) -> d/draw (s,9), written.
__global__ void __launch_bounds__(256)
gsr_unmap_kernel(const float* __restrict__ raw, const float* __restrict__ g_sigmas,
                 const float* __restrict__ g_coords, const float* __restrict__ g_colors,
                 float* __restrict__ g_raw, int s, int h, int w, float step,
                 const GsrBDesc* __restrict__ bdesc = nullptr, int bn = 0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s) return;
  if (bdesc) {
    const GsrBDesc d = bdesc[i / bn];
    h = d.h;
    w = d.w;
    step = d.step;
  }
  const float* p = raw + 9 * (size_t)i;
  float* o = g_raw + 9 * (size_t)i;
  const float kx = 2.0f / (step * (float)(w - 1)), ky = 2.0f / (step * (float)(h - 1));
  const float s0 = gsr_sigmoid(__ldg(p + 0)), s1 = gsr_sigmoid(__ldg(p + 1));
  const float th = tanhf(__ldg(p + 2));
  const float al = gsr_sigmoid(__ldg(p + 3));
  // sigmas[.,0] = sigma_y * kx (from raw column 1); sigmas[.,1] = sigma_x * ky (raw column 0)
  o[0] = g_sigmas[3 * (size_t)i + 1] * ky * 0.99999f * s0 * (1.0f - s0);
  o[1] = g_sigmas[3 * (size_t)i + 0] * kx * 0.99999f * s1 * (1.0f - s1);
  o[2] = g_sigmas[3 * (size_t)i + 2] * 0.999999f * (1.0f - th * th);
  float ga = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float sc = gsr_sigmoid(__ldg(p + 4 + c));
    const float gc = g_colors[3 * (size_t)i + c];
    o[4 + c] = gc * al * sc * (1.0f - sc);
    ga = fmaf(gc, sc, ga);
  }
  o[3] = ga * al * (1.0f - al);
  o[7] = g_coords[2 * (size_t)i + 0] * 2.0f * (float)w / (float)(w - 1);
  o[8] = g_coords[2 * (size_t)i + 1] * 2.0f * (float)h / (float)(h - 1);
}
