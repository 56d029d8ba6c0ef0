// gsr_loss.cuh -- the pixel loss of the training loop, fused with its crop and its gradient.
//
// gsasr_model.py:191-233 renders every sample at its own size, pads it to the batch's largest size (F.pad, :212),
// crops output and ground truth back to the sample's size (:221-224) and adds L1Loss(reduction='mean') per sample
// (:226-228), divided by the batch size (:234):
//     loss = weight / B * sum_b  1 / (3 h_b w_b) * sum_{c, y < h_b, x < w_b} |sr[b,c,y,x] - gt[b,c,y,x]|
// One pass here: read sr and gt once, write dL/dsr (zero outside a sample's crop: the padding carries no loss) in
// sr's own layout -- the array gsr_backward_batch_padded takes as `grads` --, and the loss.  The reduction is
// two-stage in a fixed order (per-CTA partial sums, summed by the last CTA to finish): deterministic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int GSR_LOSS_MAX_BATCH = 128;  // samples per launch (sizes travel as a kernel argument)
constexpr int GSR_LOSS_THREADS = 256;

struct GsrLossArgs {
  const float* sr;
  const float* gt;
  float* grad;
  long long sr_n, sr_c, sr_h, sr_w;  // strides in floats: element (b, c, y, x)
  long long gt_n, gt_c, gt_h, gt_w;
  int batch, hmax, wmax;
  float weight;                      // loss_weight / total batch size
  double* partial;                   // gridDim.x doubles
  unsigned int* counter;             // zero on entry, left at zero
  float* loss;                       // accumulate == 0: written; else added to (second and later launches)
  int accumulate;
  short hw[GSR_LOSS_MAX_BATCH][2];
};

__global__ void __launch_bounds__(GSR_LOSS_THREADS) gsr_l1_crop_kernel(const GsrLossArgs a) {
  __shared__ double wsum[GSR_LOSS_THREADS / 32];
  __shared__ bool last;
  const long long per = (long long)a.hmax * a.wmax, total = per * a.batch;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const int r = (int)(i - (long long)b * per), y = r / a.wmax, x = r - y * a.wmax;
    const int hb = a.hw[b][0], wb = a.hw[b][1];
    const bool in = y < hb && x < wb;
    const float sc = a.weight / (3.0f * (float)hb * (float)wb);
    const float* ps = a.sr + b * a.sr_n + y * a.sr_h + x * a.sr_w;
    const float* pg = a.gt + b * a.gt_n + y * a.gt_h + x * a.gt_w;
    float* po = a.grad + b * a.sr_n + y * a.sr_h + x * a.sr_w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = 0.f;
      if (in) {
        const float d = ps[c * a.sr_c] - pg[c * a.gt_c];
        acc += fabsf(d) * sc;
        g = d > 0.f ? sc : (d < 0.f ? -sc : 0.f);  // torch: sign(0) = 0
      }
      po[c * a.sr_c] = g;
    }
  }
  double v = (double)acc;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < GSR_LOSS_THREADS / 32; ++k) t += wsum[k];
    a.partial[blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {  // fixed order: the same bits every run
    __threadfence();
    double t = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) t += *(volatile double*)(a.partial + k);
    *a.loss = a.accumulate ? *a.loss + (float)t : (float)t;
    *a.counter = 0u;
  }
}
