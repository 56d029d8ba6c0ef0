// ubench_tma.cu -- record staging of the forward raster, isolated: per-lane cp.async gathers by index (what
// gsr_forward_region_kernel does) against ONE cp.async.bulk (TMA, 1-D) of a contiguous record run + mbarrier
// (what a region-sorted record STREAM would allow), both double-buffered and feeding the same 2x2-block
// evaluation loop over every staged record.  Persistent warps, 24 per SM (6 CTAs x 4 warps), 64 records per
// stage as in the product kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tma tools/ubench_tma.cu && tools/ubench_tma
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

constexpr int CH = 64, WARPS = 4, STAGE = CH * 32;

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// evaluation of every record of a stage: the product kernel's per-record work (21 instructions per 128 evaluations)
__device__ __forceinline__ void eval_stage(uint32_t buf, f2 px2, f2 py2, f2& r0, f2& g0, f2& b0, f2& r1, f2& g1, f2& b1) {
#pragma unroll 8
  for (int j = 0; j < CH; ++j) {
    const float4 q0 = lds128(buf + j * 32), q1 = lds128(buf + j * 32 + 16);
    const f2 dx2 = add2(px2, pk(q0.x, q0.x)), dy2 = add2(py2, pk(q0.y, q0.y));
    const f2 t1 = mul2(pk(q0.w, q0.w), dy2), t0 = mul2(mul2(pk(q1.x, q1.x), dy2), dy2);
    float t1l, t1h, t0l, t0h;
    upk(t1, t1l, t1h);
    upk(t0, t0l, t0h);
    const f2 e0 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1l, t1l)), pk(t0l, t0l));
    const f2 e1 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1h, t1h)), pk(t0h, t0h));
    float e00, e01, e10, e11;
    upk(e0, e00, e01);
    upk(e1, e10, e11);
    const f2 v0 = pk(ex2(e00), ex2(e01)), v1 = pk(ex2(e10), ex2(e11));
    r0 = fma2(v0, pk(q1.y, q1.y), r0); g0 = fma2(v0, pk(q1.z, q1.z), g0); b0 = fma2(v0, pk(q1.w, q1.w), b0);
    r1 = fma2(v1, pk(q1.y, q1.y), r1); g1 = fma2(v1, pk(q1.z, q1.z), g1); b1 = fma2(v1, pk(q1.w, q1.w), b1);
  }
}

// MODE 0: gather by index with cp.async (2 x 16 B per record, two records per lane), wait_group
// MODE 1: one cp.async.bulk of the stage's 2 KB contiguous run, completion on an mbarrier
// MODE 2: evaluation only (records staged once): the floor both are measured against
template <int MODE>
__global__ void __launch_bounds__(32 * WARPS, 6)
stage_bench(const float4* __restrict__ rec, const float4* __restrict__ stream, const int* __restrict__ idx, int nchunks,
            float* sink, long long* cycles) {
  __shared__ __align__(128) unsigned char smem[WARPS][2][STAGE];
  __shared__ __align__(8) unsigned long long bar[WARPS][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
  const uint32_t buf0 = (uint32_t)__cvta_generic_to_shared(&smem[warp][0][0]);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bar[warp][0]);
  if (MODE == 1 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const f2 px2 = pk(0.01f * (lane & 3), 0.01f * (lane & 3) + 0.005f), py2 = pk(0.01f * (lane >> 2), 0.01f * (lane >> 2) + 0.005f);
  f2 r0 = pk(0, 0), g0 = r0, b0 = r0, r1 = r0, g1 = r0, b1 = r0;
  auto stage = [&](int c, int st) {
    if (c >= nchunks) return;
    const uint32_t dst = buf0 + st * STAGE;
    if (MODE == 0) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int k = lane + 32 * t;
        const float4* src = rec + 2 * (size_t)__ldg(idx + (size_t)c * CH + k);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + k * 32), "l"(src) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + k * 32 + 16), "l"(src + 1) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (MODE == 1) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + st * 8), "r"(STAGE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(stream + 2 * (size_t)c * CH), "r"(STAGE), "r"(bar0 + st * 8) : "memory");
      }
    }
  };
  const long long t0 = clock64();
  int c = gw, cur = 0;
  unsigned phase[2] = {0u, 0u};
  if (MODE == 2) {
    for (int k = lane; k < 2 * CH * 2; k += 32) reinterpret_cast<float4*>(&smem[warp][0][0])[k] = __ldg(stream + (size_t)gw * CH * 2 + (k % (CH * 2)));
    __syncwarp();
  } else {
    stage(c, 0);
  }
  for (; c < nchunks; c += nw) {
    if (MODE == 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    } else if (MODE == 1) {
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar0 + cur * 8), "r"(phase[cur]) : "memory");
      }
      phase[cur] ^= 1u;
      __syncwarp();
    }
    if (MODE != 2) stage(c + nw, cur ^ 1);   // the other stage was fully consumed one iteration ago
    eval_stage(buf0 + cur * STAGE, px2, py2, r0, g0, b0, r1, g1, b1);
    __syncwarp();
    if (MODE != 2) cur ^= 1;
  }
  const long long t1 = clock64();
  float x, y;
  upk(add2(add2(add2(r0, g0), add2(b0, r1)), add2(g1, b1)), x, y);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int N = 2097152, per_region = 160, nreg = 65536;          // HL: 2M records, ~160 entries per 16x8 region
  const int nchunks = nreg * per_region / CH;                       // chunks of 64 staged records
  std::vector<float4> rec(2 * (size_t)N);
  for (size_t i = 0; i < rec.size(); ++i) rec[i] = make_float4(-0.01f * (i & 7), -0.01f * (i & 3), -30.f, 2.f);
  // indices with the locality of a region bucket: the Gaussians of a region come from a few grid rows around it
  std::vector<int> idx((size_t)nchunks * CH);
  uint32_t s = 12345u;
  for (int r = 0; r < nreg; ++r)
    for (int e = 0; e < per_region; ++e) {
      s = s * 1664525u + 1013904223u;
      const int row = (r / 256) * 4 + (int)((s >> 8) % 24) - 10, col = (r % 256) * 8 + (int)((s >> 16) % 26) - 9;
      long long g = (long long)std::min(std::max(row, 0), 1023) * 2048 + std::min(std::max(col, 0), 2047);
      idx[(size_t)r * per_region + e] = (int)g;
    }
  std::vector<float4> stream(2 * idx.size());
  for (size_t i = 0; i < idx.size(); ++i) { stream[2 * i] = rec[2 * (size_t)idx[i]]; stream[2 * i + 1] = rec[2 * (size_t)idx[i] + 1]; }
  float4 *drec, *dstream; int* didx; float* sink; long long* cyc;
  cudaMalloc(&drec, rec.size() * 16); cudaMalloc(&dstream, stream.size() * 16); cudaMalloc(&didx, idx.size() * 4);
  cudaMalloc(&sink, (size_t)nsm * 6 * 128 * 4); cudaMalloc(&cyc, nsm * 6 * 8);
  cudaMemcpy(drec, rec.data(), rec.size() * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(dstream, stream.data(), stream.size() * 16, cudaMemcpyHostToDevice);
  cudaMemcpy(didx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
  const int grid = nsm * 6;
  auto run = [&](int mode, const char* name) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a);
      if (mode == 0) stage_bench<0><<<grid, 32 * WARPS>>>(drec, dstream, didx, nchunks, sink, cyc);
      if (mode == 1) stage_bench<1><<<grid, 32 * WARPS>>>(drec, dstream, didx, nchunks, sink, cyc);
      if (mode == 2) stage_bench<2><<<grid, 32 * WARPS>>>(drec, dstream, didx, nchunks, sink, cyc);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); best = std::min(best, ms);
    }
    cudaError_t e = cudaGetLastError();
    printf("%-62s %8.1f us  (%d chunks of %d records, %.0f MB staged)%s\n", name, best * 1e3, nchunks, CH,
           (double)nchunks * STAGE / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  run(2, "evaluation only (records staged once)");
  run(0, "cp.async gather by index (product kernel's staging)");
  run(1, "cp.async.bulk of a contiguous record run + mbarrier (TMA)");
  return 0;
}
