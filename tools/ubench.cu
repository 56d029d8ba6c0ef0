// ubench.cu -- pipe-throughput microbenchmarks behind the forward kernel's design:
// how many SM cycles one warp instruction of each kind costs when the SM is saturated with them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

constexpr int ITERS = 2048, UNROLL = 8;

enum Kind { FFMA, FFMA2, MUFU, LDS128_SAME, LDS128_HALF, LDS128_QUARTER, LDS128_EIGHTH, LDS128_LANE, LDS64_SAME,
            LDS32_SAME, SHFL, MIX2x2, MIX1x2, MIX2x2_NOCOL, MIX_MMA, FFMA2_BCAST, MIX2x2_SCOL, MIX2x2_SCOLEXP, MIX2x2_SCALAR, NKIND };
static const char* kNames[NKIND] = {"ffma", "ffma2", "mufu.ex2", "lds128 one address/warp", "lds128 one address/half-warp",
                                    "lds128 one address/quarter-warp", "lds128 one address/4 lanes",
                                    "lds128 lane-distinct (conflict-free)", "lds64 one address/warp",
                                    "lds32 one address/warp", "shfl.idx",
                                    "eval 2x2 px/lane (per record-iteration)", "eval 1x2 px/lane (per record-iteration)",
                                    "eval 2x2 px/lane, colour FMAs removed", "eval column/lane + tf32 MMA colours (per 128 evals)",
                                    "ffma2 scalar-broadcast operands, distinct regs",
                                    "eval 2x2 px/lane, scalar colour FMAs", "eval 2x2 px/lane, scalar colour + exponent FMAs",
                                    "eval 2x2 px/lane, all scalar"};

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

template <int KIND>
__global__ void __launch_bounds__(1024) bench(long long* cycles, float* sink, float seed) {
  __shared__ __align__(16) float sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = seed * (i & 15) * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  uint32_t addr = base;
  if (KIND == LDS128_HALF) addr = base + (lane >> 4) * 32;
  if (KIND == LDS128_QUARTER) addr = base + (lane >> 3) * 32;
  if (KIND == LDS128_EIGHTH) addr = base + (lane >> 2) * 32;
  if (KIND == LDS128_LANE) addr = base + lane * 16;
  float a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  f2 p0 = pk(a0, a1), p1 = pk(a2, a3), p2 = pk(a4, a5), p3 = pk(a6, a7), p4 = p0, p5 = p1, p6 = p2, p7 = p3;
  const f2 pc = pk(seed, seed);
  __syncthreads();
  const long long t0 = clock64();
  if (KIND == MIX2x2 || KIND == MIX1x2) {
    // the forward kernel's inner loop, one record per iteration per (half-)warp
    const float py0 = seed * lane, py1 = py0 + seed;
    const f2 px2 = pk(seed * (lane & 3), seed * (lane & 3) + seed), py2 = pk(py0, py1);
    f2 r0 = pk(0, 0), g0 = r0, b0 = r0, r1 = r0, g1 = r0, b1 = r0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint32_t ad = (KIND == MIX2x2 ? base + (lane >> 4) * 2048 : base) + ((it * UNROLL + u) & 63) * 32;
        const float4 q0 = lds128(ad), q1 = lds128(ad + 16);  // -x, -y, a, b | c, r, g, bl
        const f2 dx2 = add2(px2, pk(q0.x, q0.x));
        if (KIND == MIX2x2) {
          const f2 dy2 = add2(py2, pk(q0.y, q0.y));
          const f2 t1 = mul2(pk(q0.w, q0.w), dy2);
          const f2 t0 = mul2(mul2(pk(q1.x, q1.x), dy2), dy2);
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
        } else {
          const float dy = py0 + q0.y;
          const float t1 = q0.w * dy, t0 = q1.x * dy * dy;
          const f2 e0 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1, t1)), pk(t0, t0));
          float e00, e01;
          upk(e0, e00, e01);
          const f2 v0 = pk(ex2(e00), ex2(e01));
          r0 = fma2(v0, pk(q1.y, q1.y), r0); g0 = fma2(v0, pk(q1.z, q1.z), g0); b0 = fma2(v0, pk(q1.w, q1.w), b0);
        }
      }
    }
    float x, y;
    upk(add2(add2(add2(r0, g0), add2(b0, r1)), add2(g1, b1)), x, y);
    a0 = x + y;
  } else if (KIND == MIX2x2_SCOL || KIND == MIX2x2_SCOLEXP || KIND == MIX2x2_SCALAR) {
    const float py0 = seed * lane, py1 = py0 + seed, px0 = seed * (lane & 3), px1 = px0 + seed;
    const f2 px2 = pk(px0, px1), py2 = pk(py0, py1);
    float c[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i] = 0.f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint32_t ad = base + (lane >> 4) * 2048 + ((it * UNROLL + u) & 63) * 32;
        const float4 q0 = lds128(ad), q1 = lds128(ad + 16);
        float e00, e01, e10, e11;
        if (KIND == MIX2x2_SCALAR) {
          const float dx0 = px0 + q0.x, dx1 = px1 + q0.x, dy0 = py0 + q0.y, dy1 = py1 + q0.y;
          const float t1a = q0.w * dy0, t1b = q0.w * dy1, t0a = q1.x * dy0 * dy0, t0b = q1.x * dy1 * dy1;
          e00 = fmaf(dx0, fmaf(q0.z, dx0, t1a), t0a); e01 = fmaf(dx1, fmaf(q0.z, dx1, t1a), t0a);
          e10 = fmaf(dx0, fmaf(q0.z, dx0, t1b), t0b); e11 = fmaf(dx1, fmaf(q0.z, dx1, t1b), t0b);
        } else {
          const f2 dx2 = add2(px2, pk(q0.x, q0.x));
          const f2 dy2 = add2(py2, pk(q0.y, q0.y));
          const f2 t1 = mul2(pk(q0.w, q0.w), dy2);
          const f2 t0 = mul2(mul2(pk(q1.x, q1.x), dy2), dy2);
          float t1l, t1h, t0l, t0h;
          upk(t1, t1l, t1h);
          upk(t0, t0l, t0h);
          if (KIND == MIX2x2_SCOLEXP) {
            float dx0, dx1;
            upk(dx2, dx0, dx1);
            e00 = fmaf(dx0, fmaf(q0.z, dx0, t1l), t0l); e01 = fmaf(dx1, fmaf(q0.z, dx1, t1l), t0l);
            e10 = fmaf(dx0, fmaf(q0.z, dx0, t1h), t0h); e11 = fmaf(dx1, fmaf(q0.z, dx1, t1h), t0h);
          } else {
            const f2 e0 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1l, t1l)), pk(t0l, t0l));
            const f2 e1 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1h, t1h)), pk(t0h, t0h));
            upk(e0, e00, e01);
            upk(e1, e10, e11);
          }
        }
        const float v00 = ex2(e00), v01 = ex2(e01), v10 = ex2(e10), v11 = ex2(e11);
        c[0] = fmaf(v00, q1.y, c[0]); c[1] = fmaf(v00, q1.z, c[1]); c[2] = fmaf(v00, q1.w, c[2]);
        c[3] = fmaf(v01, q1.y, c[3]); c[4] = fmaf(v01, q1.z, c[4]); c[5] = fmaf(v01, q1.w, c[5]);
        c[6] = fmaf(v10, q1.y, c[6]); c[7] = fmaf(v10, q1.z, c[7]); c[8] = fmaf(v10, q1.w, c[8]);
        c[9] = fmaf(v11, q1.y, c[9]); c[10] = fmaf(v11, q1.z, c[10]); c[11] = fmaf(v11, q1.w, c[11]);
      }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) a0 += c[i];
  } else if (KIND == MIX2x2_NOCOL) {
    const float py0 = seed * lane, py1 = py0 + seed;
    const f2 px2 = pk(seed * (lane & 3), seed * (lane & 3) + seed), py2 = pk(py0, py1);
    f2 r0 = pk(0, 0), r1 = r0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint32_t ad = base + (lane >> 4) * 2048 + ((it * UNROLL + u) & 63) * 32;
        const float4 q0 = lds128(ad), q1 = lds128(ad + 16);
        const f2 dx2 = add2(px2, pk(q0.x, q0.x));
        const f2 dy2 = add2(py2, pk(q0.y, q0.y));
        const f2 t1 = mul2(pk(q0.w, q0.w), dy2);
        const f2 t0 = mul2(mul2(pk(q1.x, q1.x), dy2), dy2);
        float t1l, t1h, t0l, t0h;
        upk(t1, t1l, t1h);
        upk(t0, t0l, t0h);
        const f2 e0 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1l, t1l)), pk(t0l, t0l));
        const f2 e1 = fma2(dx2, fma2(pk(q0.z, q0.z), dx2, pk(t1h, t1h)), pk(t0h, t0h));
        float e00, e01, e10, e11;
        upk(e0, e00, e01);
        upk(e1, e10, e11);
        const f2 v0 = pk(ex2(e00), ex2(e01)), v1 = pk(ex2(e10), ex2(e11));
        r0 = add2(r0, v0);
        r1 = add2(r1, v1);
      }
    }
    float x, y;
    upk(add2(r0, r1), x, y);
    a0 = x + y;
  } else if (KIND == MIX_MMA) {
    // warp = one 8x8 region; lane (g = lane >> 2, t = lane & 3) evaluates records t and t + 4 of each
    // group of 8 at pixel column g, rows 0..7; colours accumulate on the tensor cores (tf32 hi/lo split).
    const int g = lane >> 2, t = lane & 3;
    const float nx = seed * g;
    f2 ny2[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) ny2[m] = pk(seed * (2 * m), seed * (2 * m + 1));
    float c[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m) c[m][0] = c[m][1] = c[m][2] = c[m][3] = 0.f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int u = 0; u < UNROLL / 4; ++u) {  // one group of 8 records = 512 evals = 4 "record-iterations" of 128
        const uint32_t ad = base + (((it * 2 + u) & 7) * 8 + t) * 32;
        uint32_t hi[2][8];
        float lo[2][8];
        uint32_t bf[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float4 q0 = lds128(ad + r * 128), q1 = lds128(ad + r * 128 + 16);
          const float dx = q0.x + nx;
          const float A = q0.z * dx * dx, Bx = q0.w * dx;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const f2 dy2 = add2(ny2[m], pk(q0.y, q0.y));
            const f2 e2 = fma2(dy2, fma2(pk(q1.x, q1.x), dy2, pk(Bx, Bx)), pk(A, A));
            float e0, e1;
            upk(e2, e0, e1);
            const float v0 = ex2(e0), v1 = ex2(e1);
            hi[r][2 * m] = __float_as_uint(v0) & 0xffffe000u;
            hi[r][2 * m + 1] = __float_as_uint(v1) & 0xffffe000u;
            float l0, l1;
            upk(add2(pk(v0, v1), pk(-__uint_as_float(hi[r][2 * m]), -__uint_as_float(hi[r][2 * m + 1]))), l0, l1);
            lo[r][2 * m] = l0;
            lo[r][2 * m + 1] = l1;
          }
          // B fragment: channel g >> 1 of this record, hi (g even) or lo (g odd) part
          float col;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(col) : "r"(ad + r * 128 + 20 + (g >> 1) * 4));
          const uint32_t ch = __float_as_uint(col) & 0xffffe000u;
          bf[r] = g >= 6 ? 0u : ((g & 1) ? __float_as_uint(col - __uint_as_float(ch)) : ch);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[m][0]), "+f"(c[m][1]), "+f"(c[m][2]), "+f"(c[m][3])
                       : "r"(hi[0][2 * m]), "r"(hi[0][2 * m + 1]), "r"(hi[1][2 * m]), "r"(hi[1][2 * m + 1]), "r"(bf[0]), "r"(bf[1]));
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[m][0]), "+f"(c[m][1]), "+f"(c[m][2]), "+f"(c[m][3])
                       : "r"(__float_as_uint(lo[0][2 * m])), "r"(__float_as_uint(lo[0][2 * m + 1])), "r"(__float_as_uint(lo[1][2 * m])),
                         "r"(__float_as_uint(lo[1][2 * m + 1])), "r"(bf[0]), "r"(bf[1]));
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) a0 += c[m][0] + c[m][1] + c[m][2] + c[m][3];
  } else if (KIND == FFMA2_BCAST) {
    const f2 qa = pk(seed, seed + 1), qb = pk(seed + 2, seed + 3);
    for (int it = 0; it < ITERS; ++it) {
      p0 = fma2(qa, pk(a0, a0), p0); p1 = fma2(qb, pk(a1, a1), p1); p2 = fma2(qa, pk(a2, a2), p2); p3 = fma2(qb, pk(a3, a3), p3);
      p4 = fma2(qa, pk(a4, a4), p4); p5 = fma2(qb, pk(a5, a5), p5); p6 = fma2(qa, pk(a6, a6), p6); p7 = fma2(qb, pk(a7, a7), p7);
    }
  } else {
    for (int it = 0; it < ITERS; ++it) {
      if (KIND == FFMA) {
        a0 = fmaf(a0, seed, seed); a1 = fmaf(a1, seed, seed); a2 = fmaf(a2, seed, seed); a3 = fmaf(a3, seed, seed);
        a4 = fmaf(a4, seed, seed); a5 = fmaf(a5, seed, seed); a6 = fmaf(a6, seed, seed); a7 = fmaf(a7, seed, seed);
      } else if (KIND == FFMA2) {
        p0 = fma2(p0, pc, pc); p1 = fma2(p1, pc, pc); p2 = fma2(p2, pc, pc); p3 = fma2(p3, pc, pc);
        p4 = fma2(p4, pc, pc); p5 = fma2(p5, pc, pc); p6 = fma2(p6, pc, pc); p7 = fma2(p7, pc, pc);
      } else if (KIND == MUFU) {
        a0 = ex2(a0); a1 = ex2(a1); a2 = ex2(a2); a3 = ex2(a3); a4 = ex2(a4); a5 = ex2(a5); a6 = ex2(a6); a7 = ex2(a7);
      } else if (KIND == SHFL) {
        a0 = __shfl_sync(0xffffffffu, a0, 3); a1 = __shfl_sync(0xffffffffu, a1, 5); a2 = __shfl_sync(0xffffffffu, a2, 7);
        a3 = __shfl_sync(0xffffffffu, a3, 9); a4 = __shfl_sync(0xffffffffu, a4, 3); a5 = __shfl_sync(0xffffffffu, a5, 5);
        a6 = __shfl_sync(0xffffffffu, a6, 7); a7 = __shfl_sync(0xffffffffu, a7, 9);
      } else if (KIND == LDS64_SAME) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          float x, y;
          asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(base + u * 32));
          a0 += x; a1 += y;
        }
      } else if (KIND == LDS32_SAME) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(base + u * 32));
          a0 += x;
        }
      } else {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const float4 v = lds128(addr + u * 1024);
          a0 += v.x; a1 += v.w;
        }
      }
    }
  }
  const long long t1 = clock64();
  float x, y;
  upk(add2(add2(p0, p1), add2(add2(p2, p3), add2(add2(p4, p5), add2(p6, p7)))), x, y);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + x + y;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(int warps) {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  long long* cyc;
  float* sink;
  cudaMalloc(&cyc, nsm * sizeof(long long));
  cudaMalloc(&sink, (size_t)nsm * 1024 * sizeof(float));
  bench<KIND><<<nsm, warps * 32>>>(cyc, sink, 0.f);
  bench<KIND><<<nsm, warps * 32>>>(cyc, sink, 0.f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", kNames[KIND], cudaGetErrorString(e)); return; }
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double per = (double)h[nsm / 2] / ((double)ITERS * UNROLL * warps);
  printf("%-44s warps/SM %2d  SM cycles per warp instruction %.3f  (x4 = per SMSP: %.2f)\n", kNames[KIND], warps, per, per * 4);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int warps : {16, 24, 32}) {
    run<FFMA>(warps); run<FFMA2>(warps); run<MUFU>(warps); run<SHFL>(warps);
    run<LDS128_SAME>(warps); run<LDS128_HALF>(warps); run<LDS128_QUARTER>(warps); run<LDS128_EIGHTH>(warps);
    run<LDS128_LANE>(warps); run<LDS64_SAME>(warps); run<LDS32_SAME>(warps);
    run<MIX1x2>(warps); run<MIX2x2>(warps); run<MIX2x2_NOCOL>(warps); run<MIX_MMA>(warps); run<FFMA2_BCAST>(warps); run<MIX2x2_SCOL>(warps); run<MIX2x2_SCOLEXP>(warps); run<MIX2x2_SCALAR>(warps);
  }
  return 0;
}
