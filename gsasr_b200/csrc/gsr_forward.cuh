// gsr_forward.cuh -- forward raster kernel (sm_100a).
//
// Replaces _gs_render_cuda (utils/gs_cuda_dmax/gs.cu:7-64; utils/gs_cuda/gs.cu:9-61).
// The reference scatters: one thread per Gaussian, three global atomics per (Gaussian,pixel).
// Here the image is cut into 32x32 tiles, one CTA per tile, 16 warps each owning an 8x8 region
// (two horizontally adjacent pixels per lane, accumulated in registers, written once).
//
// Per CTA:
//   stage A  (one thread per candidate)  candidates are the Gaussians of the home bins within
//            reach of the tile -- contiguous runs of the sorted arrays, one run per bin row --
//            plus the "large" list.  Cull box vs tile, then an ellipse-vs-region mask
//            (gsr_region_mask).  Survivors are copied to shared memory and their slot is
//            appended to the list of every region they touch.
//   stage C  (one warp per region)  walk the region's list; per Gaussian 2 LDS.128 (broadcast),
//            then per pixel: 1 FADD + 2 FFMA + MUFU.EX2 + 3 FFMA.
// The bound is the MUFU pipe (16 ex2/clk/SM): see DESIGN.md.
#pragma once
#include "gsr_prepass.cuh"

constexpr int GSR_FWD_THREADS = 512;
constexpr int GSR_FWD_WARPS = GSR_FWD_THREADS / 32;
constexpr int GSR_FWD_CAP = 1024;  // survivor slots per flush
constexpr int GSR_FWD_MAXRUNS = 2 * ((GSR_LARGE_PX + GSR_BIN - 1) / GSR_BIN) + GSR_TILE / GSR_BIN + 2;
static_assert(GSR_FWD_WARPS == (GSR_TILE / GSR_REGION) * (GSR_TILE / GSR_REGION), "one warp per region");
static_assert(GSR_FWD_MAXRUNS <= 32, "run table is built by one warp");

struct GsrFwdSmem {
  float4 rec[GSR_FWD_CAP * 2];
  uint2 box[GSR_FWD_CAP];
  uint16_t list[GSR_FWD_WARPS][GSR_FWD_CAP];
  int list_n[GSR_FWD_WARPS];
  int run_start[GSR_FWD_MAXRUNS];
  int run_prefix[GSR_FWD_MAXRUNS + 1];
  int wcnt[2][GSR_FWD_WARPS];
  int nruns;
};

__device__ __forceinline__ float gsr_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct GsrFwdArgs {
  const GsrRec* rec;
  const uint2* box;
  const int* bin_off;
  const int* stats;
  const float* px_tab;
  const float* py_tab;
  float* img;
  int h, w, nbx, nby, nb;
  float ecut;
  uint32_t flags;
};

// Builds the table of candidate runs for a pixel rectangle [x0,x1]x[y0,y1] (inclusive):
// one run per bin row within reach + the large list.  Executed by warp 0.
__device__ __forceinline__ void gsr_build_runs(const int* __restrict__ bin_off,
                                               const int* __restrict__ stats, int nbx, int nby,
                                               int nb, int x0, int x1, int y0, int y1, int lane,
                                               int* run_start, int* run_prefix, int* nruns) {
  const int ext_x = __ldg(stats + 0), ext_y = __ldg(stats + 1);
  const int bx_lo = max(x0 - ext_x, 0) / GSR_BIN;
  const int bx_hi = min((x1 + ext_x) / GSR_BIN, nbx - 1);
  const int by_lo = max(y0 - ext_y, 0) / GSR_BIN;
  const int by_hi = min((y1 + ext_y) / GSR_BIN, nby - 1);
  const int nrows = by_hi - by_lo + 1;
  int st = 0, len = 0;
  if (lane < nrows) {
    const int row = (by_lo + lane) * nbx;
    st = __ldg(bin_off + row + bx_lo);
    len = __ldg(bin_off + row + bx_hi + 1) - st;
  } else if (lane == nrows) {
    st = __ldg(bin_off + nb);
    len = __ldg(bin_off + nb + 1) - st;
  }
  // compact away empty runs and prefix-sum the lengths
  const unsigned have = __ballot_sync(0xffffffffu, len > 0);
  int incl = len;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const int pos = __popc(have & ((1u << lane) - 1u));
  if (len > 0) {
    run_start[pos] = st;
    run_prefix[pos] = incl - len;
  }
  const int n = __popc(have);
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  if (lane == 0) {
    run_prefix[n] = total;
    *nruns = n;
  }
}

__global__ void __launch_bounds__(GSR_FWD_THREADS, 2) gsr_forward_kernel(GsrFwdArgs p) {
  extern __shared__ __align__(16) unsigned char gsr_smem_raw[];
  GsrFwdSmem& sm = *reinterpret_cast<GsrFwdSmem*>(gsr_smem_raw);
  constexpr int NR = GSR_TILE / GSR_REGION;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx0 = blockIdx.x * GSR_TILE, ty0 = blockIdx.y * GSR_TILE;

  // this thread's two pixels
  const int wi0 = tx0 + (warp % NR) * GSR_REGION + (lane & 3) * 2;
  const int hi = ty0 + (warp / NR) * GSR_REGION + (lane >> 2);
  const float px0 = __ldg(p.px_tab + min(wi0, p.w - 1));
  const float px1 = __ldg(p.px_tab + min(wi0 + 1, p.w - 1));
  const float py = __ldg(p.py_tab + min(hi, p.h - 1));
  float r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;

  if (warp == 0)
    gsr_build_runs(p.bin_off, p.stats, p.nbx, p.nby, p.nb, tx0, tx0 + GSR_TILE - 1, ty0,
                   ty0 + GSR_TILE - 1, lane, sm.run_start, sm.run_prefix, &sm.nruns);
  if (tid < GSR_FWD_WARPS) sm.list_n[tid] = 0;
  __syncthreads();
  int nsurv = 0;  // survivors waiting in shared memory (same value in every thread)
  const int nruns = sm.nruns;
  const int total = sm.run_prefix[nruns];

  for (int base = 0; base < total; base += GSR_FWD_THREADS) {
    // ---------------- stage A: cull one candidate per thread ----------------
    const int cnd = base + tid;
    uint32_t mask = 0;
    uint2 pb = make_uint2(0, 0);
    float4 q0, q1;
    if (cnd < total) {
      int r = 0;
      while (cnd >= sm.run_prefix[r + 1]) ++r;
      const int idx = sm.run_start[r] + (cnd - sm.run_prefix[r]);
      pb = __ldg(p.box + idx);
      int bx0, bx1, by0, by1;
      bool binds;
      gsr_box_unpack(pb, bx0, bx1, by0, by1, binds);
      if (bx1 >= tx0 && bx0 < tx0 + GSR_TILE && by1 >= ty0 && by0 < ty0 + GSR_TILE) {
        const float4* rp = reinterpret_cast<const float4*>(p.rec + idx);
        q0 = __ldg(rp);
        q1 = __ldg(rp + 1);
        GsrRec g;
        g.x = q0.x; g.y = q0.y; g.a = q0.z; g.b = q0.w;
        g.c = q1.x; g.r = q1.y; g.g = q1.z; g.bl = q1.w;
        mask = gsr_region_mask(g, bx0, bx1, by0, by1, tx0, ty0, p.h, p.w, p.ecut);
      }
    }
    // Deterministic slot assignment: per-warp survivor counts -> block prefix.  The count
    // buffers alternate between chunks, so one barrier per chunk is enough.
    const unsigned bal = __ballot_sync(0xffffffffu, mask != 0);
    int* wcnt = sm.wcnt[(base / GSR_FWD_THREADS) & 1];
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    {
      const int mine = lane < GSR_FWD_WARPS ? wcnt[lane] : 0;
      int incl = mine;
#pragma unroll
      for (int d = 1; d < GSR_FWD_WARPS; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const int before = __shfl_sync(0xffffffffu, incl - mine, warp);
      const int chunk_total = __shfl_sync(0xffffffffu, incl, GSR_FWD_WARPS - 1);
      if (mask) {
        const int slot = nsurv + before + __popc(bal & ((1u << lane) - 1u));
        sm.rec[2 * slot] = q0;
        sm.rec[2 * slot + 1] = q1;
        sm.box[slot] = pb;
        const uint16_t entry = (uint16_t)(slot | ((pb.x & 0x8000u) ? 0x8000 : 0));
        while (mask) {
          const int rg = __ffs(mask) - 1;
          mask &= mask - 1;
          const int pos = atomicAdd(&sm.list_n[rg], 1);
          sm.list[rg][pos] = entry;
        }
      }
      nsurv += chunk_total;
    }
    const bool last = base + GSR_FWD_THREADS >= total;
    if (!last && nsurv + GSR_FWD_THREADS <= GSR_FWD_CAP) continue;
    __syncthreads();

    // ---------------- stage C: every warp walks its region's list ----------------
    const int n = sm.list_n[warp];
    const uint16_t* mylist = sm.list[warp];
#pragma unroll 2
    for (int i = 0; i < n; ++i) {
      const uint32_t entry = mylist[i];
      const int slot = entry & 0x3ff;
      const float4 a0 = sm.rec[2 * slot];
      const float4 a1 = sm.rec[2 * slot + 1];
      const float dy = py - a0.y;
      const float t1 = a0.w * dy;
      const float t0 = a1.x * dy * dy;
      const float dx0 = px0 - a0.x;
      const float dx1 = px1 - a0.x;
      const float e0 = fmaf(dx0, fmaf(a0.z, dx0, t1), t0);
      const float e1 = fmaf(dx1, fmaf(a0.z, dx1, t1), t0);
      float v0 = gsr_ex2(e0);
      float v1 = gsr_ex2(e1);
      if (entry & 0x8000u) {  // dmax window cuts this Gaussian: exact inclusion test
        int bx0, bx1, by0, by1;
        bool binds;
        gsr_box_unpack(sm.box[slot], bx0, bx1, by0, by1, binds);
        const bool iny = hi >= by0 && hi <= by1;
        if (!(iny && wi0 >= bx0 && wi0 <= bx1)) v0 = 0.f;
        if (!(iny && wi0 + 1 >= bx0 && wi0 + 1 <= bx1)) v1 = 0.f;
      }
      r0 = fmaf(v0, a1.y, r0);
      g0 = fmaf(v0, a1.z, g0);
      b0 = fmaf(v0, a1.w, b0);
      r1 = fmaf(v1, a1.y, r1);
      g1 = fmaf(v1, a1.z, g1);
      b1 = fmaf(v1, a1.w, b1);
    }
    __syncthreads();
    if (tid < GSR_FWD_WARPS) sm.list_n[tid] = 0;
    nsurv = 0;
  }

  // ---------------- write-out ----------------
  if (hi < p.h) {
    const bool over = (p.flags & 1u) != 0;
    if (p.flags & 2u) {  // CHW
      const size_t plane = (size_t)p.h * p.w;
      float* o = p.img + (size_t)hi * p.w + wi0;
      if (wi0 < p.w) {
        o[0] = over ? r0 : o[0] + r0;
        o[plane] = over ? g0 : o[plane] + g0;
        o[2 * plane] = over ? b0 : o[2 * plane] + b0;
      }
      if (wi0 + 1 < p.w) {
        o[1] = over ? r1 : o[1] + r1;
        o[plane + 1] = over ? g1 : o[plane + 1] + g1;
        o[2 * plane + 1] = over ? b1 : o[2 * plane + 1] + b1;
      }
    } else {  // HWC: 6 contiguous floats per lane
      float* o = p.img + ((size_t)hi * p.w + wi0) * 3;
      if (wi0 < p.w) {
        o[0] = over ? r0 : o[0] + r0;
        o[1] = over ? g0 : o[1] + g0;
        o[2] = over ? b0 : o[2] + b0;
      }
      if (wi0 + 1 < p.w) {
        o[3] = over ? r1 : o[3] + r1;
        o[4] = over ? g1 : o[4] + g1;
        o[5] = over ? b1 : o[5] + b1;
      }
    }
  }
}
