// gsr_prepass.cuh -- O(N) set-up pipeline shared by forward and backward:
//
//   K1 gsr_bin_kernel     per Gaussian: exact dmax window /\ k-sigma box -> cull box, home bin,
//                         rank inside the bin (atomic), global reach statistics; also fills the
//                         pixel coordinate tables (the reference's double-precision rule).
//   K2 gsr_scan_kernel    exclusive scan of the bin histogram (single CTA).
//   K3 gsr_scatter_kernel counting-sort scatter: writes the 32 B raster record, the packed
//                         cull box and the original index at offset[bin] + rank.
//
// After K3 the Gaussians of one home bin are contiguous, bins are row-major, and the
// "large" Gaussians (cull box half-extent > GSR_LARGE_PX) form one extra bin at the end.
// Memory is bounded by sizes alone (no data-dependent list lengths, no host sync).
#pragma once
#include "gsr_common.cuh"

struct GsrWorkspace {
  int* bin_count;   // nb + 1        (zeroed per call, contiguous with stats)
  int* stats;       // 8 ints        [0] max ext_x (small), [1] max ext_y (small)
  int* bin_off;     // nb + 2        exclusive offsets; [nb] = start of large, [nb+1] = n_live
  float* px_tab;    // w
  float* py_tab;    // h
  uint2* box_tmp;   // s   (unsorted)
  int2* keyrank;    // s   (unsorted)  key = bin id, -1 = skipped
  GsrRec* rec;      // s   (sorted)
  uint2* box;       // s   (sorted)
  int* ids;         // s   (sorted -> original index)
  int nbx, nby, nb;
  size_t bytes;
};

static inline size_t gsr_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves the caller's workspace.  base may be NULL to compute the size only.
static inline GsrWorkspace gsr_carve(void* base, int s, int h, int w) {
  GsrWorkspace ws;
  ws.nbx = (w + GSR_BIN - 1) / GSR_BIN;
  ws.nby = (h + GSR_BIN - 1) / GSR_BIN;
  ws.nb = ws.nbx * ws.nby;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += gsr_align_up(bytes, 256);
    return (void*)r;
  };
  const size_t sn = (size_t)(s > 0 ? s : 1);
  // bin_count and stats must be contiguous: they are cleared by one memset.
  ws.bin_count = (int*)take(((size_t)ws.nb + 1 + 8) * sizeof(int));
  ws.stats = ws.bin_count ? ws.bin_count + ws.nb + 1 : nullptr;
  ws.bin_off = (int*)take(((size_t)ws.nb + 2) * sizeof(int));
  ws.px_tab = (float*)take((size_t)w * sizeof(float));
  ws.py_tab = (float*)take((size_t)h * sizeof(float));
  ws.box_tmp = (uint2*)take(sn * sizeof(uint2));
  ws.keyrank = (int2*)take(sn * sizeof(int2));
  ws.rec = (GsrRec*)take(sn * sizeof(GsrRec));
  ws.box = (uint2*)take(sn * sizeof(uint2));
  ws.ids = (int*)take(sn * sizeof(int));
  ws.bytes = off;
  return ws;
}

__global__ void __launch_bounds__(256)
gsr_bin_kernel(const float* __restrict__ sigmas, const float* __restrict__ coords,
               const float* __restrict__ colors, int s, int h, int w, float dmax, float ksigma,
               GsrWorkspace ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s) return;
  const float sx = __ldg(sigmas + 3 * (size_t)i + 0);
  const float sy = __ldg(sigmas + 3 * (size_t)i + 1);
  const float rho = __ldg(sigmas + 3 * (size_t)i + 2);
  const float x = __ldg(coords + 2 * (size_t)i + 0);
  const float y = __ldg(coords + 2 * (size_t)i + 1);
  const float cr = __ldg(colors + 3 * (size_t)i + 0);
  const float cg = __ldg(colors + 3 * (size_t)i + 1);
  const float cb = __ldg(colors + 3 * (size_t)i + 2);
  GsrSetup st = gsr_setup(sx, sy, rho, x, y, cr, cg, cb, h, w, dmax, ksigma, ws.px_tab, ws.py_tab);
  if (st.live) {
    const GsrRec r = gsr_make_rec(sx, sy, rho, x, y, cr, cg, cb);
    if (!(gsr_finite(r.a) && gsr_finite(r.b) && gsr_finite(r.c))) st.live = false;
  }
  int key = -1, rank = 0;
  if (st.live) {
    key = st.large ? ws.nb : st.bin_y * ws.nbx + st.bin_x;
    rank = atomicAdd(ws.bin_count + key, 1);
    ws.box_tmp[i] = gsr_box_pack(st.x0, st.x1, st.y0, st.y1, st.binds);
  }
  ws.keyrank[i] = make_int2(key, rank);
  // reach statistics: one atomic per warp, and only while the maximum still grows
  const bool small = st.live && !st.large;
  const unsigned act = __activemask();
  const int ex = __reduce_max_sync(act, small ? st.ext_x : 0);
  const int ey = __reduce_max_sync(act, small ? st.ext_y : 0);
  if ((threadIdx.x & 31) == (__ffs(act) - 1)) {
    if (ex > *(volatile int*)(ws.stats + 0)) atomicMax(ws.stats + 0, ex);
    if (ey > *(volatile int*)(ws.stats + 1)) atomicMax(ws.stats + 1, ey);
  }
}

// Pixel coordinate tables: the reference's rule (gs.cu:39,46), evaluated once per axis entry.
__global__ void __launch_bounds__(256) gsr_table_kernel(float* __restrict__ px_tab,
                                                        float* __restrict__ py_tab, int h, int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w) px_tab[i] = gsr_pix_coord(i, w);
  if (i < h) py_tab[i] = gsr_pix_coord(i, h);
}

// Exclusive scan of n = nb + 1 counters into n + 1 offsets; one CTA of 1024 threads.
__global__ void __launch_bounds__(1024) gsr_scan_kernel(const int* __restrict__ count,
                                                        int* __restrict__ off, int n) {
  __shared__ int warp_sums[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  // process in slabs of 1024*4 elements so that loads stay coalesced (int4 per thread)
  for (int base = 0; base < n; base += 4096) {
    const int i0 = base + tid * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? count[i0 + k] : 0;
    const int local = v[0] + v[1] + v[2] + v[3];
    int incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int ws = warp_sums[lane];
      int wi = ws;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      warp_sums[lane] = wi - ws;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    const int carry = carry_s;
    int run = carry + warp_sums[warp] + incl - local;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) off[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
    if (tid == 1023) carry_s = run;
    __syncthreads();
  }
  if (tid == 0) off[n] = carry_s;
}

__global__ void __launch_bounds__(256)
gsr_scatter_kernel(const float* __restrict__ sigmas, const float* __restrict__ coords,
                   const float* __restrict__ colors, int s, GsrWorkspace ws) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s) return;
  const int2 kr = ws.keyrank[i];
  if (kr.x < 0) return;
  const int dst = __ldg(ws.bin_off + kr.x) + kr.y;
  const float sx = __ldg(sigmas + 3 * (size_t)i + 0);
  const float sy = __ldg(sigmas + 3 * (size_t)i + 1);
  const float rho = __ldg(sigmas + 3 * (size_t)i + 2);
  const float x = __ldg(coords + 2 * (size_t)i + 0);
  const float y = __ldg(coords + 2 * (size_t)i + 1);
  const float cr = __ldg(colors + 3 * (size_t)i + 0);
  const float cg = __ldg(colors + 3 * (size_t)i + 1);
  const float cb = __ldg(colors + 3 * (size_t)i + 2);
  const GsrRec r = gsr_make_rec(sx, sy, rho, x, y, cr, cg, cb);
  float4* dr = reinterpret_cast<float4*>(ws.rec + dst);
  dr[0] = make_float4(r.x, r.y, r.a, r.b);
  dr[1] = make_float4(r.c, r.r, r.g, r.bl);
  ws.box[dst] = ws.box_tmp[i];
  ws.ids[dst] = i;
}
