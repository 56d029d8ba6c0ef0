// gsr_prepass.cuh -- O(N) set-up pipeline shared by forward and backward:
//
//   K1 gsr_bin_kernel     per Gaussian: exact dmax window /\ k-sigma box -> cull box, home bin,
//                         rank inside the bin (atomic), global reach statistics; also fills the
//                         pixel coordinate tables (the reference's double-precision rule).
//   K2 gsr_scan_kernel    exclusive scan of the bin histogram (decoupled look-back, one CTA per
//                         4096 bins).
//   K3 gsr_scatter_kernel counting-sort scatter: writes the 32 B raster record, the packed
//                         cull box and the original index at offset[bin] + rank.
//
// After K3 the Gaussians of one home bin are contiguous, bins are row-major, and the
// "large" Gaussians (cull box half-extent > GSR_LARGE_PX) form one extra bin at the end.
// Memory is bounded by sizes alone (no data-dependent list lengths, no host sync).
#pragma once
#include "gsr_common.cuh"

struct GsrWorkspace {
  int* bin_count;   // nb + 1        (zeroed per call, contiguous with stats and scan_state)
  int* stats;       // 8 ints        [0] max ext_x (small), [1] max ext_y (small)
  int* scan_state;  // 2 * nscan     per scan CTA: [2b] = ready flag, [2b+1] = CTA total
  int nscan;        // CTAs of the scan kernel
  size_t zero_bytes;  // bytes of the block cleared per call
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

constexpr int GSR_SCAN_CHUNK = 4096;  // counters per scan CTA (1024 threads x 4)

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
  // bin_count, stats and scan_state are contiguous: they are cleared by one memset.
  ws.nscan = (ws.nb + 1 + GSR_SCAN_CHUNK - 1) / GSR_SCAN_CHUNK;
  ws.zero_bytes = ((size_t)ws.nb + 1 + 8 + 2 * (size_t)ws.nscan) * sizeof(int);
  ws.bin_count = (int*)take(ws.zero_bytes);
  ws.stats = ws.bin_count ? ws.bin_count + ws.nb + 1 : nullptr;
  ws.scan_state = ws.bin_count ? ws.stats + 8 : nullptr;
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

// Exclusive scan of n = nb + 1 counters into n + 1 offsets.  One CTA of 1024 threads per
// GSR_SCAN_CHUNK counters; CTAs publish their totals (state[2b+1], then flag state[2b]) and sum
// the totals of their predecessors (decoupled look-back: totals do not depend on anything, so
// every flag is raised as soon as its CTA has been scheduled; CTAs are dispatched in index order,
// so a waiting CTA never starves the ones it waits for).  `state` is zero on entry.
__device__ __forceinline__ int gsr_block_sum_1024(int v, int* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if (lane == 0) warp_sums[warp] = v;
  __syncthreads();
  int t = warp_sums[lane];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(1024) gsr_scan_kernel(const int* __restrict__ count,
                                                        int* __restrict__ off, int n,
                                                        int* state) {
  __shared__ int warp_sums[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const int i0 = b * GSR_SCAN_CHUNK + tid * 4;
  int v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? count[i0 + k] : 0;
  const int local = v[0] + v[1] + v[2] + v[3];
  // inclusive scan of the thread sums inside the CTA
  int incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int ws = warp_sums[lane], wi = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    warp_sums[lane] = wi - ws;  // exclusive prefix of the warp totals
    if (lane == 31) {           // wi = CTA total: publish it
      *(volatile int*)(state + 2 * b + 1) = wi;
      __threadfence();
      *(volatile int*)(state + 2 * b) = 1;
    }
  }
  __syncthreads();
  const int in_cta = warp_sums[warp] + incl - local;
  __syncthreads();
  // look back: sum the totals of CTAs 0 .. b-1
  int carry = 0;
  for (int j = tid; j < b; j += 1024) {
    while (*(volatile int*)(state + 2 * j) == 0) {
    }
    __threadfence();
    carry += *(volatile int*)(state + 2 * j + 1);
  }
  carry = gsr_block_sum_1024(carry, warp_sums);
  int run = carry + in_cta;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (i0 + k < n) off[i0 + k] = run;
    run += v[k];
  }
  if (b == gridDim.x - 1 && tid == 1023) off[n] = run;
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
