// gsr_prepass.cuh -- O(N) set-up pipelines shared by forward and backward.
//
// Region-bucket pipeline (forward fast path), ONE pass over the Gaussians:
//   T1 gsr_region_build2_kernel per Gaussian: exact dmax window /\ k-sigma box -> cull box, raster
//                               record (written in input order); then for every 16x8-pixel region
//                               the ellipse touches a 4-byte entry {index | cell mask << 23 | binds << 31}
//                               is appended to the region's bucket -- the mask says which of the region's
//                               eight 4x4-pixel cells the ellipse reaches.  Optionally fed with the RAW head
//                               output: the activations and the unit mapping of the front end are then
//                               applied here (gsr_map_one).
//   Every region owns a fixed-capacity bucket (GSR_ENTRIES_PER_GAUSSIAN N / regions + 32 entries: GSASR
//   emits its Gaussians on a regular grid, utils/fea2gs.py:553-563, so the load per region is uniform); an
//   entry that does not fit raises the overflow flag and the forward falls back to the home-bin pipeline.
//   The forward kernel then needs no culling at all: a warp per region streams its bucket.
//
// Home-bin pipeline (backward; forward fallback when the tile lists overflow their capacity):
//   K1 gsr_bin_kernel     per Gaussian: cull box, home bin, rank inside the bin (atomic), reach.
//   K2 gsr_scan_kernel    exclusive scan of the bin histogram.
//   K3 gsr_scatter_kernel counting-sort scatter: 32 B record, packed cull box, original index at
//                         offset[bin] + rank.  Bins are row-major; Gaussians whose cull box reaches
//                         further than GSR_LARGE_PX form one extra bin at the end.
//
// Memory is bounded by sizes alone (no data-dependent allocation, no host sync): the choice between
// the two forward paths is made ON THE DEVICE through stats[GSR_STAT_OVERFLOW]; kernels of the path
// not taken return at once (`guard`).  In a one-call forward K1..K3 run as phases of the fallback raster
// kernel itself (gsr_forward_fallback_kernel, gsr_forward.cuh), so the normal path pays for ONE idle launch.
#pragma once
#include "gsr_common.cuh"
struct gsr_window;

// One sample of a padded (ragged) batch: its own image size and window inside the common (hmax, wmax)
// block of the stacked canvas, and the ratios between the canvas' and its own coordinate normalisation
// (pixel i sits at 2i/(n-1)-1 for an n-pixel axis, so px_own = ax * px_canvas + (ax - 1), ax = (W-1)/(w-1)).
struct GsrBDesc {
  int h, w;
  float dmax;
  float step;     // step size of the sample (fused front end only)
  double ax, ay;  // double: a float ratio would move the centres by 1e-4 pixel
};
constexpr int GSR_BDESC_MAX = 1024;  // samples per stacked launch

constexpr int GSR_STAT_EXT_X = 0, GSR_STAT_EXT_Y = 1, GSR_STAT_OVERFLOW = 2, GSR_STAT_ENTRIES = 3;
// bucket capacity per region = 28 * N / regions + 32: a Gaussian of the x4 head touches 4.4 of the 16x8
// regions on average, one of the x8 head 12.3 (5 sigma = 16.7 px on average, 33 px at most); 4 bytes per slot
constexpr int GSR_ENTRIES_PER_GAUSSIAN = 28;
constexpr int GSR_STAT_KSIGMA = 6;                   // effective k-sigma of the home-bin set-up (float bits), for the backward
constexpr int GSR_STAT_UNSORTED = 7;                 // deterministic mode: a bucket was too long to be sorted
constexpr int GSR_STAT_BARRIER = 8;                  // arrival counter of the fallback kernel's grid barriers
constexpr int GSR_STAT_UNIT = 4, GSR_STAT_DONE = 5;  // work counter / finished-warp counter of the raster kernel

struct GsrWorkspace {
  // ---- one block, cleared per call ----
  int* bin_count;    // nb + 1
  int* stats;        // 16 ints, see GSR_STAT_*
  int* scan_state;   // 2 * nscan     look-back state of the bin scan
  int* reg_count;    // nreg      entries appended to each region's bucket (may exceed reg_cap)
  size_t zero_bytes;
  // ---- tile-list pipeline ----
  GsrRec* rec_in;    // s        records, input order
  uint2* box_in;     // s        packed cull boxes, input order (only read for window-binding ones)
  uint32_t* entries; // nreg * reg_cap   Gaussian index | cell mask << 23 | binds << 31
  int nrx, nry, nreg, reg_cap;  // GSR_RGW x GSR_RGH pixel regions
  // ---- home-bin pipeline ----
  int* bin_off;      // nb + 2   exclusive offsets; [nb] = start of large, [nb+1] = n_live
  uint2* box_tmp;    // s   (unsorted)
  int2* keyrank;     // s   (unsorted)  key = bin id, -1 = skipped
  GsrRec* rec;       // s   (sorted)
  uint2* box;        // s   (sorted)
  int* ids;
  float* mom;  // (s, 8) moment rows of the region backward          // s   (sorted -> original index)
  int nbx, nby, nb, nscan;
  // ---- both ----
  float* px_tab;     // w
  float* py_tab;     // h
  size_t bytes;
  int hf, row0;      // row-band view: the image is rows [row0, row0 + h) of an hf-row image (hf = 0: whole)
  const struct gsr_window* win;  // host-side only: destination window of gsr_forward_window (NULL: plain image)
  int bn, bhs;       // uniform batch: the image is a stack of samples, bhs rows each, bn Gaussians each (0: single)
  GsrBDesc* bdesc;   // GSR_BDESC_MAX descriptors; read by the kernels only when `ragged`
  int ragged;        // padded batch: per-sample (h, w, dmax) from bdesc, records rescaled to canvas coordinates
};

constexpr int GSR_SCAN_CHUNK = 4096;  // counters per scan CTA (1024 threads x 4)

static inline size_t gsr_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves the caller's workspace.  base may be NULL to compute the size only.
static inline GsrWorkspace gsr_carve(void* base, int s, int h, int w) {
  GsrWorkspace ws;
  ws.nbx = (w + GSR_BIN - 1) / GSR_BIN;
  ws.nby = (h + GSR_BIN - 1) / GSR_BIN;
  ws.nb = ws.nbx * ws.nby;
  ws.nrx = (w + GSR_RGW - 1) / GSR_RGW;
  ws.nry = (h + GSR_RGH - 1) / GSR_RGH;
  ws.nreg = ws.nrx * ws.nry;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += gsr_align_up(bytes, 256);
    return (void*)r;
  };
  const size_t sn = (size_t)(s > 0 ? s : 1);
  ws.nscan = (ws.nb + 1 + GSR_SCAN_CHUNK - 1) / GSR_SCAN_CHUNK;
  ws.zero_bytes = ((size_t)ws.nb + 1 + 16 + 2 * (size_t)ws.nscan + (size_t)ws.nreg) * sizeof(int);
  ws.bin_count = (int*)take(ws.zero_bytes);
  ws.stats = ws.bin_count ? ws.bin_count + ws.nb + 1 : nullptr;
  ws.scan_state = ws.bin_count ? ws.stats + 16 : nullptr;
  ws.reg_count = ws.bin_count ? ws.scan_state + 2 * ws.nscan : nullptr;
  ws.rec_in = (GsrRec*)take(sn * sizeof(GsrRec));
  ws.box_in = (uint2*)take(sn * sizeof(uint2));
  const size_t per_reg = ((size_t)GSR_ENTRIES_PER_GAUSSIAN * (size_t)(s > 0 ? s : 0) + ws.nreg - 1) / ws.nreg + 32;
  ws.reg_cap = (int)(per_reg > 0x3fffffffu ? 0x3fffffffu : per_reg);
  ws.entries = (uint32_t*)take((size_t)ws.nreg * (size_t)ws.reg_cap * sizeof(uint32_t));
  ws.bin_off = (int*)take(((size_t)ws.nb + 2) * sizeof(int));
  ws.px_tab = (float*)take((size_t)w * sizeof(float));
  ws.py_tab = (float*)take((size_t)h * sizeof(float));
  ws.bdesc = (GsrBDesc*)take((size_t)GSR_BDESC_MAX * sizeof(GsrBDesc));
  ws.box_tmp = (uint2*)take(sn * sizeof(uint2));
  ws.keyrank = (int2*)take(sn * sizeof(int2));
  ws.rec = (GsrRec*)take(sn * sizeof(GsrRec));
  ws.box = (uint2*)take(sn * sizeof(uint2));
  ws.ids = (int*)take(sn * sizeof(int));
  ws.mom = (float*)take(sn * 8 * sizeof(float));
  ws.bytes = off;
  ws.hf = 0;
  ws.row0 = 0;
  ws.bn = 0;
  ws.bhs = 0;
  ws.ragged = 0;
  ws.win = nullptr;
  return ws;
}

// A kernel of the path not taken returns at once: guard == nullptr means "always run".
__device__ __forceinline__ bool gsr_guard_skip(const int* guard, int want) {
  return guard != nullptr && *(const volatile int*)guard != want;
}


// What a Gaussian's sample looks like: the image it is set up in (hl x wl, dmax), the row offset of its
// block in the stacked canvas, and the coordinate ratios of a padded batch (1 otherwise).
struct GsrSampleView {
  int hl, wl, yoff;
  float dmax;
  double ax, ay;
  const float* px_tab;
  const float* py_tab;
};
// RAGGED is a compile-time switch: the common (single image / uniform batch) instantiations of the set-up
// kernels carry none of the padded-batch code (it cost 9 us at HL when it was a run-time branch).
template <bool RAGGED>
__device__ __forceinline__ GsrSampleView gsr_sample_view(const GsrWorkspace& ws, int i, int h, int w, float dmax) {
  GsrSampleView v;
  v.hl = ws.bn > 0 ? ws.bhs : h;
  v.wl = w;
  v.yoff = ws.bn > 0 ? (i / ws.bn) * ws.bhs : 0;
  v.dmax = dmax;
  v.ax = v.ay = 1.0;
  v.px_tab = ws.px_tab;
  v.py_tab = ws.py_tab;
  if (RAGGED) {
    const int b = i / ws.bn;
    const GsrBDesc d = ws.bdesc[b < GSR_BDESC_MAX ? b : GSR_BDESC_MAX - 1];  // (threads past s: unused)
    v.hl = d.h;
    v.wl = d.w;
    v.dmax = d.dmax;
    v.ax = d.ax;
    v.ay = d.ay;
    v.px_tab = v.py_tab = nullptr;  // the canvas' tables are not the sample's: evaluate the rule directly
  }
  return v;
}
// Padded batch: a sample whose width / height is not a multiple of the region size shares its last regions
// with padding pixels, which must stay 0: Gaussians whose box reaches those regions take the exact per-pixel
// box test.
template <bool RAGGED>
__device__ __forceinline__ bool gsr_edge_binds(const GsrSampleView& v, const GsrSetup& st) {
  return RAGGED && (((v.wl % GSR_RGW) != 0 && st.x1 / GSR_RGW == (v.wl - 1) / GSR_RGW) ||
                       ((v.hl % GSR_RGH) != 0 && (st.y1 - v.yoff) / GSR_RGH == (v.hl - 1) / GSR_RGH));
}
// Record in CANVAS coordinates: d_own = a_ * d_canvas, so x_c = (x + 1) / ax - 1 and the conic scales.
__device__ __forceinline__ void gsr_rescale_rec(GsrRec& r, const GsrSampleView& v) {
  if (v.ax == 1.0 && v.ay == 1.0) return;
  r.x = (float)(((double)r.x + 1.0) / v.ax - 1.0);
  r.y = (float)(((double)r.y + 1.0) / v.ay - 1.0);
  r.a = (float)((double)r.a * v.ax * v.ax);
  r.b = (float)((double)r.b * v.ax * v.ay);
  r.c = (float)((double)r.c * v.ay * v.ay);
}

template <bool RAGGED>
__device__ __forceinline__ void gsr_bin_one(const float* __restrict__ sigmas,
                                            const float* __restrict__ coords,
                                            const float* __restrict__ colors, int i, int h, int w,
                                            float dmax, float ksigma, const GsrWorkspace& ws) {
  const float sx = __ldg(sigmas + 3 * (size_t)i + 0);
  const float sy = __ldg(sigmas + 3 * (size_t)i + 1);
  const float rho = __ldg(sigmas + 3 * (size_t)i + 2);
  const float x = __ldg(coords + 2 * (size_t)i + 0);
  const float y = __ldg(coords + 2 * (size_t)i + 1);
  const float cr = __ldg(colors + 3 * (size_t)i + 0);
  const float cg = __ldg(colors + 3 * (size_t)i + 1);
  const float cb = __ldg(colors + 3 * (size_t)i + 2);
  // batches: set up in the sample's own image, then move to its block of rows of the stack
  const GsrSampleView sv = gsr_sample_view<RAGGED>(ws, i, h, w, dmax);
  const int yoff = sv.yoff;
  GsrSetup st = gsr_setup(sx, sy, rho, x, y, cr, cg, cb, sv.hl, sv.wl, sv.dmax, ksigma, sv.px_tab, sv.py_tab, ws.hf,
                          ws.row0);
  if (st.live) {
    const GsrRec r = gsr_make_rec(sx, sy, rho, x, y, cr, cg, cb);
    if (!(gsr_finite(r.a) && gsr_finite(r.b) && gsr_finite(r.c))) st.live = false;
    st.y0 += yoff;
    st.y1 += yoff;
    st.bin_y += yoff / GSR_BIN;
    if (gsr_edge_binds<RAGGED>(sv, st)) st.binds = true;
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
    if (ex > *(volatile int*)(ws.stats + 0)) atomicMax(ws.stats + GSR_STAT_EXT_X, ex);
    if (ey > *(volatile int*)(ws.stats + 1)) atomicMax(ws.stats + GSR_STAT_EXT_Y, ey);
  }
}

// Grid-stride: a guarded launch of the path not taken uses a small grid and costs ~2 us.
template <bool RAGGED>
__global__ void __launch_bounds__(256)
gsr_bin_kernel(const float* __restrict__ sigmas, const float* __restrict__ coords,
               const float* __restrict__ colors, int s, int h, int w, float dmax, float ksigma,
               GsrWorkspace ws, const int* guard, int want) {
  if (gsr_guard_skip(guard, want)) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) ws.stats[GSR_STAT_KSIGMA] = __float_as_int(ksigma);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s; i += gridDim.x * blockDim.x)
    gsr_bin_one<RAGGED>(sigmas, coords, colors, i, h, w, dmax, ksigma, ws);
}

// Pixel coordinate tables: the reference's rule (gs.cu:39,46), evaluated once per axis entry -- and the clearing
// of the workspace's counter block (zero16 x 16 bytes from `zero`), which would otherwise be a launch of its own.
__global__ void __launch_bounds__(256) gsr_table_kernel(float* __restrict__ px_tab,
                                                        float* __restrict__ py_tab, int h, int w,
                                                        int hf, int row0, int bhs, uint4* __restrict__ zero,
                                                        size_t zero16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w) px_tab[i] = gsr_pix_coord(i, w);
  if (i < h) py_tab[i] = bhs > 0 ? gsr_pix_coord(i % bhs, bhs) : hf > 0 ? gsr_pix_coord(i + row0, hf) : gsr_pix_coord(i, h);
  for (size_t k = (size_t)i; k < zero16; k += (size_t)gridDim.x * blockDim.x) zero[k] = make_uint4(0u, 0u, 0u, 0u);
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

// Optional extras: `cur` receives a copy of the offsets (fill cursors); when cap >= 0 the last CTA
// stores the grand total in stats[GSR_STAT_ENTRIES] and raises stats[GSR_STAT_OVERFLOW] if it
// exceeds cap.
__global__ void __launch_bounds__(1024) gsr_scan_kernel(const int* __restrict__ count,
                                                        int* __restrict__ off, int* __restrict__ cur,
                                                        int n, int* state, int cap, int* stats,
                                                        const int* guard, int want) {
  if (gsr_guard_skip(guard, want)) return;
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
    if (i0 + k < n) {
      off[i0 + k] = run;
      if (cur) cur[i0 + k] = run;
    }
    run += v[k];
  }
  if (b == gridDim.x - 1 && tid == 1023) {
    off[n] = run;
    if (cap >= 0) {
      stats[GSR_STAT_ENTRIES] = run;
      if (run > cap) stats[GSR_STAT_OVERFLOW] = 1;
    }
  }
}

__device__ __forceinline__ void gsr_scatter_one(const float* __restrict__ sigmas,
                                                const float* __restrict__ coords,
                                                const float* __restrict__ colors, int i,
                                                const GsrWorkspace& ws) {
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
  GsrRec r = gsr_make_rec(sx, sy, rho, x, y, cr, cg, cb);
  if (ws.ragged) gsr_rescale_rec(r, gsr_sample_view<true>(ws, i, 0, 0, 0.f));  // padded batch: canvas coordinates
  float4* dr = reinterpret_cast<float4*>(ws.rec + dst);
  dr[0] = make_float4(r.x, r.y, r.a, r.b);
  dr[1] = make_float4(r.c, r.r, r.g, r.bl);
  ws.box[dst] = ws.box_tmp[i];
  ws.ids[dst] = i;
}

__global__ void __launch_bounds__(256)
gsr_scatter_kernel(const float* __restrict__ sigmas, const float* __restrict__ coords,
                   const float* __restrict__ colors, int s, GsrWorkspace ws, const int* guard,
                   int want) {
  if (gsr_guard_skip(guard, want)) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s; i += gridDim.x * blockDim.x)
    gsr_scatter_one(sigmas, coords, colors, i, ws);
}

// ---- the home-bin pipeline as phases of ONE resident grid (the forward's fallback kernel) -------------------
// Barrier across a grid whose CTAs are all resident (the caller sizes the grid by occupancy): arrival counter in
// the cleared block of the workspace, `target` = arrivals expected so far (phase number x gridDim.x).
__device__ __forceinline__ void gsr_grid_barrier(int* counter, int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1);
    while (*(volatile int*)counter < target) {
    }
    __threadfence();
  }
  __syncthreads();
}
// Exclusive scan of n counters into n + 1 offsets by the whole grid (256-thread CTAs), two phases around one
// barrier: chunk totals, then offsets.  `part` holds one int per GSR_SCAN_CHUNK counters.
__device__ __forceinline__ void gsr_grid_scan(const int* __restrict__ count, int* __restrict__ off, int n, int* part,
                                              int* barrier, int& phase) {
  __shared__ int red[256 / 32 + 1];
  constexpr int PER = GSR_SCAN_CHUNK / 256;  // counters per thread
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nchunks = (n + GSR_SCAN_CHUNK - 1) / GSR_SCAN_CHUNK;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    int v = 0;
    for (int k = 0; k < PER; ++k) {
      const int i = c * GSR_SCAN_CHUNK + tid * PER + k;
      v += i < n ? count[i] : 0;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int k = 0; k < 256 / 32; ++k) t += red[k];
      part[c] = t;
    }
    __syncthreads();
  }
  gsr_grid_barrier(barrier, (++phase) * gridDim.x);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    int carry = 0;
    for (int j = tid; j < c; j += 256) carry += *(volatile int*)(part + j);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) carry += __shfl_xor_sync(0xffffffffu, carry, d);
    if (lane == 0) red[warp] = carry;
    __syncthreads();
    carry = 0;
    for (int k = 0; k < 256 / 32; ++k) carry += red[k];
    __syncthreads();
    // this thread's PER counters, then an exclusive scan of the thread sums over the CTA
    int loc[PER], sum = 0;
    for (int k = 0; k < PER; ++k) {
      const int i = c * GSR_SCAN_CHUNK + tid * PER + k;
      loc[k] = i < n ? count[i] : 0;
      sum += loc[k];
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    int wpre = 0;
    for (int k = 0; k < warp; ++k) wpre += red[k];
    int run = carry + wpre + incl - sum;
    for (int k = 0; k < PER; ++k) {
      const int i = c * GSR_SCAN_CHUNK + tid * PER + k;
      if (i < n) off[i] = run;
      run += loc[k];
      if (i == n - 1) off[n] = run;
    }
    __syncthreads();
  }
}

// ---- region-bucket pipeline ---------------------------------------------------------------------
// Front end, folded in (utils/gaussian_splatting.py:174-180 activations, :121-123 unit / coordinate mapping of
// rendering_cuda_dmax): raw (s,9) = (sx, sy, rho, alpha, r, g, b, mu_x, mu_y) ->
//   sigmas = (sy/step*2/(w-1), sx/step*2/(h-1), rho)     NOTE the x/y swap (:121)
//   coords = ((mu*2-1) + 1 - 1/n) * n / (n-1) - 1         (:122-123)
//   colors = sigmoid(rgb) * sigmoid(alpha)                (:177-180)
// The arithmetic follows what PyTorch executes on the INFERENCE path (inference_paper.py:113-131: sr_size and
// scale_modify are CPU tensors, so CUDA tensor / CPU-scalar divisions run as a multiplication by the fp32
// reciprocal); every operation is individually rounded (no FMA contraction) so the mapped parameters agree
// with the unfused path to the last bit or one ulp.
__device__ __forceinline__ float gsr_sigmoid(float x) {
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
struct GsrMapped {
  float sx, sy, rho, x, y, cr, cg, cb;
};
__device__ __forceinline__ GsrMapped gsr_map_one(const float* __restrict__ p, int h, int w, float step) {
  const float inv_step = __fdiv_rn(1.0f, step);
  const float inv_w1 = __fdiv_rn(1.0f, (float)(w - 1)), inv_h1 = __fdiv_rn(1.0f, (float)(h - 1));
  const float inv_w = __fdiv_rn(1.0f, (float)w), inv_h = __fdiv_rn(1.0f, (float)h);
  const float sgx = __fadd_rn(__fmul_rn(0.99999f, gsr_sigmoid(__ldg(p + 0))), 1e-6f);
  const float sgy = __fadd_rn(__fmul_rn(0.99999f, gsr_sigmoid(__ldg(p + 1))), 1e-6f);
  const float alpha = gsr_sigmoid(__ldg(p + 3));
  GsrMapped m;
  m.sx = __fmul_rn(__fmul_rn(__fmul_rn(sgy, inv_step), 2.0f), inv_w1);
  m.sy = __fmul_rn(__fmul_rn(__fmul_rn(sgx, inv_step), 2.0f), inv_h1);
  m.rho = __fmul_rn(0.999999f, tanhf(__ldg(p + 2)));
  const float mx = __fsub_rn(__fmul_rn(__ldg(p + 7), 2.0f), 1.0f);
  const float my = __fsub_rn(__fmul_rn(__ldg(p + 8), 2.0f), 1.0f);
  m.x = __fsub_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fadd_rn(mx, 1.0f), inv_w), (float)w), inv_w1), 1.0f);
  m.y = __fsub_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fadd_rn(my, 1.0f), inv_h), (float)h), inv_h1), 1.0f);
  m.cr = __fmul_rn(gsr_sigmoid(__ldg(p + 4)), alpha);
  m.cg = __fmul_rn(gsr_sigmoid(__ldg(p + 5)), alpha);
  m.cb = __fmul_rn(gsr_sigmoid(__ldg(p + 6)), alpha);
  return m;
}

// Bucket append.  One thread per Gaussian walks the region bands of its box and collects its entries
// {index | cell mask << 23 | binds << 31} with their region ids in a PRIVATE shared-memory list (no atomics, no
// barrier: a thread only ever touches its own column).  The lists are then appended to the regions' buckets four
// positions at a time: the 32 Gaussians of a warp are consecutive in the input, which for a fea2gs field means
// spatially adjacent (utils/fea2gs.py:553-563), so at the same list position neighbouring lanes name the same
// region; runs of lanes that do form a group, the group's first lane reserves the group's slots with ONE global
// atomicAdd and every lane stores its entry at its rank.  The four atomics of a batch are issued back to back
// before the first result is used, so a warp exposes one atomic round trip per four list positions instead of
// one per position.  All loops are warp-uniform (trip counts are warp maxima, lanes past their own range are
// predicated off): no divergent nesting, and no special path for wide (x8) or incoherent input, which merely
// forms smaller groups.
constexpr int GSR_RB_ECAP = 16;  // list positions per Gaussian between two flushes

// list: [GSR_RB_ECAP][STRIDE] items, this lane's column is `col`.
template <int STRIDE>
__device__ __forceinline__ void gsr_bucket_flush(const uint2* list, int col, int lane, int n,
                                                 int* __restrict__ cnt, uint32_t* __restrict__ ent, int cap,
                                                 int* overflow) {
  const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
  const int nmax = __reduce_max_sync(full, n);
  for (int e0 = 0; e0 < nmax; e0 += 4) {
    unsigned grp[4];
    int base[4];
    uint2 it[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const bool v = e0 + t < n;
      it[t] = list[((e0 + t) & (GSR_RB_ECAP - 1)) * STRIDE + col];
      // runs of consecutive lanes naming the same region form a group (lanes without an entry at this
      // position get an id no region has): heads from one shuffle + ballot, no match instruction
      const int key = v ? (int)it[t].y : -1 - lane;
      const int prev = __shfl_up_sync(full, key, 1);
      const unsigned heads = __ballot_sync(full, lane == 0 || prev != key);
      const unsigned upto = lt | (1u << lane);
      const int lead = 31 - __clz(heads & upto);          // my run's first lane
      const unsigned above = heads & ~upto;
      const int end = above ? __ffs(above) - 1 : 32;      // one past my run's last lane
      grp[t] = v ? (unsigned)((lead << 8) | (lane - lead)) | 0x10000u : 0u;  // {valid, leader, rank}
      base[t] = 0;
      // the run's first lane reserves the run's slots (predicated atomic: no divergent region, and the result
      // is not waited for before the second loop)
      asm volatile("{\n\t.reg .pred pa;\n\tsetp.ne.s32 pa, %3, 0;\n\t@pa atom.global.add.s32 %0, [%1], %2;\n\t}"
                   : "+r"(base[t]) : "l"(cnt + (v ? (int)it[t].y : 0)), "r"(end - lead), "r"((int)(v && lane == lead))
                   : "memory");
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const unsigned g = grp[t];
      const int b = __shfl_sync(full, base[t], (g >> 8) & 31);
      if (g) {
        const int pos = b + (int)(g & 0xffu);
        if (pos < cap) ent[(size_t)it[t].y * cap + pos] = it[t].x;
        else *overflow = 1;
      }
    }
  }
}

// ---- region build ------------------------------------------------------------------------------------------------
// Arranged so that no lane waits for the widest Gaussian of its warp (round 1-2's kernel ran every lane for the
// warp's maximum number of bands and columns: 1960 warp instructions per 32 Gaussians at HL, now ~1750 with a
// third of the atomics).
// A warp takes 32 consecutive Gaussians.
//   phase A  one lane per Gaussian: set-up, record, ellipse -- ellipse and cull box go to the warp's shared memory;
//   phase B  one lane per (Gaussian, region band) ITEM: the bitmaps of the band's two cell rows, the band's first
//            region column and its column count, left in shared memory; the items' (band, column) ENTRIES are
//            listed column-major per batch of 32 items, so that neighbouring list positions hold neighbouring
//            Gaussians at the same relative position -- which mostly name the same region;
//   phase C  one lane per entry: cell mask from the item's bitmaps; runs of lanes naming the same region are
//            found with one shuffle + ballot and the run's first lane reserves the run's bucket slots with ONE
//            atomic (two list positions per lane in flight before the first result is used).
// A warp holding a Gaussian wider than 32 cells or taller than GSR_RB2_MAXB bands walks the generic way
// (gsr_rb_walk_generic: per-lane lists, appended four positions at a time).
// RAW: `sigmas` points to the raw head output (s,9); the mapped parameters are written to msig (s,3), mcrd (s,2),
// mcol (s,3) -- what the backward and the autograd boundary need -- and used from registers: no second pass
// over them.  step: the front end's step size (padded batches: per sample).
#ifndef GSR_CFG_RB2_MIN_CTAS
#define GSR_CFG_RB2_MIN_CTAS 8
#endif
constexpr int GSR_RB2_WARPS = 4;
constexpr int GSR_RB2_THREADS = 32 * GSR_RB2_WARPS;
constexpr int GSR_RB2_MAXB = 12;    // region bands per Gaussian on the balanced path
constexpr int GSR_RB2_PASS = 128;   // items per pass (four batches of 32)
constexpr int GSR_RB2_MAXC = 8;     // region columns per item: 32 cells from the first column's first cell

struct GsrRb2Warp {
  union {
    struct {
      float4 p0[32];                                 // cx, cy, inv_a, kappa
      float4 p1[32];                                 // cp, box.x, box.y (gsr_box_pack, binds in bit 15 of box.x), -
      uint4 item[GSR_RB2_PASS];                      // row bits 0, row bits 1, first region id, shift | lane << 8 | binds << 13
      uint16_t items[32 * GSR_RB2_MAXB];             // lane | band offset << 5, k-major
      uint16_t ents[GSR_RB2_PASS * GSR_RB2_MAXC];    // item slot | column << 7
    } b;
    uint2 list[GSR_RB_ECAP][32];                     // generic walk: {entry, region id} per lane
  };
};

// Generic walk of one lane's box (any width): the old kernel's loop, on a warp-private list.  Warp-collective.
__device__ __noinline__ void gsr_rb_walk_generic(uint2 (*list)[32], int lane, const GsrSetup st, const GsrEllipse e,
                                                 uint32_t entry, float ecut, int nrx, int* __restrict__ cnt,
                                                 uint32_t* __restrict__ ent, int cap, int* overflow);

template <bool RAGGED, bool RAW>
__global__ void __launch_bounds__(GSR_RB2_THREADS, GSR_CFG_RB2_MIN_CTAS)
gsr_region_build2_kernel(const float* __restrict__ sigmas, const float* __restrict__ coords,
                         const float* __restrict__ colors, float* __restrict__ msig, float* __restrict__ mcrd,
                         float* __restrict__ mcol, int s, int h, int w, float dmax, float ksigma, float ecut,
                         float step, GsrWorkspace ws) {
  __shared__ GsrRb2Warp smw[GSR_RB2_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  GsrRb2Warp& sm = smw[warp];
  const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
  int* const overflow = ws.stats + GSR_STAT_OVERFLOW;
  const int nchunks = (s + 31) / 32;

  for (int chunk = blockIdx.x * GSR_RB2_WARPS + warp; chunk < nchunks; chunk += gridDim.x * GSR_RB2_WARPS) {
    const int i = chunk * 32 + lane;
    // ---- phase A: one lane per Gaussian (batches: set up in the sample's own image, then moved to its block of rows of the stack)
    const GsrSampleView sv = gsr_sample_view<RAGGED>(ws, i < s ? i : 0, h, w, dmax);
    const int yoff = sv.yoff, hl = sv.hl, wl = sv.wl;
    GsrSetup st;
    st.live = false;
    st.binds = false;
    st.x0 = st.y0 = 1;
    st.x1 = st.y1 = 0;
    GsrEllipse e;
    e.cx = e.cy = e.inv_a = e.kappa = e.cp = 0.f;
    if (i < s) {
      float sx, sy, rho, x, y, cr, cg, cb;
      if (RAW) {
        float stp = step;
        if (RAGGED) stp = ws.bdesc[i / ws.bn].step;
        const GsrMapped m = gsr_map_one(sigmas + 9 * (size_t)i, ws.hf > 0 ? ws.hf : hl, wl, stp);
        sx = m.sx, sy = m.sy, rho = m.rho, x = m.x, y = m.y, cr = m.cr, cg = m.cg, cb = m.cb;
        float* ms = msig + 3 * (size_t)i;
        float* mc = mcrd + 2 * (size_t)i;
        float* mk = mcol + 3 * (size_t)i;
        ms[0] = sx, ms[1] = sy, ms[2] = rho;
        mc[0] = x, mc[1] = y;
        mk[0] = cr, mk[1] = cg, mk[2] = cb;
      } else {
        sy = __ldg(sigmas + 3 * (size_t)i + 1);
        y = __ldg(coords + 2 * (size_t)i + 1);
        sx = rho = x = cr = cg = cb = 0.f;
        // row-band view: a Gaussian whose k-sigma rows miss the band by more than a pixel is dropped after two loads
        bool reach = true;
        if (ws.hf > 0) {
          const float hyf = 0.5f * (float)(ws.hf - 1);
          const float cyf = (y + 1.0f) * hyf, eyf = ksigma * fabsf(sy) * hyf + 1.0f;
          reach = !(cyf + eyf < (float)ws.row0 - 1.0f || cyf - eyf > (float)(ws.row0 + h));
        }
        if (reach) {
          sx = __ldg(sigmas + 3 * (size_t)i + 0);
          rho = __ldg(sigmas + 3 * (size_t)i + 2);
          x = __ldg(coords + 2 * (size_t)i + 0);
          cr = __ldg(colors + 3 * (size_t)i + 0);
          cg = __ldg(colors + 3 * (size_t)i + 1);
          cb = __ldg(colors + 3 * (size_t)i + 2);
        } else {
          sy = 0.f;  // sigma = 0: gsr_setup returns "not live" at its first test
        }
      }
      st = gsr_setup(sx, sy, rho, x, y, cr, cg, cb, hl, wl, sv.dmax, ksigma, sv.px_tab, sv.py_tab, ws.hf, ws.row0);
      if (st.live) {
        const GsrRec r = gsr_make_rec(sx, sy, rho, x, y, cr, cg, cb);
        if (!(gsr_finite(r.a) & gsr_finite(r.b) & gsr_finite(r.c))) st.live = false;
        st.y0 += yoff;
        st.y1 += yoff;
        if (gsr_edge_binds<RAGGED>(sv, st)) st.binds = true;
        if (st.live) {
          GsrRec rs = r;
          if (RAGGED) gsr_rescale_rec(rs, sv);
          float4* dr = reinterpret_cast<float4*>(ws.rec_in + i);
          dr[0] = make_float4(rs.x, rs.y, rs.a, rs.b);
          dr[1] = make_float4(rs.c, rs.r, rs.g, rs.bl);
          if (st.binds) ws.box_in[i] = gsr_box_pack(st.x0, st.x1, st.y0, st.y1, true);
          e = gsr_ellipse(r, hl, wl, ws.hf, ws.row0);
          e.cy += (float)yoff;
        }
      }
    }
    if (!st.live) {
      st.x0 = st.y0 = 1;
      st.x1 = st.y1 = 0;
    }
    const uint32_t entry = (uint32_t)i | (st.binds ? 0x80000000u : 0u);
    // (box corners of a live Gaussian are non-negative: unsigned divisions are shifts)
    const int nb = st.live ? (int)((unsigned)st.y1 / GSR_RGH) - (int)((unsigned)st.y0 / GSR_RGH) + 1 : 0;
    const int cbase0 = (int)((unsigned)st.x0 / GSR_RGW) * GSR_CELLS_X;
    const bool fits = !st.live || ((int)((unsigned)st.x1 / GSR_CELL) - cbase0 < 32 && nb <= GSR_RB2_MAXB);
    if (!__all_sync(full, fits)) {
      gsr_rb_walk_generic(sm.list, lane, st, e, entry, ecut, ws.nrx, ws.reg_count, ws.entries, ws.reg_cap, overflow);
      __syncwarp();
      continue;
    }
    const uint2 pk = gsr_box_pack(st.x0, st.x1, st.y0, st.y1, st.binds);
    sm.b.p0[lane] = make_float4(e.cx, e.cy, e.inv_a, e.kappa);
    sm.b.p1[lane] = make_float4(e.cp, __uint_as_float(pk.x), __uint_as_float(pk.y), 0.f);
    // item list, band-major: (every lane whose box reaches band b), b = first band of the warp, ...  Lanes next to
    // each other in the list are then neighbouring Gaussians in the SAME band (whatever their jitter across a band
    // boundary).  Input whose bands lie far apart (incoherent order, a chunk across two samples of a batch) lists
    // by band offset instead.
    const int b0 = (int)((unsigned)st.y0 / GSR_RGH);
    const int bmin = __reduce_min_sync(full, nb > 0 ? b0 : 0x7fffffff);
    const int bmax = __reduce_max_sync(full, nb > 0 ? b0 + nb - 1 : -1);
    const bool by_band = bmax - bmin < 2 * GSR_RB2_MAXB;
    const int KB = by_band ? bmax - bmin + 1 : __reduce_max_sync(full, nb);
    const int kfirst = by_band ? bmin - b0 : 0;  // this lane's band offset at list round 0
    int nitems = 0;
    for (int kk = 0; kk < KB; ++kk) {
      const int k = kfirst + kk;
      const bool has = k >= 0 && k < nb;
      const unsigned have = __ballot_sync(full, has);
      if (has) sm.b.items[nitems + __popc(have & lt)] = (uint16_t)(lane | (k << 5));
      nitems += __popc(have);
    }
    __syncwarp();
    const uint32_t ibase = (uint32_t)(chunk * 32);

    for (int pass0 = 0; pass0 < nitems; pass0 += GSR_RB2_PASS) {
      const int pend = min(nitems, pass0 + GSR_RB2_PASS);
      // ---- phase B: one lane per (Gaussian, band)
      int nent = 0;
      for (int t0 = pass0; t0 < pend; t0 += 32) {
        const bool act = t0 + lane < pend;
        const uint32_t item = act ? sm.b.items[t0 + lane] : 0u;
        const int g = (int)(item & 31u), k = (int)(item >> 5);
        const float4 q0 = sm.b.p0[g], q1 = sm.b.p1[g];
        GsrEllipse eg;
        eg.cx = q0.x;
        eg.cy = q0.y;
        eg.inv_a = q0.z;
        eg.kappa = q0.w;
        eg.cp = q1.x;
        int x0, x1, y0, y1;
        bool binds;
        gsr_box_unpack(make_uint2(__float_as_uint(q1.y), __float_as_uint(q1.z)), x0, x1, y0, y1, binds);
        const int b = (int)((unsigned)y0 / GSR_RGH) + k;
        const int cbase = (int)((unsigned)x0 / GSR_RGW) * GSR_CELLS_X;
        uint32_t rb[2] = {0u, 0u};
        int ca = 0, ncol = 0;
        if (act && gsr_band_rowbits(eg, ecut, b, x0, x1, y0, y1, cbase, rb)) {
          const uint32_t any = rb[0] | rb[1];
          ca = (int)((unsigned)(cbase + __ffs(any) - 1) / GSR_CELLS_X);
          ncol = (int)((unsigned)(cbase + 31 - __clz(any)) / GSR_CELLS_X) - ca + 1;
        }
        const int slot = t0 - pass0 + lane;
        sm.b.item[slot] = make_uint4(rb[0], rb[1], (uint32_t)(b * ws.nrx + ca),
                                     (uint32_t)(ca * GSR_CELLS_X - cbase) | ((uint32_t)g << 8) | (binds ? 0x2000u : 0u));
        // entries of this batch, column-major (by absolute region column when the batch's columns lie close)
        const int cmin = __reduce_min_sync(full, ncol > 0 ? ca : 0x7fffffff);
        const int cmax = __reduce_max_sync(full, ncol > 0 ? ca + ncol - 1 : -1);
        const bool by_col = cmax - cmin < 2 * GSR_RB2_MAXC;
        const int KC = by_col ? cmax - cmin + 1 : __reduce_max_sync(full, ncol);
        const int jfirst = by_col ? cmin - ca : 0;
        for (int jj = 0; jj < KC; ++jj) {
          const int j = jfirst + jj;
          const bool has = j >= 0 && j < ncol;
          const unsigned have = __ballot_sync(full, has);
          if (has) sm.b.ents[nent + __popc(have & lt)] = (uint16_t)(slot | (j << 7));
          nent += __popc(have);
        }
      }
      __syncwarp();
      // ---- phase C: one lane per entry, two list positions per lane and round
      for (int t0 = 0; t0 < nent; t0 += 64) {
        uint32_t m[2], word[2];
        int rid[2], base[2], lead[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int t = t0 + 32 * u + lane;
          const uint32_t en = t < nent ? sm.b.ents[t] : 0u;
          const uint4 it = sm.b.item[en & 127u];
          const int j = (int)(en >> 7);
          const unsigned sh = (it.w & 31u) + 4u * (unsigned)j;
          m[u] = t < nent ? ((it.x >> sh) & 0xfu) | (((it.y >> sh) & 0xfu) << GSR_CELLS_X) : 0u;
          rid[u] = (int)it.z + j;
          word[u] = (ibase + ((it.w >> 8) & 31u)) | ((it.w & 0x2000u) << 18) | (m[u] << GSR_ENT_MASK_SHIFT);
          // runs of consecutive lanes naming the same region form a group (lanes without an entry get an id no
          // region has): heads from one shuffle + ballot
          const int key = m[u] ? rid[u] : -1 - lane;
          const int prev = __shfl_up_sync(full, key, 1);
          const unsigned heads = __ballot_sync(full, lane == 0 || prev != key);
          const unsigned upto = lt | (1u << lane);
          lead[u] = 31 - __clz(heads & upto);                 // my run's first lane
          const unsigned above = heads & ~upto;
          const int end = above ? __ffs(above) - 1 : 32;      // one past my run's last lane
          base[u] = 0;
          // the run's first lane reserves the run's slots (predicated atomic: no divergent region; the result is
          // not waited for before the second loop)
          asm volatile("{\n\t.reg .pred pa;\n\tsetp.ne.s32 pa, %3, 0;\n\t@pa atom.global.add.s32 %0, [%1], %2;\n\t}"
                       : "+r"(base[u]) : "l"(ws.reg_count + (m[u] ? rid[u] : 0)), "r"(end - lead[u]),
                         "r"((int)(m[u] != 0u && lane == lead[u])) : "memory");
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int bs = __shfl_sync(full, base[u], lead[u]);
          if (m[u]) {
            const int pos = bs + lane - lead[u];
            if (pos < ws.reg_cap) ws.entries[(size_t)rid[u] * ws.reg_cap + pos] = word[u];
            else *overflow = 1;
          }
        }
      }
      __syncwarp();  // items and entries are rewritten by the next pass / chunk
    }
  }
}

__device__ __noinline__ void gsr_rb_walk_generic(uint2 (*list)[32], int lane, const GsrSetup st, const GsrEllipse e,
                                                 uint32_t entry, float ecut, int nrx, int* __restrict__ cnt,
                                                 uint32_t* __restrict__ ent, int cap, int* overflow) {
  const unsigned full = 0xffffffffu;
  const int b0 = (int)((unsigned)st.y0 / GSR_RGH), nb = st.live ? (int)((unsigned)st.y1 / GSR_RGH) - b0 + 1 : 0;
  const int KB = __reduce_max_sync(full, nb);
  int n = 0;
  for (int k = 0; k < KB; ++k) {
    const int b = b0 + k;
    int cl[2] = {1, 1}, ch[2] = {0, 0}, ca = 0, ncol = 0;
    if (k < nb && gsr_band_cells(e, ecut, b, st.x0, st.x1, st.y0, st.y1, cl, ch)) {
      int cb_;
      gsr_band_columns(cl, ch, ca, cb_);
      ncol = cb_ - ca + 1;
    }
    const int KC = __reduce_max_sync(full, ncol);
    for (int j0 = 0; j0 < KC; j0 += 4) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = j0 + jj, c = ca + j;
        const uint32_t m = j < ncol ? gsr_cell_mask(cl, ch, c) : 0u;
        if (m != 0u) {
          list[n][lane] = make_uint2(entry | (m << GSR_ENT_MASK_SHIFT), (uint32_t)(b * nrx + c));
          ++n;
        }
      }
      if (__any_sync(full, n > GSR_RB_ECAP - 4)) {  // the next four might not fit
        gsr_bucket_flush<32>(&list[0][0], lane, lane, n, cnt, ent, cap, overflow);
        n = 0;
      }
    }
  }
  gsr_bucket_flush<32>(&list[0][0], lane, lane, n, cnt, ent, cap, overflow);
}

// ---- deterministic mode (GSR_FLAG_DETERMINISTIC) -------------------------------------------------------------
// Bucket ranks come from atomics, so the ORDER of a bucket's entries -- and with it the order in which a pixel's
// terms are added in fp32 -- differs from run to run (as do the reference's atomicAdds).  Sorting every bucket by
// Gaussian index makes buckets, chunks and cell lists functions of the input alone: the forward becomes
// bit-reproducible (the backward already is: one warp owns a Gaussian, fixed sweep order, no atomics).  Indices
// are unique within a bucket.  One warp per bucket: a stable LSD radix sort in shared memory on the index minus the
// bucket's smallest one (two 8-bit passes cover a fea2gs bucket); buckets longer than GSR_SORT_MAX entries are ranked
// by a whole CTA through the bucket itself (an entry's rank is the number of smaller keys: rank pass, barrier, write
// pass).
constexpr int GSR_SORT_WARPS = 4;
constexpr int GSR_SORT_MAX = 1024;  // entries per warp-sorted bucket (4 KB of shared memory per warp)

__global__ void __launch_bounds__(32 * GSR_SORT_WARPS)
gsr_bucket_sort_kernel(uint32_t* __restrict__ entries, const int* __restrict__ reg_count, int cap, int nreg,
                       const int* guard, int want) {
  if (gsr_guard_skip(guard, want)) return;
  __shared__ __align__(16) uint32_t buf[GSR_SORT_WARPS][2][GSR_SORT_MAX + 4];  // entries, ping-pong
  __shared__ int hist[GSR_SORT_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned full = 0xffffffffu, lt = (1u << lane) - 1u;
  int* hs = hist[warp];
  for (int r = blockIdx.x * GSR_SORT_WARPS + warp; r < nreg; r += gridDim.x * GSR_SORT_WARPS) {
    const int n = min(__ldg(reg_count + r), cap);
    if (n < 2 || n > GSR_SORT_MAX) continue;  // (longer buckets: gsr_bucket_sort_long_kernel)
    uint32_t* e = entries + (size_t)r * cap;
    uint32_t* src = buf[warp][0];
    uint32_t* dst = buf[warp][1];
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int i = lane; i < n; i += 32) {
      const uint32_t v = e[i], k = v & GSR_ENT_INDEX;
      src[i] = v;
      kmin = min(kmin, k);
      kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(full, kmin);
    kmax = __reduce_max_sync(full, kmax);
    // Stable LSD radix sort on (index - smallest index of the bucket), 8 bits per pass: the Gaussians of a region come
    // from a few rows of the field, so two passes cover the range of a fea2gs bucket (three always suffice: 23 bits).
    const uint32_t range = kmax - kmin;
    const int passes = range < 256u ? 1 : range < 65536u ? 2 : 3;
    for (int ps = 0; ps < passes; ++ps) {
      const int shift = 8 * ps;
#pragma unroll
      for (int k = 0; k < 8; ++k) hs[lane * 8 + k] = 0;
      __syncwarp();
      for (int i = lane; i < n; i += 32) atomicAdd(hs + ((((src[i] & GSR_ENT_INDEX) - kmin) >> shift) & 255u), 1);
      __syncwarp();
      {  // exclusive scan of the 256 counters: eight per lane
        int c[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          c[k] = hs[lane * 8 + k];
          sum += c[k];
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(full, incl, d);
          if (lane >= d) incl += t;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          hs[lane * 8 + k] = run;
          run += c[k];
        }
      }
      __syncwarp();
      for (int i0 = 0; i0 < n; i0 += 32) {  // in order: entries with equal digits keep their order (stable)
        const int i = i0 + lane;
        const bool valid = i < n;
        const uint32_t v = valid ? src[i] : 0u;
        const int d = valid ? (int)((((v & GSR_ENT_INDEX) - kmin) >> shift) & 255u) : 256 + lane;
        const unsigned peers = __match_any_sync(full, d);
        const int base = valid ? hs[d] : 0;
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) hs[d] = base + __popc(peers);
        __syncwarp();
        if (valid) dst[base + __popc(peers & lt)] = v;
      }
      __syncwarp();
      uint32_t* t = src;
      src = dst;
      dst = t;
    }
    for (int i = lane; i < n; i += 32) e[i] = src[i];
    __syncwarp();
  }
}

// Buckets longer than GSR_SORT_MAX (dense or clustered fields): one CTA per bucket, keys read from the bucket
// itself (L1/L2), ranks kept in registers until every thread has finished reading.
constexpr int GSR_SORT_LONG_PER_THREAD = 32;  // 256 threads x 32 = 8192 entries; longer buckets stay unsorted
__global__ void __launch_bounds__(256)
gsr_bucket_sort_long_kernel(uint32_t* __restrict__ entries, const int* __restrict__ reg_count, int cap, int nreg,
                            const int* guard, int want, int* unsorted) {
  if (gsr_guard_skip(guard, want)) return;
  for (int r = blockIdx.x; r < nreg; r += gridDim.x) {
    const int n = min(__ldg(reg_count + r), cap);
    if (n <= GSR_SORT_MAX) continue;
    if (n > 256 * GSR_SORT_LONG_PER_THREAD) {
      if (threadIdx.x == 0) *unsorted = 1;
      continue;
    }
    uint32_t* e = entries + (size_t)r * cap;
    uint32_t v[GSR_SORT_LONG_PER_THREAD];
    int rank[GSR_SORT_LONG_PER_THREAD];
#pragma unroll 1
    for (int t = 0; t < GSR_SORT_LONG_PER_THREAD; ++t) {
      const int i = threadIdx.x + 256 * t;
      v[t] = 0u;
      rank[t] = 0;
      if (i < n) {
        v[t] = e[i];
        const uint32_t key = v[t] & GSR_ENT_INDEX;
        int rk = 0;
        for (int j = 0; j < n; ++j) rk += ((*(const volatile uint32_t*)(e + j)) & GSR_ENT_INDEX) < key;
        rank[t] = rk;
      }
    }
    __syncthreads();  // every read of the bucket is done
#pragma unroll 1
    for (int t = 0; t < GSR_SORT_LONG_PER_THREAD; ++t)
      if (threadIdx.x + 256 * t < n) e[rank[t]] = v[t];
    __syncthreads();
  }
}

