// gsr_prepass.cuh -- O(N) set-up pipelines shared by forward and backward.
//
// Region-bucket pipeline (forward fast path), ONE pass over the Gaussians:
//   T1 gsr_region_build_kernel  per Gaussian: exact dmax window /\ k-sigma box -> cull box, raster
//                               record (written in input order); then for every 8x8-pixel region
//                               the ellipse touches a 4-byte entry {index | binds << 31} is appended
//                               to the region's bucket (small persistent CTAs collect their entries
//                               per region in shared memory and reserve bucket slots with one global
//                               atomic per (CTA, region); a warp-ballot path serves incoherent input).
//   Every region owns a fixed-capacity bucket (40 N / regions + 32 entries: GSASR emits its Gaussians
//   on a regular grid, utils/fea2gs.py:553-563, so the load per region is uniform); an entry that
//   does not fit raises the overflow flag and the forward falls back to the home-bin pipeline.
//   The forward kernel then needs no culling at all: half a warp per region streams its bucket.
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
// not taken return at once (`guard`).
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
// bucket capacity per region = 40 * N / regions + 32: a Gaussian of the x4 head touches 6.8 regions on
// average, one of the x8 head 23 (5 sigma = 33 px); 4 bytes per slot
constexpr int GSR_ENTRIES_PER_GAUSSIAN = 40;

struct GsrWorkspace {
  // ---- one block, cleared per call ----
  int* bin_count;    // nb + 1
  int* stats;        // 8 ints, see GSR_STAT_*
  int* scan_state;   // 2 * nscan     look-back state of the bin scan
  int* reg_count;    // nreg      entries appended to each region's bucket (may exceed reg_cap)
  size_t zero_bytes;
  // ---- tile-list pipeline ----
  GsrRec* rec_in;    // s        records, input order
  uint2* box_in;     // s        packed cull boxes, input order (only read for window-binding ones)
  uint32_t* entries; // nreg * reg_cap   Gaussian index | binds << 31
  int ntx, nty, nrx, nry, nreg, reg_cap;
  // ---- home-bin pipeline ----
  int* bin_off;      // nb + 2   exclusive offsets; [nb] = start of large, [nb+1] = n_live
  uint2* box_tmp;    // s   (unsorted)
  int2* keyrank;     // s   (unsorted)  key = bin id, -1 = skipped
  GsrRec* rec;       // s   (sorted)
  uint2* box;        // s   (sorted)
  int* ids;          // s   (sorted -> original index)
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
  ws.ntx = (w + GSR_TILE_W - 1) / GSR_TILE_W;
  ws.nty = (h + GSR_TILE_H - 1) / GSR_TILE_H;
  ws.nrx = ws.ntx * GSR_NRX;  // region grid padded to whole tiles
  ws.nry = ws.nty * GSR_NRY;
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
  ws.zero_bytes = ((size_t)ws.nb + 1 + 8 + 2 * (size_t)ws.nscan + (size_t)ws.nreg) * sizeof(int);
  ws.bin_count = (int*)take(ws.zero_bytes);
  ws.stats = ws.bin_count ? ws.bin_count + ws.nb + 1 : nullptr;
  ws.scan_state = ws.bin_count ? ws.stats + 8 : nullptr;
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
  return RAGGED && (((v.wl % GSR_REGION) != 0 && st.x1 / GSR_REGION == (v.wl - 1) / GSR_REGION) ||
                       ((v.hl % GSR_REGION) != 0 && (st.y1 - v.yoff) / GSR_REGION == (v.hl - 1) / GSR_REGION));
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
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s; i += gridDim.x * blockDim.x)
    gsr_bin_one<RAGGED>(sigmas, coords, colors, i, h, w, dmax, ksigma, ws);
}

// Pixel coordinate tables: the reference's rule (gs.cu:39,46), evaluated once per axis entry.
__global__ void __launch_bounds__(256) gsr_table_kernel(float* __restrict__ px_tab,
                                                        float* __restrict__ py_tab, int h, int w,
                                                        int hf, int row0, int bhs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < w) px_tab[i] = gsr_pix_coord(i, w);
  if (i < h) py_tab[i] = bhs > 0 ? gsr_pix_coord(i % bhs, bhs) : hf > 0 ? gsr_pix_coord(i + row0, hf) : gsr_pix_coord(i, h);
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

// ---- tile-list pipeline -------------------------------------------------------------------------
// Warp-cooperative bucket append.  The 32 Gaussians of a warp are consecutive in the input, which
// for a fea2gs field means spatially adjacent (utils/fea2gs.py:553-563), so their region sets
// overlap heavily.  Regions are addressed as (band, column): band = 8-row strip of the image.
//   1. per band of the warp's union every lane computes the column range its ellipse covers
//      (gsr_band_xrange); per (band, column) of the union one ballot tells which lanes touch the
//      region, and non-empty pairs are dealt to the lanes, 32 at a time;
//   2. lane = pair: ONE atomicAdd reserves the pair's slots (32 reservations in flight per
//      instruction, so the atomic round trip is paid once per batch, not once per pair), then the
//      lane streams the entries of its pair -- fetched from the owning lanes with shuffles -- into
//      the region's bucket.
// If the union is large (incoherent input order) every lane walks its own regions with one atomic
// per entry.  All 32 lanes must call this function (dead lanes pass live = false).
constexpr int GSR_COOP_MAX_PAIRS = 256;

__device__ __forceinline__ void gsr_warp_flush_pairs(unsigned bal, int rid, uint32_t entry,
                                                     int* __restrict__ cnt, uint32_t* __restrict__ ent,
                                                     int cap, int* overflow) {
  const unsigned full = 0xffffffffu;
  uint32_t* dst = ent;
  int room = 0;
  if (bal) {
    const int base = atomicAdd(cnt + rid, __popc(bal));
    dst = ent + (size_t)rid * cap + base;
    room = cap - base;  // entries that still fit
    if (room < __popc(bal)) *overflow = 1;
  }
  const int iters = __reduce_max_sync(full, __popc(bal));
  for (int k = 0; k < iters; ++k) {
    const int src = bal ? __ffs(bal) - 1 : 0;
    const uint32_t en = __shfl_sync(full, entry, src);
    if (bal && k < room) dst[k] = en;
    bal &= bal - 1;
  }
}

__device__ __forceinline__ void gsr_warp_append(bool live, const GsrRec& r, uint32_t entry, int x0,
                                                int x1, int y0, int y1, int h, int w, int nrx,
                                                float ecut, int* __restrict__ cnt,
                                                uint32_t* __restrict__ ent, int cap, int* overflow,
                                                int hf = 0, int row0 = 0, int yoff = 0) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int b0 = live ? y0 / GSR_REGION : 0x3fffffff, b1 = live ? y1 / GSR_REGION : -1;
  const int c0 = live ? x0 / GSR_REGION : 0x3fffffff, c1 = live ? x1 / GSR_REGION : -1;
  const int ub0 = __reduce_min_sync(full, b0), ub1 = __reduce_max_sync(full, b1);
  const int uc0 = __reduce_min_sync(full, c0), uc1 = __reduce_max_sync(full, c1);
  if (ub1 < ub0 || uc1 < uc0) return;  // no live lane
  GsrEllipse e;
  if (live) {
    e = gsr_ellipse(r, h, w, hf, row0);
    e.cy += (float)yoff;  // uniform batch: the sample's block of rows
  }
  if ((long long)(ub1 - ub0 + 1) * (uc1 - uc0 + 1) > GSR_COOP_MAX_PAIRS) {
    if (live)
      for (int b = b0; b <= b1; ++b) {
        int ya = b * GSR_REGION, yb = ya + GSR_REGION - 1, xl, xh;
        ya = ya > y0 ? ya : y0;
        yb = yb < y1 ? yb : y1;
        if (!gsr_band_xrange(e, ecut, ya, yb, x0, x1, xl, xh)) continue;
        for (int c = xl / GSR_REGION; c <= xh / GSR_REGION; ++c) {
          const int rid = b * nrx + c;
          const int pos = atomicAdd(cnt + rid, 1);
          if (pos < cap) ent[(size_t)rid * cap + pos] = entry;
          else *overflow = 1;
        }
      }
    return;
  }
  int slot = 0;  // pairs dealt in the current batch (warp-uniform)
  unsigned mybal = 0;
  int myrid = 0;
  for (int b = ub0; b <= ub1; ++b) {
    int rc0 = 1, rc1 = 0;  // this lane's column range in band b (empty by default)
    if (live && b >= b0 && b <= b1) {
      int ya = b * GSR_REGION, yb = ya + GSR_REGION - 1, xl, xh;
      ya = ya > y0 ? ya : y0;
      yb = yb < y1 ? yb : y1;
      if (gsr_band_xrange(e, ecut, ya, yb, x0, x1, xl, xh)) {
        rc0 = xl / GSR_REGION;
        rc1 = xh / GSR_REGION;
      }
    }
    // columns that any lane touches in this band
    const int bc0 = __reduce_min_sync(full, rc0 <= rc1 ? rc0 : 0x3fffffff);
    const int bc1 = __reduce_max_sync(full, rc0 <= rc1 ? rc1 : -1);
    for (int c = bc0; c <= bc1; ++c) {
      const unsigned bal = __ballot_sync(full, rc0 <= c && c <= rc1);
      if (bal) {
        if (lane == slot) {
          mybal = bal;
          myrid = b * nrx + c;
        }
        if (++slot == 32) {
          gsr_warp_flush_pairs(mybal, myrid, entry, cnt, ent, cap, overflow);
          slot = 0;
          mybal = 0;
        }
      }
    }
  }
  if (slot) gsr_warp_flush_pairs(mybal, myrid, entry, cnt, ent, cap, overflow);
}

// CTA-cooperative bucket append (the common case).  The Gaussians of a CTA are consecutive in the input
// -- for a fea2gs field a piece of one grid row -- so the regions they touch form a small rectangle
// (18 x 4 regions for 64 Gaussians at x4).  When it has at most GSR_RB_MAXR regions the CTA collects its
// entries per region in shared memory (one shared-memory atomic per entry for the slot), reserves each
// non-empty region's slots in the global bucket with ONE atomicAdd (a thread per region) and copies the
// short lists out.  A list longer than GSR_RB_LCAP spills its tail straight to the bucket (one global
// atomic per entry); a CTA whose rectangle is larger (incoherent input order) takes the ballot-based
// path above.  CTAs are small (two warps: the barriers cost little and many CTAs are in different
// phases at any time) and persistent: each strides over the chunks of the input and requests the
// parameters of its next chunk before it processes the current one.
#ifndef GSR_CFG_RB_THREADS
#define GSR_CFG_RB_THREADS 64
#endif
constexpr int GSR_RB_THREADS = GSR_CFG_RB_THREADS;
constexpr int GSR_RB_MAXR = 2 * GSR_RB_THREADS;
constexpr int GSR_RB_LCAP = 16;

struct GsrRegionBuildSmem {
  int cnt[GSR_RB_MAXR];
  uint32_t list[GSR_RB_MAXR][GSR_RB_LCAP + 1];  // odd row stride: a thread per region reads conflict-free
  int red[4][(GSR_RB_THREADS + 31) / 32];
};

#ifndef GSR_CFG_RB_MIN_CTAS
#define GSR_CFG_RB_MIN_CTAS (1024 / GSR_CFG_RB_THREADS)
#endif
template <bool RAGGED>
__global__ void __launch_bounds__(GSR_RB_THREADS, GSR_CFG_RB_MIN_CTAS)
gsr_region_build_kernel(const float* __restrict__ sigmas, const float* __restrict__ coords,
                      const float* __restrict__ colors, int s, int h, int w, float dmax,
                      float ksigma, float ecut, GsrWorkspace ws) {
  __shared__ GsrRegionBuildSmem sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned full = 0xffffffffu;
  int* const overflow = ws.stats + GSR_STAT_OVERFLOW;
  const int nchunks = (s + GSR_RB_THREADS - 1) / GSR_RB_THREADS;

  float pf[8];
  auto request = [&](int chunk) {
    const int i = chunk * GSR_RB_THREADS + tid;
#pragma unroll
    for (int k = 0; k < 8; ++k) pf[k] = 0.f;
    if (chunk < nchunks && i < s) {
      pf[0] = __ldg(sigmas + 3 * (size_t)i + 0);
      pf[1] = __ldg(sigmas + 3 * (size_t)i + 1);
      pf[2] = __ldg(sigmas + 3 * (size_t)i + 2);
      pf[3] = __ldg(coords + 2 * (size_t)i + 0);
      pf[4] = __ldg(coords + 2 * (size_t)i + 1);
      pf[5] = __ldg(colors + 3 * (size_t)i + 0);
      pf[6] = __ldg(colors + 3 * (size_t)i + 1);
      pf[7] = __ldg(colors + 3 * (size_t)i + 2);
    }
  };
  request(blockIdx.x);

  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int i = chunk * GSR_RB_THREADS + tid;
    const float sx = pf[0], sy = pf[1], rho = pf[2], x = pf[3], y = pf[4], cr = pf[5], cg = pf[6], cb = pf[7];
    request(chunk + gridDim.x);
    // batches: set up in the sample's own image (hl x wl, its dmax), then move to its block of rows of the
    // stack; a padded batch stores the record rescaled to the canvas' coordinate normalisation, the raw
    // one (r) keeps describing the ellipse in the sample's own pixels for the region masks
    const GsrSampleView sv = gsr_sample_view<RAGGED>(ws, i, h, w, dmax);
    const int yoff = sv.yoff, hl = sv.hl, wl = sv.wl;
    GsrSetup st;
    st.live = false;
    st.binds = false;
    st.x0 = st.y0 = 1;
    st.x1 = st.y1 = 0;
    GsrRec r;
    r.x = r.y = r.a = r.b = r.c = r.r = r.g = r.bl = 0.f;
    if (i < s) {
      st = gsr_setup(sx, sy, rho, x, y, cr, cg, cb, hl, wl, sv.dmax, ksigma, sv.px_tab, sv.py_tab, ws.hf, ws.row0);
      if (st.live) {
        r = gsr_make_rec(sx, sy, rho, x, y, cr, cg, cb);
        if (!(gsr_finite(r.a) && gsr_finite(r.b) && gsr_finite(r.c))) st.live = false;
        st.y0 += yoff;
        st.y1 += yoff;
        if (gsr_edge_binds<RAGGED>(sv, st)) st.binds = true;
      }
      if (st.live) {
        GsrRec rs = r;
        if (RAGGED) gsr_rescale_rec(rs, sv);
        float4* dr = reinterpret_cast<float4*>(ws.rec_in + i);
        dr[0] = make_float4(rs.x, rs.y, rs.a, rs.b);
        dr[1] = make_float4(rs.c, rs.r, rs.g, rs.bl);
        if (st.binds) ws.box_in[i] = gsr_box_pack(st.x0, st.x1, st.y0, st.y1, true);
      }
    }
    const uint32_t entry = (uint32_t)i | (st.binds ? 0x80000000u : 0u);

    // ---- the CTA's region rectangle
    const int b0 = st.live ? st.y0 / GSR_REGION : 0x3fffffff, b1 = st.live ? st.y1 / GSR_REGION : -1;
    const int c0 = st.live ? st.x0 / GSR_REGION : 0x3fffffff, c1 = st.live ? st.x1 / GSR_REGION : -1;
    {
      const int wb0 = __reduce_min_sync(full, b0), wb1 = __reduce_max_sync(full, b1);
      const int wc0 = __reduce_min_sync(full, c0), wc1 = __reduce_max_sync(full, c1);
      if (lane == 0) {
        sm.red[0][warp] = wb0;
        sm.red[1][warp] = wb1;
        sm.red[2][warp] = wc0;
        sm.red[3][warp] = wc1;
      }
    }
    for (int k = tid; k < GSR_RB_MAXR; k += GSR_RB_THREADS) sm.cnt[k] = 0;
    __syncthreads();
    int B0 = 0x3fffffff, B1 = -1, C0 = 0x3fffffff, C1 = -1;
#pragma unroll
    for (int k = 0; k < (GSR_RB_THREADS + 31) / 32; ++k) {
      B0 = min(B0, sm.red[0][k]);
      B1 = max(B1, sm.red[1][k]);
      C0 = min(C0, sm.red[2][k]);
      C1 = max(C1, sm.red[3][k]);
    }
    const bool any_live = B1 >= B0 && C1 >= C0;  // CTA-uniform
    const int NC = C1 - C0 + 1, NR = any_live ? NC * (B1 - B0 + 1) : 0;
    if (any_live && (long long)NC * (B1 - B0 + 1) > GSR_RB_MAXR) {
      gsr_warp_append(st.live, r, entry, st.x0, st.x1, st.y0, st.y1, hl, wl, ws.nrx, ecut, ws.reg_count, ws.entries,
                      ws.reg_cap, overflow, ws.hf, ws.row0, yoff);
    } else if (any_live) {
      // ---- collect: one shared-memory atomic per (Gaussian, region)
      if (st.live) {
        GsrEllipse e = gsr_ellipse(r, hl, wl, ws.hf, ws.row0);
        e.cy += (float)yoff;
        for (int b = b0; b <= b1; ++b) {
          int ya = b * GSR_REGION, yb = ya + GSR_REGION - 1, xl, xh;
          ya = ya > st.y0 ? ya : st.y0;
          yb = yb < st.y1 ? yb : st.y1;
          if (!gsr_band_xrange(e, ecut, ya, yb, st.x0, st.x1, xl, xh)) continue;
          const int row = (b - B0) * NC - C0;
          for (int c = xl / GSR_REGION; c <= xh / GSR_REGION; ++c) {
            const int slot = atomicAdd(&sm.cnt[row + c], 1);
            if (slot < GSR_RB_LCAP) {
              sm.list[row + c][slot] = entry;
            } else {  // list full: straight to the bucket
              const int rid = b * ws.nrx + c;
              const int pos = atomicAdd(ws.reg_count + rid, 1);
              if (pos < ws.reg_cap) ws.entries[(size_t)rid * ws.reg_cap + pos] = entry;
              else *overflow = 1;
            }
          }
        }
      }
      __syncthreads();
      // ---- reserve and copy out: a thread per region of the rectangle -- ONE global atomic for the
      // region's slots, then its short list goes to the bucket
      for (int k = tid; k < NR; k += GSR_RB_THREADS) {
        const int n = min(sm.cnt[k], GSR_RB_LCAP);
        if (n > 0) {
          const int rb = k / NC;
          const int rid = (B0 + rb) * ws.nrx + C0 + (k - rb * NC);
          const int pos = atomicAdd(ws.reg_count + rid, n);
          uint32_t* dst = ws.entries + (size_t)rid * ws.reg_cap + pos;
          const int room = ws.reg_cap - pos;
          if (room < n) *overflow = 1;
          for (int j = 0; j < n && j < room; ++j) dst[j] = sm.list[k][j];
        }
      }
    }
    __syncthreads();  // shared memory is reused by the next chunk
  }
}
