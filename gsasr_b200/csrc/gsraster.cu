// gsraster.cu -- the C ABI of libgsraster.so (see include/gsraster.h for the contract and the
// reference interfaces each entry point replaces).  One translation unit: set-up pipeline,
// forward and backward raster kernels, fused front end, and CPU test hooks.
#include "../../include/gsraster.h"
#include "gsr_backward.cuh"
#include "gsr_backward_region.cuh"
#include "gsr_frontend.cuh"
#include "gsr_loss.cuh"
#include <atomic>
#include <cstdlib>

static thread_local int g_last_cuda_error = 0;

#define GSR_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) {                           \
      g_last_cuda_error = (int)e__;                     \
      (void)cudaGetLastError();                         \
      return GSR_ERR_CUDA;                              \
    }                                                   \
  } while (0)

extern "C" int gsr_version(void) { return GSR_VERSION; }

extern "C" const char* gsr_status_string(int status) {
  switch (status) {
    case GSR_OK: return "ok";
    case GSR_ERR_NULL_POINTER: return "null pointer argument";
    case GSR_ERR_BAD_SHAPE: return "bad shape: need s >= 0 and 2 <= h, w <= 32767";
    case GSR_ERR_BAD_CHANNELS: return "bad channel count: the rasteriser supports c == 3 only";
    case GSR_ERR_WORKSPACE: return "workspace missing, misaligned (256 B) or too small";
    case GSR_ERR_BAD_ARGUMENT: return "bad argument";
    case GSR_ERR_CUDA: return "CUDA runtime error (see gsr_last_cuda_error)";
    default: return "unknown status";
  }
}

extern "C" int gsr_last_cuda_error(void) { return g_last_cuda_error; }

static bool gsr_dims_ok(int s, int h, int w) {
  return s >= 0 && h >= 2 && w >= 2 && h <= GSR_MAX_DIM && w <= GSR_MAX_DIM;
}

extern "C" size_t gsr_workspace_bytes(int s, int h, int w) {
  if (!gsr_dims_ok(s, h, w)) return 0;
  return gsr_carve(nullptr, s, h, w).bytes;
}

static int gsr_check_ws(void* workspace, size_t bytes, size_t need) {
  if (!workspace || ((uintptr_t)workspace & 255u) || bytes < need) return GSR_ERR_WORKSPACE;
  return GSR_OK;
}


// Resident grid of a persistent kernel on the current device: SM count x CTAs per SM.  Queried once
// per (device, kernel) and cached -- immutable facts of the device; the cache is a table of relaxed atomics
// (every thread that races to fill an entry writes the same value).
static int gsr_sm_count(int* out) {
  static std::atomic<int> cache[64];
  int dev = 0;
  GSR_CUDA(cudaGetDevice(&dev));
  int v = (dev >= 0 && dev < 64) ? cache[dev].load(std::memory_order_relaxed) : 0;
  if (v <= 0) {
    GSR_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < 64) cache[dev].store(v, std::memory_order_relaxed);
  }
  *out = v;
  return GSR_OK;
}

template <typename K>
static int gsr_resident_grid(K kernel, int threads, int slot, int* out, size_t dyn_smem = 0) {
  static std::atomic<int> cache[4][64];  // [kernel slot][device], 0 = not yet queried
  int dev = 0;
  GSR_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64) {
    const int v = cache[slot][dev].load(std::memory_order_relaxed);
    if (v > 0) {
      *out = v;
      return GSR_OK;
    }
  }
  int nsm = 0, per_sm = 0;
  const int rc = gsr_sm_count(&nsm);
  if (rc) return rc;
  GSR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem));
  *out = nsm * (per_sm > 0 ? per_sm : 1);
  if (dev >= 0 && dev < 64) cache[slot][dev].store(*out, std::memory_order_relaxed);
  return GSR_OK;
}

// Kernels that need more than 48 KB of dynamic shared memory opt in once per (device, kernel).
template <typename K>
static int gsr_optin_smem(K kernel, int bytes, int slot) {
  static std::atomic<int> done[8][64];
  int dev = 0;
  GSR_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[slot][dev].load(std::memory_order_relaxed)) return GSR_OK;
  GSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (dev >= 0 && dev < 64) done[slot][dev].store(1, std::memory_order_relaxed);
  return GSR_OK;
}

static int gsr_clear_and_tables(int h, int w, const GsrWorkspace& ws, cudaStream_t st) {
  // one launch: the tables and the zeroing of the counter block (carved in multiples of 256 bytes)
  const int n = w > h ? w : h;
  const size_t zero16 = gsr_align_up(ws.zero_bytes, 256) / 16;
  const size_t want = (zero16 + 255) / 256 > (size_t)((n + 255) / 256) ? (zero16 + 255) / 256 : (size_t)((n + 255) / 256);
  gsr_table_kernel<<<(int)(want < 1184 ? want : 1184), 256, 0, st>>>(ws.px_tab, ws.py_tab, h, w, ws.hf, ws.row0, ws.bhs,
                                                                    reinterpret_cast<uint4*>(ws.bin_count), zero16);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Home-bin pipeline K1..K3 (leaves the sorted arrays in ws).  guard/want: see gsr_guard_skip.
static int gsr_run_bins(const float* sigmas, const float* coords, const float* colors, int s,
                        int h, int w, float dmax, float keff, const GsrWorkspace& ws,
                        const int* guard, int want, cudaStream_t st) {
  // guarded (fallback) launches use a small grid-stride grid: the no-op case must stay cheap
  const int gs_grid = guard ? (s + 255) / 256 < 592 ? (s + 255) / 256 : 592 : (s + 255) / 256;
  if (s > 0) {
    if (ws.ragged) gsr_bin_kernel<true><<<gs_grid, 256, 0, st>>>(sigmas, coords, colors, s, h, w, dmax, keff, ws, guard, want);
    else gsr_bin_kernel<false><<<gs_grid, 256, 0, st>>>(sigmas, coords, colors, s, h, w, dmax, keff, ws, guard, want);
  }
  gsr_scan_kernel<<<ws.nscan, 1024, 0, st>>>(ws.bin_count, ws.bin_off, nullptr, ws.nb + 1,
                                             ws.scan_state, -1, ws.stats, guard, want);
  if (s > 0)
    gsr_scatter_kernel<<<gs_grid, 256, 0, st>>>(sigmas, coords, colors, s, ws, guard, want);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Region-bucket pipeline (one kernel).  Raises stats[GSR_STAT_OVERFLOW] when a bucket overflows.
// raw != NULL: the kernel is fed with the raw head output (s,9) and applies the front end itself, leaving the
// mapped parameters in `mapped` (= sigmas / coords / colors, which then need not be initialised).
static int gsr_run_tiles(const float* sigmas, const float* coords, const float* colors, int s,
                         int h, int w, float dmax, float keff, const GsrWorkspace& ws,
                         cudaStream_t st, const float* raw = nullptr, float step = 0.f) {
  if (s > GSR_BUCKET_MAX_S) {
    // the bucket entries hold 23-bit indices: larger calls take the home-bin path (flag = 1, little endian)
    if (raw) {
      gsr_map_kernel<<<(s + 255) / 256, 256, 0, st>>>(raw, (float*)sigmas, (float*)coords, (float*)colors, s, h, w, step,
                                                      ws.ragged ? ws.bdesc : nullptr, ws.bn);
      GSR_CUDA(cudaGetLastError());
    }
    GSR_CUDA(cudaMemsetAsync(ws.stats + GSR_STAT_OVERFLOW, 1, 1, st));
    return GSR_OK;
  }
  if (s > 0) {
    const float ec = gsr_ecut(keff);
    float *ms = (float*)sigmas, *mc = (float*)coords, *mk = (float*)colors;  // RAW: outputs
    // one warp per 32 Gaussians (warps stride over the chunks only when the grid limit is hit)
    const long long want = ((long long)s + GSR_RB2_THREADS - 1) / GSR_RB2_THREADS;
    const int grid = (int)(want < (1 << 20) ? want : (1 << 20));
#define GSR_RB_LAUNCH(RG, RW, a0, a1, a2, o0, o1, o2, stp) \
  gsr_region_build2_kernel<RG, RW><<<grid, GSR_RB2_THREADS, 0, st>>>(a0, a1, a2, o0, o1, o2, s, h, w, dmax, keff, ec, stp, ws)
    if (raw) {
      if (ws.ragged) GSR_RB_LAUNCH(true, true, raw, nullptr, nullptr, ms, mc, mk, step);
      else GSR_RB_LAUNCH(false, true, raw, nullptr, nullptr, ms, mc, mk, step);
    } else if (ws.ragged) {
      GSR_RB_LAUNCH(true, false, sigmas, coords, colors, nullptr, nullptr, nullptr, 0.f);
    } else {
      GSR_RB_LAUNCH(false, false, sigmas, coords, colors, nullptr, nullptr, nullptr, 0.f);
    }
#undef GSR_RB_LAUNCH
  }
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

static GsrFwdArgs gsr_fwd_args(const GsrWorkspace& ws, float* img, int h, int w, float keff,
                               uint32_t flags) {
  GsrFwdArgs a;
  a.rec = ws.rec;
  a.box = ws.box;
  a.bin_off = ws.bin_off;
  a.stats = ws.stats;
  a.px_tab = ws.px_tab;
  a.py_tab = ws.py_tab;
  a.img = img;
  a.h = h;
  a.w = w;
  a.nbx = ws.nbx;
  a.nby = ws.nby;
  a.nb = ws.nb;
  a.ecut = gsr_ecut(keff);
  a.flags = flags;
  a.guard = ws.stats + GSR_STAT_OVERFLOW;
  a.want = 0;
  a.reg_count = ws.reg_count;
  a.reg_cap = ws.reg_cap;
  a.nrx = ws.nrx;
  a.nry = ws.nry;
  a.entries = ws.entries;
  a.rec_in = ws.rec_in;
  a.box_in = ws.box_in;
  a.sched = ws.stats + GSR_STAT_UNIT;
  a.hf = ws.hf;
  a.row0 = ws.row0;
  a.bhs = ws.bn > 0 ? ws.bhs : 0;
  if (flags & GSR_FLAG_CHW) {
    a.row_stride = w;
    a.pix_stride = 1;
    a.chan_stride = (long long)h * w;
  } else {
    a.row_stride = 3LL * w;
    a.pix_stride = 3;
    a.chan_stride = 1;
  }
  a.nclip = 0;
  if (ws.win) {  // a window of a larger destination (gsr_forward_window)
    a.row_stride = ws.win->row_stride;
    a.pix_stride = ws.win->pix_stride;
    a.chan_stride = ws.win->chan_stride;
    a.nclip = ws.win->nclip;
    for (int k = 0; k < a.nclip; ++k)
      for (int j = 0; j < 4; ++j) a.clip[k][j] = ws.win->clip[k][j];
  }
  return a;
}

// Raster over the region buckets (runs when every bucket fitted) ...
// Two kernels share the evaluation loop.  gsr_forward_region_kernel (every warp stages and evaluates) wins when a
// warp has many regions to walk (HL: 18 per warp, 283 vs 303 us; C3: 405 vs 440 us); the warp-specialised
// gsr_forward_region_ws_kernel (producer / consumer warp pairs, gsr_forward_ws.cuh) wins when there are only a few
// regions per warp, where its deeper staging pipeline and coarser warp count balance better (C2: 84 -> 62 us,
// C2d: 245 -> 162 us).  GSR_FR_WS=0/1 in the environment forces one of them (tuning aid, read once).
#ifndef GSR_CFG_FR_WS_UNITS_PER_WARP
#define GSR_CFG_FR_WS_UNITS_PER_WARP 12
#endif
static int gsr_ws_override() {
  static const int v = [] {
    const char* e = getenv("GSR_FR_WS");
    return e && (e[0] == '0' || e[0] == '1') ? e[0] - '0' : -1;
  }();
  return v;
}
static int gsr_launch_forward_region(const GsrWorkspace& ws, float* img, int h, int w, float keff,
                                   uint32_t flags, cudaStream_t st) {
  GsrFwdArgs a = gsr_fwd_args(ws, img, h, w, keff, flags);
  a.want = 0;
  const int nunits = ws.nrx * ws.nry;
  int cap = 0;
  int rc = gsr_resident_grid(gsr_forward_region_kernel<false>, GSR_FR_THREADS, 1, &cap);
  if (rc) return rc;
  const int ov = gsr_ws_override();
  const bool use_ws = ov >= 0 ? ov != 0 : (long long)nunits < (long long)GSR_CFG_FR_WS_UNITS_PER_WARP * cap * GSR_FR_WARPS;
  if (use_ws) {
    // persistent warp pairs: one resident wave, every producer warp takes regions from the work counter
    rc = gsr_optin_smem(gsr_forward_region_ws_kernel<false>, (int)sizeof(GsrWsSmem), 2);
    if (rc) return rc;
    rc = gsr_optin_smem(gsr_forward_region_ws_kernel<true>, (int)sizeof(GsrWsSmem), 3);
    if (rc) return rc;
    rc = gsr_resident_grid(gsr_forward_region_ws_kernel<false>, GSR_WS_THREADS, 2, &cap, sizeof(GsrWsSmem));
    if (rc) return rc;
    const int want = (nunits + GSR_WS_PAIRS - 1) / GSR_WS_PAIRS, grid = want < cap ? want : cap;
    if (ws.win) gsr_forward_region_ws_kernel<true><<<grid, GSR_WS_THREADS, sizeof(GsrWsSmem), st>>>(a);
    else gsr_forward_region_ws_kernel<false><<<grid, GSR_WS_THREADS, sizeof(GsrWsSmem), st>>>(a);
  } else {
    // persistent warps: one resident wave, every warp takes regions from the work counter
    const int want = (nunits + GSR_FR_WARPS - 1) / GSR_FR_WARPS;
    if (ws.win) gsr_forward_region_kernel<true><<<want < cap ? want : cap, GSR_FR_THREADS, 0, st>>>(a);
    else gsr_forward_region_kernel<false><<<want < cap ? want : cap, GSR_FR_THREADS, 0, st>>>(a);
  }
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// ... or over the home bins (runs when they overflowed).
static int gsr_launch_forward_bins(const GsrWorkspace& ws, float* img, int h, int w, float keff,
                                   uint32_t flags, cudaStream_t st) {
  GsrFwdArgs a = gsr_fwd_args(ws, img, h, w, keff, flags);
  a.want = 1;
  int rc = gsr_optin_smem(gsr_forward_bins_kernel, (int)sizeof(GsrFwdSmem), 0);
  if (rc) return rc;
  int nsm = 0;
  rc = gsr_sm_count(&nsm);
  if (rc) return rc;
  const int ntiles = ((w + GSR_TILE_W - 1) / GSR_TILE_W) * ((h + GSR_TILE_H - 1) / GSR_TILE_H);
  const int grid = ntiles < nsm * GSR_CFG_MIN_CTAS ? ntiles : nsm * GSR_CFG_MIN_CTAS;
  gsr_forward_bins_kernel<<<grid, GSR_FWD_THREADS, sizeof(GsrFwdSmem), st>>>(a);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Whole forward set-up: the region buckets (the home-bin fallback is part of the fallback raster kernel).
static int gsr_prepare_forward(const float* sigmas, const float* coords, const float* colors, int s,
                               int h, int w, float dmax, float keff, const GsrWorkspace& ws,
                               cudaStream_t st, const float* raw = nullptr, float step = 0.f) {
  return gsr_run_tiles(sigmas, coords, colors, s, h, w, dmax, keff, ws, st, raw, step);
}

#ifndef GSR_CFG_FALLBACK_COOP
#define GSR_CFG_FALLBACK_COOP 1
#endif
// The fallback as one launch: home-bin set-up + raster over the home bins, all of it skipped on the device unless
// a bucket overflowed (gsr_forward_fallback_kernel).
static int gsr_launch_forward_fallback(const GsrWorkspace& ws, float* img, int h, int w, float keff, uint32_t flags,
                                       cudaStream_t st, const float* sigmas, const float* coords,
                                       const float* colors, int s, float dmax) {
  GsrFwdArgs a = gsr_fwd_args(ws, img, h, w, keff, flags);
  a.want = 1;
  int rc = gsr_optin_smem(gsr_forward_fallback_kernel<false>, (int)sizeof(GsrFwdSmem), 4);
  if (rc) return rc;
  rc = gsr_optin_smem(gsr_forward_fallback_kernel<true>, (int)sizeof(GsrFwdSmem), 5);
  if (rc) return rc;
  int cap = 0;  // every CTA must be resident: the phases are separated by grid barriers
  rc = gsr_resident_grid(gsr_forward_fallback_kernel<false>, GSR_FWD_THREADS, 3, &cap, sizeof(GsrFwdSmem));
  if (rc) return rc;
  // A COOPERATIVE launch: the driver starts the grid only when all of its CTAs can be resident at once, so the
  // barriers cannot deadlock against other work on the device (e.g. a second forward call on another stream).
  GsrWorkspace wsv = ws;
  void* args[] = {(void*)&a, (void*)&sigmas, (void*)&coords, (void*)&colors, (void*)&s, (void*)&dmax, (void*)&keff,
                  (void*)&wsv};
  const void* fn = ws.ragged ? (const void*)gsr_forward_fallback_kernel<true> : (const void*)gsr_forward_fallback_kernel<false>;
#if GSR_CFG_FALLBACK_COOP
  GSR_CUDA(cudaLaunchCooperativeKernel(fn, dim3(cap), dim3(GSR_FWD_THREADS), args, sizeof(GsrFwdSmem), st));
#else
  GSR_CUDA(cudaLaunchKernel(fn, dim3(cap), dim3(GSR_FWD_THREADS), args, sizeof(GsrFwdSmem), st));
#endif
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// GSR_FLAG_DETERMINISTIC: every bucket sorted by Gaussian index before it is rasterised (idempotent: a prepared
// workspace may be rasterised any number of times).
static int gsr_sort_buckets(const GsrWorkspace& ws, cudaStream_t st) {
  int nsm = 0;
  const int rc = gsr_sm_count(&nsm);
  if (rc) return rc;
  const int* guard = ws.stats + GSR_STAT_OVERFLOW;
  const int want = (ws.nreg + GSR_SORT_WARPS - 1) / GSR_SORT_WARPS, cap = nsm * 16;
  gsr_bucket_sort_kernel<<<want < cap ? want : cap, 32 * GSR_SORT_WARPS, 0, st>>>(ws.entries, ws.reg_count, ws.reg_cap,
                                                                                  ws.nreg, guard, 0);
  if (ws.reg_cap > GSR_SORT_MAX)  // only then can a bucket be longer than the warp sort takes
    gsr_bucket_sort_long_kernel<<<ws.nreg < nsm * 8 ? ws.nreg : nsm * 8, 256, 0, st>>>(
        ws.entries, ws.reg_count, ws.reg_cap, ws.nreg, guard, 0, ws.stats + GSR_STAT_UNSORTED);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// bins_ready: gsr_prepare has set up the home bins (split-phase API); otherwise (one-call forward) the fallback
// kernel does it when needed.
static int gsr_raster_forward(const GsrWorkspace& ws, float* img, int h, int w, float keff,
                              uint32_t flags, cudaStream_t st, bool bins_ready, const float* sigmas = nullptr,
                              const float* coords = nullptr, const float* colors = nullptr, int s = 0,
                              float dmax = 0.f) {
  int rc = (flags & GSR_FLAG_DETERMINISTIC) ? gsr_sort_buckets(ws, st) : GSR_OK;
  if (rc) return rc;
  rc = gsr_launch_forward_region(ws, img, h, w, keff, flags, st);
  if (rc) return rc;
  if (bins_ready) return gsr_launch_forward_bins(ws, img, h, w, keff, flags, st);
  return gsr_launch_forward_fallback(ws, img, h, w, keff, flags, st, sigmas, coords, colors, s, dmax);
}

static int gsr_launch_backward(const GsrWorkspace& ws, const float* sigmas, const float* grads,
                               float* gs, float* gc, float* gk, int s, int h, int w,
                               uint32_t flags, cudaStream_t st, bool guarded = false) {
  if (s == 0) return GSR_OK;
  GsrBwdArgs a;
  a.guard = guarded ? ws.stats + GSR_STAT_OVERFLOW : nullptr;
  a.want = 1;
  a.rec = ws.rec;
  a.box = ws.box;
  a.ids = ws.ids;
  a.bin_off = ws.bin_off;
  a.stats = ws.stats;
  a.px_tab = ws.px_tab;
  a.py_tab = ws.py_tab;
  a.grads = grads;
  a.sigmas = sigmas;
  a.g_sigmas = gs;
  a.g_coords = gc;
  a.g_colors = gk;
  a.h = h;
  a.w = w;
  a.nbx = ws.nbx;
  a.nby = ws.nby;
  a.nb = ws.nb;
  a.tiles_x = (w + GSR_BWD_TILE - 1) / GSR_BWD_TILE;
  a.tiles_y = (h + GSR_BWD_TILE - 1) / GSR_BWD_TILE;
  a.flags = flags;
  a.bdesc = ws.bdesc;
  a.bn = ws.bn;
  a.ragged = ws.ragged;
  a.hf = ws.hf;
  a.row0 = ws.row0;
  a.bhs = ws.bn > 0 ? ws.bhs : 0;
  const int rcs = gsr_optin_smem(gsr_backward_kernel, (int)sizeof(GsrBwdSmem), 1);
  if (rcs) return rcs;
  const int nblocks = a.tiles_x * a.tiles_y + (s + GSR_BWD_LARGE_CHUNK - 1) / GSR_BWD_LARGE_CHUNK;
  int grid = nblocks;
  if (guarded) {  // normally a no-op: keep it cheap
    int nsm = 0;
    const int rc = gsr_sm_count(&nsm);
    if (rc) return rc;
    grid = nblocks < 2 * nsm ? nblocks : 2 * nsm;
  }
  gsr_backward_kernel<<<grid, GSR_BWD_THREADS, sizeof(GsrBwdSmem), st>>>(a, nblocks);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Backward over the region buckets (gsr_backward_region.cuh) + the chain rule per Gaussian; both skipped on the
// device when a bucket overflowed (the Gaussian-centric kernel over the home bins then runs instead).
#ifndef GSR_CFG_BWD_REGION
#define GSR_CFG_BWD_REGION 1
#endif
static int gsr_bwd_region_override() {
  static const int v = [] {
    const char* e = getenv("GSR_BWD_REGION");
    return e && (e[0] == '0' || e[0] == '1') ? e[0] - '0' : -1;
  }();
  return v;
}
static bool gsr_bwd_use_region(uint32_t flags) {
  if (flags & GSR_FLAG_DETERMINISTIC) return false;  // the atomics' order is not reproducible
  const int ov = gsr_bwd_region_override();
  return ov >= 0 ? ov != 0 : GSR_CFG_BWD_REGION != 0;
}
static int gsr_clear_moments(const GsrWorkspace& ws, int s, cudaStream_t st) {
  if (s > 0) GSR_CUDA(cudaMemsetAsync(ws.mom, 0, (size_t)s * 8 * sizeof(float), st));
  return GSR_OK;
}
static int gsr_launch_backward_region(const GsrWorkspace& ws, const float* sigmas, const float* grads, float* gs,
                                      float* gc, float* gk, int s, int h, int w, float keff, uint32_t flags,
                                      cudaStream_t st) {
  if (s == 0) return GSR_OK;
  GsrFwdArgs a = gsr_fwd_args(ws, nullptr, h, w, keff, flags);
  a.want = 0;
  GsrBwdRegionArgs q;
  q.grads = grads;
  q.mom = ws.mom;
  const int nunits = ws.nrx * ws.nry;
  int cap = 0;
  int rc = gsr_optin_smem(gsr_backward_region_kernel, (int)sizeof(GsrBwdRegionSmem), 6);
  if (rc) return rc;
  rc = gsr_resident_grid(gsr_backward_region_kernel, GSR_FR_THREADS, 0, &cap, sizeof(GsrBwdRegionSmem));
  if (rc) return rc;
  const int want = (nunits + GSR_FR_WARPS - 1) / GSR_FR_WARPS;
  gsr_backward_region_kernel<<<want < cap ? want : cap, GSR_FR_THREADS, sizeof(GsrBwdRegionSmem), st>>>(a, q);
  const int cg = (s + 255) / 256;
  gsr_bwd_chain_kernel<<<cg < 1184 ? cg : 1184, 256, 0, st>>>(ws.mom, ws.rec_in, sigmas, gs, gc, gk, s, a.guard, 0,
                                                             ws.ragged ? ws.bdesc : nullptr, ws.bn);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Padded batches: the per-sample descriptors travel as KERNEL PARAMETERS (by value, 96 per launch), not through a
// copy from pageable host memory: nothing of the caller's host memory is referenced after the call returns and
// the call can be captured into a CUDA graph.
constexpr int GSR_BDESC_PER_LAUNCH = 96;
struct GsrBDescBlock {
  GsrBDesc d[GSR_BDESC_PER_LAUNCH];
};
static_assert(sizeof(GsrBDescBlock) <= 3500, "kernel parameter space");
__global__ void gsr_bdesc_store_kernel(GsrBDescBlock blk, GsrBDesc* __restrict__ dst, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = blk.d[threadIdx.x];
}
static int gsr_stage_bdesc(const GsrWorkspace& ws, const GsrBDesc* host, int n, cudaStream_t st) {
  for (int o = 0; o < n; o += GSR_BDESC_PER_LAUNCH) {
    GsrBDescBlock blk;
    const int m = n - o < GSR_BDESC_PER_LAUNCH ? n - o : GSR_BDESC_PER_LAUNCH;
    for (int k = 0; k < m; ++k) blk.d[k] = host[o + k];
    gsr_bdesc_store_kernel<<<1, GSR_BDESC_PER_LAUNCH, 0, st>>>(blk, ws.bdesc + o, m);
  }
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Rows [row0, row0 + h) of an hf-row image (hf = 0: the whole h-row image).
static bool gsr_band_ok(int h, int hf, int row0) {
  return hf == 0 || (hf >= 2 && hf <= GSR_MAX_DIM && row0 >= 0 && row0 + h <= hf);
}

static int gsr_forward_impl(const float* sigmas, const float* coords, const float* colors, float* img,
                            int s, int h, int w, int c, int hf, int row0, float dmax, float ksigma,
                            uint32_t flags, void* workspace, size_t workspace_bytes, void* stream,
                            int bn = 0, int bhs = 0, const gsr_window* win = nullptr,
                            const GsrBDesc* bdesc_host = nullptr, int nsamples = 0, const float* raw = nullptr,
                            float step = 0.f) {
  if (c != 3) return GSR_ERR_BAD_CHANNELS;
  if (!gsr_dims_ok(s, h, w) || !gsr_band_ok(h, hf, row0)) return GSR_ERR_BAD_SHAPE;
  if (!img || (s > 0 && (!sigmas || !coords || !colors))) return GSR_ERR_NULL_POINTER;
  if ((flags & GSR_FLAG_U8) && (!(flags & GSR_FLAG_OVERWRITE) || (flags & GSR_FLAG_CHW) || win))
    return GSR_ERR_BAD_ARGUMENT;  // uint8 output: written once, (h,w,3), whole image
  if ((flags & GSR_FLAG_BGR) && !(flags & GSR_FLAG_U8)) return GSR_ERR_BAD_ARGUMENT;
  GsrWorkspace ws = gsr_carve(workspace, s, h, w);
  ws.hf = hf;
  ws.row0 = row0;
  ws.bn = bn;
  ws.bhs = bhs;
  int rc = gsr_check_ws(workspace, workspace_bytes, ws.bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const float keff = gsr_effective_ksigma(ksigma);
  if (bdesc_host) {  // padded batch: per-sample descriptors
    ws.ragged = 1;
    rc = gsr_stage_bdesc(ws, bdesc_host, nsamples, st);
    if (rc) return rc;
  }
  rc = gsr_clear_and_tables(h, w, ws, st);
  if (rc) return rc;
  ws.win = win;
  rc = gsr_prepare_forward(sigmas, coords, colors, s, h, w, dmax, keff, ws, st, raw, step);
  if (rc) return rc;
  // (raw: the set-up kernel has left the mapped parameters in sigmas / coords / colors)
  return gsr_raster_forward(ws, img, h, w, keff, flags, st, false, sigmas, coords, colors, s, dmax);
}

static int gsr_backward_impl(const float* sigmas, const float* coords, const float* colors,
                             const float* grads, float* grads_sigmas, float* grads_coords,
                             float* grads_colors, int s, int h, int w, int c, int hf, int row0,
                             float dmax, float ksigma, uint32_t flags, void* workspace,
                             size_t workspace_bytes, void* stream, int bn = 0, int bhs = 0,
                             const GsrBDesc* bdesc_host = nullptr, int nsamples = 0) {
  if (c != 3) return GSR_ERR_BAD_CHANNELS;
  if (!gsr_dims_ok(s, h, w) || !gsr_band_ok(h, hf, row0)) return GSR_ERR_BAD_SHAPE;
  if (!grads || (s > 0 && (!sigmas || !coords || !colors || !grads_sigmas || !grads_coords ||
                           !grads_colors)))
    return GSR_ERR_NULL_POINTER;
  GsrWorkspace ws = gsr_carve(workspace, s, h, w);
  ws.hf = hf;
  ws.row0 = row0;
  ws.bn = bn;
  ws.bhs = bhs;
  int rc = gsr_check_ws(workspace, workspace_bytes, ws.bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const float keff = gsr_effective_ksigma(ksigma);
  if (bdesc_host) {
    ws.ragged = 1;
    rc = gsr_stage_bdesc(ws, bdesc_host, nsamples, st);
    if (rc) return rc;
  }
  rc = gsr_clear_and_tables(h, w, ws, st);
  if (rc) return rc;
  if (!gsr_bwd_use_region(flags)) {
    rc = gsr_run_bins(sigmas, coords, colors, s, h, w, dmax, keff, ws, nullptr, 0, st);
    if (rc) return rc;
    return gsr_launch_backward(ws, sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, flags, st);
  }
  // region buckets; if one overflows, the home-bin sort and the Gaussian-centric kernel (guarded launches)
  rc = gsr_clear_moments(ws, s, st);
  if (rc) return rc;
  rc = gsr_run_tiles(sigmas, coords, colors, s, h, w, dmax, keff, ws, st);
  if (rc) return rc;
  rc = gsr_launch_backward_region(ws, sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, keff, flags, st);
  if (rc) return rc;
  rc = gsr_run_bins(sigmas, coords, colors, s, h, w, dmax, keff, ws, ws.stats + GSR_STAT_OVERFLOW, 1, st);
  if (rc) return rc;
  return gsr_launch_backward(ws, sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, flags, st, true);
}

extern "C" int gsr_forward(const float* sigmas, const float* coords, const float* colors,
                           float* img, int s, int h, int w, int c, float dmax, float ksigma,
                           uint32_t flags, void* workspace, size_t workspace_bytes, void* stream) {
  return gsr_forward_impl(sigmas, coords, colors, img, s, h, w, c, 0, 0, dmax, ksigma, flags, workspace,
                          workspace_bytes, stream);
}

extern "C" int gsr_backward(const float* sigmas, const float* coords, const float* colors,
                            const float* grads, float* grads_sigmas, float* grads_coords,
                            float* grads_colors, int s, int h, int w, int c, float dmax,
                            float ksigma, uint32_t flags, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return gsr_backward_impl(sigmas, coords, colors, grads, grads_sigmas, grads_coords, grads_colors, s, h, w,
                           c, 0, 0, dmax, ksigma, flags, workspace, workspace_bytes, stream);
}

// ---- render into a window of a larger destination ---------------------------------------------
extern "C" int gsr_forward_window(const float* sigmas, const float* coords, const float* colors,
                                  float* origin, const gsr_window* win, int s, int h, int w, int c,
                                  float dmax, float ksigma, uint32_t flags, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!win) return GSR_ERR_NULL_POINTER;
  if (win->nclip < 0 || win->nclip > GSR_MAX_CLIP || (flags & GSR_FLAG_CHW)) return GSR_ERR_BAD_ARGUMENT;
  return gsr_forward_impl(sigmas, coords, colors, origin, s, h, w, c, 0, 0, dmax, ksigma, flags, workspace,
                          workspace_bytes, stream, 0, 0, win);
}

// ---- row bands of one image (multi-GPU split of a single large image) ---------------------------
extern "C" int gsr_forward_band(const float* sigmas, const float* coords, const float* colors,
                                float* img_band, int s, int h, int w, int c, int row0, int rows,
                                float dmax, float ksigma, uint32_t flags, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (h < 2 || h > GSR_MAX_DIM) return GSR_ERR_BAD_SHAPE;
  return gsr_forward_impl(sigmas, coords, colors, img_band, s, rows, w, c, h, row0, dmax, ksigma, flags,
                          workspace, workspace_bytes, stream);
}

extern "C" int gsr_backward_band(const float* sigmas, const float* coords, const float* colors,
                                 const float* grads_band, float* grads_sigmas, float* grads_coords,
                                 float* grads_colors, int s, int h, int w, int c, int row0, int rows,
                                 float dmax, float ksigma, uint32_t flags, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (h < 2 || h > GSR_MAX_DIM) return GSR_ERR_BAD_SHAPE;
  return gsr_backward_impl(sigmas, coords, colors, grads_band, grads_sigmas, grads_coords, grads_colors, s,
                           rows, w, c, h, row0, dmax, ksigma, flags, workspace, workspace_bytes, stream);
}

// ---- uniform batches: B samples of the same shape rendered as ONE stacked image ------------------
// Samples whose height is a multiple of the region height (8) are stacked into one (B*h, w) image --
// every Gaussian is set up in its own sample's coordinates and moved to the sample's block of rows --
// so a whole training batch takes one set-up and one raster launch (and one backward launch) instead
// of B of each.  Stacks are cut to GSR_MAX_DIM rows; other shapes (and the CHW layout, whose batch
// stride differs) take one call per sample.
static int gsr_batch_group(int batch, int h, uint32_t flags) {
  if (h % GSR_REGION != 0 || (flags & GSR_FLAG_CHW)) return 1;
  int g = GSR_MAX_DIM / h;  // samples that fit one stack ...
  if (g < 1 || batch < 1) return 1;
  const int launches = (batch + g - 1) / g;
  g = (batch + launches - 1) / launches;  // ... dealt evenly over the launches (32 x 1024 rows: 16 + 16)
  return g;
}

extern "C" size_t gsr_workspace_bytes_batch_uniform(int batch, int s_per, int h, int w) {
  if (batch < 0 || !gsr_dims_ok(s_per, h, w)) return 0;
  if (batch == 0) return 256;
  const int g = gsr_batch_group(batch, h, 0);
  if ((long long)g * s_per > 0x7fffffffLL) return 0;
  const size_t a = gsr_workspace_bytes(g * s_per, g * h, w), b = gsr_workspace_bytes(s_per, h, w);
  return a > b ? a : b;
}

static int gsr_forward_batch_uniform_impl(const float* sigmas, const float* coords, const float* colors,
                                          float* imgs, int batch, int s_per, int h, int w, int c,
                                          float dmax, float ksigma, uint32_t flags, void* workspace,
                                          size_t workspace_bytes, void* stream, const float* raw = nullptr,
                                          float step = 0.f) {
  if (batch < 0) return GSR_ERR_BAD_ARGUMENT;
  if (!gsr_dims_ok(s_per, h, w)) return GSR_ERR_BAD_SHAPE;
  if (batch > 0 && !imgs) return GSR_ERR_NULL_POINTER;
  const int g = gsr_batch_group(batch, h, flags);
  for (int b0 = 0; b0 < batch; b0 += g) {
    const int nb = batch - b0 < g ? batch - b0 : g;
    const size_t go = (size_t)b0 * s_per;
    const int rc = gsr_forward_impl(sigmas ? sigmas + 3 * go : nullptr, coords ? coords + 2 * go : nullptr,
                                    colors ? colors + 3 * go : nullptr,
                                    (flags & GSR_FLAG_U8) ? (float*)((unsigned char*)imgs + (size_t)b0 * h * w * 3)
                                                          : imgs + (size_t)b0 * h * w * 3,
                                    nb * s_per,
                                    nb * h, w, c, 0, 0, dmax, ksigma, flags, workspace, workspace_bytes, stream,
                                    nb > 1 ? s_per : 0, nb > 1 ? h : 0, nullptr, nullptr, 0,
                                    raw ? raw + 9 * go : nullptr, step);
    if (rc) return rc;
  }
  return GSR_OK;
}

extern "C" int gsr_forward_batch_uniform(const float* sigmas, const float* coords, const float* colors,
                                         float* imgs, int batch, int s_per, int h, int w, int c,
                                         float dmax, float ksigma, uint32_t flags, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  return gsr_forward_batch_uniform_impl(sigmas, coords, colors, imgs, batch, s_per, h, w, c, dmax, ksigma, flags,
                                        workspace, workspace_bytes, stream);
}

extern "C" int gsr_backward_batch_uniform(const float* sigmas, const float* coords, const float* colors,
                                          const float* grads, float* grads_sigmas, float* grads_coords,
                                          float* grads_colors, int batch, int s_per, int h, int w, int c,
                                          float dmax, float ksigma, uint32_t flags, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  if (batch < 0) return GSR_ERR_BAD_ARGUMENT;
  if (!gsr_dims_ok(s_per, h, w)) return GSR_ERR_BAD_SHAPE;
  if (batch > 0 && !grads) return GSR_ERR_NULL_POINTER;
  const int g = gsr_batch_group(batch, h, flags);
  for (int b0 = 0; b0 < batch; b0 += g) {
    const int nb = batch - b0 < g ? batch - b0 : g;
    const size_t go = (size_t)b0 * s_per;
    const int rc = gsr_backward_impl(sigmas ? sigmas + 3 * go : nullptr, coords ? coords + 2 * go : nullptr,
                                     colors ? colors + 3 * go : nullptr, grads + (size_t)b0 * h * w * 3,
                                     grads_sigmas ? grads_sigmas + 3 * go : nullptr,
                                     grads_coords ? grads_coords + 2 * go : nullptr,
                                     grads_colors ? grads_colors + 3 * go : nullptr, nb * s_per, nb * h, w, c, 0, 0,
                                     dmax, ksigma, flags, workspace, workspace_bytes, stream, nb > 1 ? s_per : 0,
                                     nb > 1 ? h : 0);
    if (rc) return rc;
  }
  return GSR_OK;
}

// ---- split-phase form ---------------------------------------------------------------------------
extern "C" int gsr_prepare(const float* sigmas, const float* coords, const float* colors, int s,
                           int h, int w, float dmax, float ksigma, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!gsr_dims_ok(s, h, w)) return GSR_ERR_BAD_SHAPE;
  if (s > 0 && (!sigmas || !coords || !colors)) return GSR_ERR_NULL_POINTER;
  const GsrWorkspace ws = gsr_carve(workspace, s, h, w);
  int rc = gsr_check_ws(workspace, workspace_bytes, ws.bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const float keff = gsr_effective_ksigma(ksigma);
  rc = gsr_clear_and_tables(h, w, ws, st);
  if (rc) return rc;
  rc = gsr_run_tiles(sigmas, coords, colors, s, h, w, dmax, keff, ws, st);
  if (rc) return rc;
  return gsr_run_bins(sigmas, coords, colors, s, h, w, dmax, keff, ws, nullptr, 0, st);
}

extern "C" int gsr_forward_prepared(float* img, int s, int h, int w, float ksigma, uint32_t flags,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (!gsr_dims_ok(s, h, w)) return GSR_ERR_BAD_SHAPE;
  if (!img) return GSR_ERR_NULL_POINTER;
  const GsrWorkspace ws = gsr_carve(workspace, s, h, w);
  int rc = gsr_check_ws(workspace, workspace_bytes, ws.bytes);
  if (rc) return rc;
  return gsr_raster_forward(ws, img, h, w, gsr_effective_ksigma(ksigma), flags, (cudaStream_t)stream, true);
}

extern "C" int gsr_backward_prepared(const float* sigmas, const float* grads, float* grads_sigmas,
                                     float* grads_coords, float* grads_colors, int s, int h, int w,
                                     uint32_t flags, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (!gsr_dims_ok(s, h, w)) return GSR_ERR_BAD_SHAPE;
  if (!grads || (s > 0 && (!sigmas || !grads_sigmas || !grads_coords || !grads_colors)))
    return GSR_ERR_NULL_POINTER;
  const GsrWorkspace ws = gsr_carve(workspace, s, h, w);
  int rc = gsr_check_ws(workspace, workspace_bytes, ws.bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!gsr_bwd_use_region(flags))
    return gsr_launch_backward(ws, sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, flags, st);
  rc = gsr_clear_moments(ws, s, st);
  if (rc) return rc;
  const float keff = gsr_effective_ksigma(0.f);  // (not used by the region backward: the buckets are already built)
  rc = gsr_launch_backward_region(ws, sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, keff, flags, st);
  if (rc) return rc;
  return gsr_launch_backward(ws, sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, flags, st, true);
}

// ---- ragged batches -------------------------------------------------------------------------
extern "C" size_t gsr_workspace_bytes_batch(const gsr_sample* samples, int n) {
  if (!samples || n < 0) return 0;
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    const size_t b = gsr_workspace_bytes(samples[i].s, samples[i].h, samples[i].w);
    if (b == 0) return 0;
    total += b;
  }
  return total;
}

extern "C" int gsr_forward_batch(const gsr_sample* samples, int n, float ksigma, uint32_t flags,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0) return GSR_ERR_BAD_ARGUMENT;
  if (n > 0 && !samples) return GSR_ERR_NULL_POINTER;
  const size_t need = gsr_workspace_bytes_batch(samples, n);
  if (n > 0 && need == 0) return GSR_ERR_BAD_SHAPE;
  int rc = n > 0 ? gsr_check_ws(workspace, workspace_bytes, need) : GSR_OK;
  if (rc) return rc;
  char* base = (char*)workspace;
  for (int i = 0; i < n; ++i) {
    const gsr_sample& sm = samples[i];
    const size_t b = gsr_workspace_bytes(sm.s, sm.h, sm.w);
    rc = gsr_forward(sm.sigmas, sm.coords, sm.colors, sm.img, sm.s, sm.h, sm.w, 3, sm.dmax, ksigma,
                     flags, base, b, stream);
    if (rc) return rc;
    base += b;
  }
  return GSR_OK;
}

extern "C" int gsr_backward_batch(const gsr_sample* samples, int n, float ksigma, uint32_t flags,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 0) return GSR_ERR_BAD_ARGUMENT;
  if (n > 0 && !samples) return GSR_ERR_NULL_POINTER;
  const size_t need = gsr_workspace_bytes_batch(samples, n);
  if (n > 0 && need == 0) return GSR_ERR_BAD_SHAPE;
  int rc = n > 0 ? gsr_check_ws(workspace, workspace_bytes, need) : GSR_OK;
  if (rc) return rc;
  char* base = (char*)workspace;
  for (int i = 0; i < n; ++i) {
    const gsr_sample& sm = samples[i];
    const size_t b = gsr_workspace_bytes(sm.s, sm.h, sm.w);
    rc = gsr_backward(sm.sigmas, sm.coords, sm.colors, sm.grads, sm.grads_sigmas, sm.grads_coords,
                      sm.grads_colors, sm.s, sm.h, sm.w, 3, sm.dmax, ksigma, flags, base, b, stream);
    if (rc) return rc;
    base += b;
  }
  return GSR_OK;
}

// ---- fused front end --------------------------------------------------------------------------
extern "C" int gsr_frontend_forward(const float* raw, float* mapped, float* img_chw, int s, int h,
                                    int w, float step_size, float dmax, float ksigma,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (!gsr_dims_ok(s, h, w)) return GSR_ERR_BAD_SHAPE;
  if (!img_chw || (s > 0 && (!raw || !mapped))) return GSR_ERR_NULL_POINTER;
  if (!(step_size > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  // the set-up kernel applies the activations and the unit mapping itself and leaves the mapped parameters
  // in `mapped` for the backward: no separate elementwise pass, no second read of them
  return gsr_forward_impl(mapped, mapped + 3 * (size_t)s, mapped + 5 * (size_t)s, img_chw, s, h, w, 3, 0, 0, dmax,
                          ksigma, GSR_FLAG_OVERWRITE | GSR_FLAG_CHW, workspace, workspace_bytes, stream, 0, 0, nullptr,
                          nullptr, 0, raw, step_size);
}

extern "C" int gsr_frontend_backward(const float* raw, const float* mapped, const float* grads_chw,
                                     float* grad_raw, int s, int h, int w, float step_size,
                                     float dmax, float ksigma, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (!gsr_dims_ok(s, h, w)) return GSR_ERR_BAD_SHAPE;
  if (!grads_chw || (s > 0 && (!raw || !mapped || !grad_raw))) return GSR_ERR_NULL_POINTER;
  if (!(step_size > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  if (s == 0) return GSR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // grad_raw doubles as the accumulation buffer for the mapped-parameter gradients:
  // [0,3s) d/dsigmas, [3s,5s) d/dcoords, [5s,8s) d/dcolors, then rewritten in place per row.
  // The (s,9) output has 9s floats, the mapped gradients need 8s: they are staged at the END
  // of the workspace instead, so the chain-rule kernel can write grad_raw freely.
  const size_t need = gsr_workspace_bytes(s, h, w);
  const size_t stage = gsr_align_up((size_t)s * 8 * sizeof(float), 256);
  if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < need + stage)
    return GSR_ERR_WORKSPACE;
  float* gm = (float*)((char*)workspace + need);
  GSR_CUDA(cudaMemsetAsync(gm, 0, (size_t)s * 8 * sizeof(float), st));
  int rc = gsr_backward(mapped, mapped + 3 * (size_t)s, mapped + 5 * (size_t)s, grads_chw, gm,
                        gm + 3 * (size_t)s, gm + 5 * (size_t)s, s, h, w, 3, dmax, ksigma,
                        GSR_FLAG_CHW, workspace, need, stream);
  if (rc) return rc;
  gsr_unmap_kernel<<<(s + 255) / 256, 256, 0, st>>>(raw, gm, gm + 3 * (size_t)s, gm + 5 * (size_t)s,
                                                    grad_raw, s, h, w, step_size);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// ---- padded (ragged) batches: samples of DIFFERENT sizes in one launch ---------------------------
// The training loop renders every sample at its own scale (gsasr_model.py:191-233: sr_size=gt_size[i])
// and pads the results to the largest size.  Here the samples -- s_per Gaussians each -- are rendered into
// the (batch, hmax, wmax, 3) padded buffer directly, stacked like a uniform batch: every Gaussian is set
// up in its own sample's (h_b, w_b, dmax_b) image, and its record is rescaled to the canvas' coordinate
// normalisation (d_own = (W-1)/(w_b-1) * d_canvas), so the raster kernels need nothing per sample.
static int gsr_padded_group(int batch, int hmax) {
  int g = GSR_MAX_DIM / hmax;
  if (g > GSR_BDESC_MAX) g = GSR_BDESC_MAX;
  if (g < 1 || batch < 1) return 1;
  const int launches = (batch + g - 1) / g;
  return (batch + launches - 1) / launches;
}

static int gsr_padded_check(int batch, int s_per, int hmax, int wmax, const int* hw_host) {
  if (batch < 0) return GSR_ERR_BAD_ARGUMENT;
  if (!gsr_dims_ok(s_per, hmax, wmax) || hmax % GSR_REGION != 0 || (long long)batch * s_per > 0x7fffffffLL)
    return GSR_ERR_BAD_SHAPE;
  if (batch > 0 && !hw_host) return GSR_ERR_NULL_POINTER;
  for (int b = 0; b < batch; ++b)
    if (hw_host[2 * b] < 2 || hw_host[2 * b] > hmax || hw_host[2 * b + 1] < 2 || hw_host[2 * b + 1] > wmax)
      return GSR_ERR_BAD_SHAPE;
  return GSR_OK;
}

static void gsr_padded_fill(GsrBDesc* d, int nb, int hmax, int wmax, const int* hw, const float* dmax_host, float dmax,
                            const float* step_host = nullptr) {
  for (int b = 0; b < nb; ++b) {
    d[b].h = hw[2 * b];
    d[b].w = hw[2 * b + 1];
    d[b].dmax = dmax_host ? dmax_host[b] : dmax;
    d[b].step = step_host ? step_host[b] : 0.0f;
    d[b].ax = (double)(wmax - 1) / (double)(hw[2 * b + 1] - 1);
    d[b].ay = (double)(hmax - 1) / (double)(hw[2 * b] - 1);
  }
}

extern "C" size_t gsr_workspace_bytes_batch_padded(int batch, int s_per, int hmax, int wmax) {
  if (batch < 0 || !gsr_dims_ok(s_per, hmax, wmax) || hmax % GSR_REGION != 0) return 0;
  if (batch == 0) return 256;
  const int g = gsr_padded_group(batch, hmax);
  if ((long long)g * s_per > 0x7fffffffLL) return 0;
  return gsr_workspace_bytes(g * s_per, g * hmax, wmax);
}

extern "C" int gsr_forward_batch_padded(const float* sigmas, const float* coords, const float* colors,
                                        float* imgs, int batch, int s_per, int hmax, int wmax,
                                        const int* hw_host, const float* dmax_host, float dmax,
                                        float ksigma, uint32_t flags, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  int rc = gsr_padded_check(batch, s_per, hmax, wmax, hw_host);
  if (rc) return rc;
  if (flags & (GSR_FLAG_CHW | GSR_FLAG_U8)) return GSR_ERR_BAD_ARGUMENT;
  if (batch > 0 && !imgs) return GSR_ERR_NULL_POINTER;
  const int g = gsr_padded_group(batch, hmax);
  GsrBDesc desc[GSR_BDESC_MAX];
  for (int b0 = 0; b0 < batch; b0 += g) {
    const int nb = batch - b0 < g ? batch - b0 : g;
    const size_t go = (size_t)b0 * s_per;
    gsr_padded_fill(desc, nb, hmax, wmax, hw_host + 2 * b0, dmax_host ? dmax_host + b0 : nullptr, dmax);
    rc = gsr_forward_impl(sigmas ? sigmas + 3 * go : nullptr, coords ? coords + 2 * go : nullptr,
                          colors ? colors + 3 * go : nullptr, imgs + (size_t)b0 * hmax * wmax * 3, nb * s_per,
                          nb * hmax, wmax, 3, 0, 0, dmax, ksigma, flags, workspace, workspace_bytes, stream, s_per,
                          hmax, nullptr, desc, nb);
    if (rc) return rc;
  }
  return GSR_OK;
}

extern "C" int gsr_backward_batch_padded(const float* sigmas, const float* coords, const float* colors,
                                         const float* grads, float* grads_sigmas, float* grads_coords,
                                         float* grads_colors, int batch, int s_per, int hmax, int wmax,
                                         const int* hw_host, const float* dmax_host, float dmax,
                                         float ksigma, uint32_t flags, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  int rc = gsr_padded_check(batch, s_per, hmax, wmax, hw_host);
  if (rc) return rc;
  if (flags & GSR_FLAG_CHW) return GSR_ERR_BAD_ARGUMENT;
  if (batch > 0 && !grads) return GSR_ERR_NULL_POINTER;
  const int g = gsr_padded_group(batch, hmax);
  GsrBDesc desc[GSR_BDESC_MAX];
  for (int b0 = 0; b0 < batch; b0 += g) {
    const int nb = batch - b0 < g ? batch - b0 : g;
    const size_t go = (size_t)b0 * s_per;
    gsr_padded_fill(desc, nb, hmax, wmax, hw_host + 2 * b0, dmax_host ? dmax_host + b0 : nullptr, dmax);
    rc = gsr_backward_impl(sigmas ? sigmas + 3 * go : nullptr, coords ? coords + 2 * go : nullptr,
                           colors ? colors + 3 * go : nullptr, grads + (size_t)b0 * hmax * wmax * 3,
                           grads_sigmas ? grads_sigmas + 3 * go : nullptr,
                           grads_coords ? grads_coords + 2 * go : nullptr,
                           grads_colors ? grads_colors + 3 * go : nullptr, nb * s_per, nb * hmax, wmax, 3, 0, 0,
                           dmax, ksigma, flags, workspace, workspace_bytes, stream, s_per, hmax, desc, nb);
    if (rc) return rc;
  }
  return GSR_OK;
}

// Fused front end for a uniform batch: raw (batch*s_per,9) -> imgs (batch,h,w,3) (HWC per sample).
extern "C" int gsr_frontend_forward_batch_uniform(const float* raw, float* mapped, float* imgs, int batch,
                                                  int s_per, int h, int w, float step_size, float dmax,
                                                  float ksigma, void* workspace, size_t workspace_bytes,
                                                  void* stream) {
  if (batch < 0) return GSR_ERR_BAD_ARGUMENT;
  if (!gsr_dims_ok(s_per, h, w) || (long long)batch * s_per > 0x7fffffffLL) return GSR_ERR_BAD_SHAPE;
  const int s = batch * s_per;
  if (!imgs || (s > 0 && (!raw || !mapped))) return GSR_ERR_NULL_POINTER;
  if (!(step_size > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  return gsr_forward_batch_uniform_impl(mapped, mapped + 3 * (size_t)s, mapped + 5 * (size_t)s, imgs, batch, s_per, h,
                                        w, 3, dmax, ksigma, GSR_FLAG_OVERWRITE, workspace, workspace_bytes, stream, raw,
                                        step_size);
}

// grads (batch,h,w,3) -> grad_raw (batch*s_per,9), written.  Workspace: gsr_workspace_bytes_batch_uniform
// + 32 bytes per Gaussian (rounded up to 256) for the mapped-parameter gradients.
extern "C" int gsr_frontend_backward_batch_uniform(const float* raw, const float* mapped, const float* grads,
                                                   float* grad_raw, int batch, int s_per, int h, int w,
                                                   float step_size, float dmax, float ksigma,
                                                   void* workspace, size_t workspace_bytes, void* stream) {
  if (batch < 0) return GSR_ERR_BAD_ARGUMENT;
  if (!gsr_dims_ok(s_per, h, w) || (long long)batch * s_per > 0x7fffffffLL) return GSR_ERR_BAD_SHAPE;
  const int s = batch * s_per;
  if (!grads || (s > 0 && (!raw || !mapped || !grad_raw))) return GSR_ERR_NULL_POINTER;
  if (!(step_size > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  if (s == 0) return GSR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = gsr_workspace_bytes_batch_uniform(batch, s_per, h, w);
  const size_t stage = gsr_align_up((size_t)s * 8 * sizeof(float), 256);
  if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < need + stage) return GSR_ERR_WORKSPACE;
  float* gm = (float*)((char*)workspace + need);
  GSR_CUDA(cudaMemsetAsync(gm, 0, (size_t)s * 8 * sizeof(float), st));
  int rc = gsr_backward_batch_uniform(mapped, mapped + 3 * (size_t)s, mapped + 5 * (size_t)s, grads, gm,
                                      gm + 3 * (size_t)s, gm + 5 * (size_t)s, batch, s_per, h, w, 3, dmax, ksigma,
                                      0, workspace, need, stream);
  if (rc) return rc;
  gsr_unmap_kernel<<<(s + 255) / 256, 256, 0, st>>>(raw, gm, gm + 3 * (size_t)s, gm + 5 * (size_t)s, grad_raw, s,
                                                    h, w, step_size);
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}

// Fused front end into a window: raw (s,9) -> activations + mapping -> render into the destination.
extern "C" int gsr_frontend_forward_window(const float* raw, float* mapped, float* origin,
                                           const gsr_window* win, int s, int h, int w, float step_size,
                                           float dmax, float ksigma, uint32_t flags, void* workspace,
                                           size_t workspace_bytes, void* stream) {
  if (!gsr_dims_ok(s, h, w)) return GSR_ERR_BAD_SHAPE;
  if (!origin || !win || (s > 0 && (!raw || !mapped))) return GSR_ERR_NULL_POINTER;
  if (!(step_size > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  if (win->nclip < 0 || win->nclip > GSR_MAX_CLIP || (flags & GSR_FLAG_CHW)) return GSR_ERR_BAD_ARGUMENT;
  return gsr_forward_impl(mapped, mapped + 3 * (size_t)s, mapped + 5 * (size_t)s, origin, s, h, w, 3, 0, 0, dmax,
                          ksigma, flags, workspace, workspace_bytes, stream, 0, 0, win, nullptr, 0, raw, step_size);
}

// Fused front end for a padded batch: raw (batch*s_per,9) -> imgs (batch,hmax,wmax,3); every sample its own
// size hw_host[b] and step size step_host[b] (HOST arrays).
extern "C" int gsr_frontend_forward_batch_padded(const float* raw, float* mapped, float* imgs, int batch,
                                                 int s_per, int hmax, int wmax, const int* hw_host,
                                                 const float* step_host, const float* dmax_host, float dmax,
                                                 float ksigma, void* workspace, size_t workspace_bytes,
                                                 void* stream) {
  int rc = gsr_padded_check(batch, s_per, hmax, wmax, hw_host);
  if (rc) return rc;
  if (batch > 0 && (!raw || !mapped || !imgs || !step_host)) return GSR_ERR_NULL_POINTER;
  for (int b = 0; b < batch; ++b)
    if (!(step_host[b] > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t S = (size_t)batch * s_per;
  float *sig = mapped, *crd = mapped + 3 * S, *col = mapped + 5 * S;
  const int g = gsr_padded_group(batch, hmax);
  GsrBDesc desc[GSR_BDESC_MAX];
  for (int b0 = 0; b0 < batch; b0 += g) {
    const int nb = batch - b0 < g ? batch - b0 : g, sg = nb * s_per;
    const size_t go = (size_t)b0 * s_per;
    const GsrWorkspace ws = gsr_carve(workspace, sg, nb * hmax, wmax);
    rc = gsr_check_ws(workspace, workspace_bytes, ws.bytes);
    if (rc) return rc;
    gsr_padded_fill(desc, nb, hmax, wmax, hw_host + 2 * b0, dmax_host ? dmax_host + b0 : nullptr, dmax, step_host + b0);
    rc = gsr_forward_impl(sig + 3 * go, crd + 2 * go, col + 3 * go, imgs + (size_t)b0 * hmax * wmax * 3, sg, nb * hmax,
                          wmax, 3, 0, 0, dmax, ksigma, GSR_FLAG_OVERWRITE, workspace, workspace_bytes, stream, s_per, hmax,
                          nullptr, desc, nb, raw + 9 * go, 1.0f);
    if (rc) return rc;
  }
  return GSR_OK;
}

// grads (batch,hmax,wmax,3) -> grad_raw (batch*s_per,9), written.  Workspace: gsr_workspace_bytes_batch_padded
// + 32 bytes per Gaussian of the largest launch group (at most the whole batch), rounded up to 256.
extern "C" int gsr_frontend_backward_batch_padded(const float* raw, const float* mapped, const float* grads,
                                                  float* grad_raw, int batch, int s_per, int hmax, int wmax,
                                                  const int* hw_host, const float* step_host,
                                                  const float* dmax_host, float dmax, float ksigma,
                                                  void* workspace, size_t workspace_bytes, void* stream) {
  int rc = gsr_padded_check(batch, s_per, hmax, wmax, hw_host);
  if (rc) return rc;
  if (batch > 0 && (!raw || !mapped || !grads || !grad_raw || !step_host)) return GSR_ERR_NULL_POINTER;
  for (int b = 0; b < batch; ++b)
    if (!(step_host[b] > 0.0f)) return GSR_ERR_BAD_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t S = (size_t)batch * s_per;
  const float *sig = mapped, *crd = mapped + 3 * S, *col = mapped + 5 * S;
  const int g = gsr_padded_group(batch, hmax);
  const size_t need = gsr_workspace_bytes_batch_padded(batch, s_per, hmax, wmax);
  const size_t stage = gsr_align_up((size_t)g * s_per * 8 * sizeof(float), 256);
  if (!workspace || ((uintptr_t)workspace & 255u) || workspace_bytes < need + stage) return GSR_ERR_WORKSPACE;
  float* gm = (float*)((char*)workspace + need);
  GsrBDesc desc[GSR_BDESC_MAX];
  for (int b0 = 0; b0 < batch; b0 += g) {
    const int nb = batch - b0 < g ? batch - b0 : g, sg = nb * s_per;
    if (sg == 0) continue;
    const size_t go = (size_t)b0 * s_per;
    const GsrWorkspace ws = gsr_carve(workspace, sg, nb * hmax, wmax);
    gsr_padded_fill(desc, nb, hmax, wmax, hw_host + 2 * b0, dmax_host ? dmax_host + b0 : nullptr, dmax, step_host + b0);
    GSR_CUDA(cudaMemsetAsync(gm, 0, (size_t)sg * 8 * sizeof(float), st));
    rc = gsr_backward_impl(sig + 3 * go, crd + 2 * go, col + 3 * go, grads + (size_t)b0 * hmax * wmax * 3, gm,
                           gm + 3 * (size_t)sg, gm + 5 * (size_t)sg, sg, nb * hmax, wmax, 3, 0, 0, dmax, ksigma, 0,
                           workspace, need, stream, s_per, hmax, desc, nb);
    if (rc) return rc;
    gsr_unmap_kernel<<<(sg + 255) / 256, 256, 0, st>>>(raw + 9 * go, gm, gm + 3 * (size_t)sg, gm + 5 * (size_t)sg,
                                                       grad_raw + 9 * go, sg, 0, 0, 0.f, ws.bdesc, s_per);
    GSR_CUDA(cudaGetLastError());
  }
  return GSR_OK;
}

// ---- CPU test hooks (no GPU needed): run the shared host/device culling code on the host ----
extern "C" void gsr_host_setup(const float* sigmas, const float* coords, const float* colors,
                               int s, int h, int w, float dmax, float ksigma, int* out /* s x 9 */) {
  const float keff = gsr_effective_ksigma(ksigma);
  for (int i = 0; i < s; ++i) {
    GsrSetup st = gsr_setup(sigmas[3 * i], sigmas[3 * i + 1], sigmas[3 * i + 2], coords[2 * i],
                            coords[2 * i + 1], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2],
                            h, w, dmax, keff);
    int* o = out + 9 * (size_t)i;
    o[0] = st.live;
    o[1] = st.x0;
    o[2] = st.x1;
    o[3] = st.y0;
    o[4] = st.y1;
    o[5] = st.binds;
    o[6] = st.large;
    o[7] = st.bin_y * ((w + GSR_BIN - 1) / GSR_BIN) + st.bin_x;
    o[8] = st.ext_x > st.ext_y ? st.ext_x : st.ext_y;
  }
}

// Same for a row band: rows [row0, row0 + rows) of the h-row image; boxes are band-local.
extern "C" void gsr_host_setup_band(const float* sigmas, const float* coords, const float* colors,
                                    int s, int h, int w, int row0, int rows, float dmax, float ksigma,
                                    int* out /* s x 6: live, x0, x1, y0, y1, binds */) {
  const float keff = gsr_effective_ksigma(ksigma);
  for (int i = 0; i < s; ++i) {
    GsrSetup st = gsr_setup(sigmas[3 * i], sigmas[3 * i + 1], sigmas[3 * i + 2], coords[2 * i],
                            coords[2 * i + 1], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2],
                            rows, w, dmax, keff, nullptr, nullptr, h, row0);
    int* o = out + 6 * (size_t)i;
    o[0] = st.live;
    o[1] = st.x0;
    o[2] = st.x1;
    o[3] = st.y0;
    o[4] = st.y1;
    o[5] = st.binds;
  }
}

extern "C" void gsr_host_window_range(int n, float ctr, float dmax, int* lo, int* hi) {
  gsr_window_range(n, ctr, dmax, *lo, *hi);
}

// Region mask of Gaussian i for the tile at (tx0, ty0); 0 if the Gaussian is not live.
extern "C" unsigned gsr_host_region_mask(const float* sigmas, const float* coords,
                                         const float* colors, int i, int h, int w, float dmax,
                                         float ksigma, int tx0, int ty0) {
  const float keff = gsr_effective_ksigma(ksigma);
  GsrSetup st = gsr_setup(sigmas[3 * i], sigmas[3 * i + 1], sigmas[3 * i + 2], coords[2 * i],
                          coords[2 * i + 1], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2], h,
                          w, dmax, keff);
  if (!st.live) return 0;
  if (st.x1 < tx0 || st.x0 >= tx0 + GSR_TILE_W || st.y1 < ty0 || st.y0 >= ty0 + GSR_TILE_H) return 0;
  GsrRec r = gsr_make_rec(sigmas[3 * i], sigmas[3 * i + 1], sigmas[3 * i + 2], coords[2 * i],
                          coords[2 * i + 1], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
  return gsr_region_mask(r, st.x0, st.x1, st.y0, st.y1, tx0, ty0, h, w, gsr_ecut(keff));
}

// Bucket entries of Gaussian i as the set-up kernel's shared-memory path writes them: triples
// (region column, region row, 8-bit cell mask), at most cap of them; returns their number (0: not live).
extern "C" int gsr_host_entries(const float* sigmas, const float* coords, const float* colors, int i, int h,
                                int w, float dmax, float ksigma, int* out, int cap) {
  const float keff = gsr_effective_ksigma(ksigma);
  GsrSetup st = gsr_setup(sigmas[3 * i], sigmas[3 * i + 1], sigmas[3 * i + 2], coords[2 * i],
                          coords[2 * i + 1], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2], h,
                          w, dmax, keff);
  if (!st.live) return 0;
  const GsrRec r = gsr_make_rec(sigmas[3 * i], sigmas[3 * i + 1], sigmas[3 * i + 2], coords[2 * i],
                                coords[2 * i + 1], colors[3 * i], colors[3 * i + 1], colors[3 * i + 2]);
  if (!(gsr_finite(r.a) && gsr_finite(r.b) && gsr_finite(r.c))) return 0;
  const GsrEllipse e = gsr_ellipse(r, h, w);
  int n = 0;
  const int cbase = (st.x0 / GSR_RGW) * GSR_CELLS_X;
  const bool narrow = st.x1 / GSR_CELL - cbase < 32;
  for (int b = st.y0 / GSR_RGH; b <= st.y1 / GSR_RGH; ++b) {
    uint32_t rb[2] = {0u, 0u};
    int cl[2], ch[2], ca, cb;
    if (narrow) {
      if (!gsr_band_rowbits(e, gsr_ecut(keff), b, st.x0, st.x1, st.y0, st.y1, cbase, rb)) continue;
      const uint32_t any = rb[0] | rb[1];
      int lo = 0, hi = 31;
      while (!((any >> lo) & 1u)) ++lo;
      while (!((any >> hi) & 1u)) --hi;
      ca = (cbase + lo) / GSR_CELLS_X;
      cb = (cbase + hi) / GSR_CELLS_X;
    } else {
      if (!gsr_band_cells(e, gsr_ecut(keff), b, st.x0, st.x1, st.y0, st.y1, cl, ch)) continue;
      gsr_band_columns(cl, ch, ca, cb);
    }
    for (int c = ca; c <= cb; ++c) {
      if (n < cap) {
        out[3 * n + 0] = c;
        out[3 * n + 1] = b;
        out[3 * n + 2] = (int)(narrow ? gsr_rowbits_mask(rb, c, cbase) : gsr_cell_mask(cl, ch, c));
      }
      ++n;
    }
  }
  return n;
}

extern "C" void gsr_host_geometry(int* tile_w, int* tile_h, int* bin, int* region, int* large_px) {
  *tile_w = GSR_TILE_W;
  *tile_h = GSR_TILE_H;
  *bin = GSR_BIN;
  *region = GSR_REGION;
  *large_px = GSR_LARGE_PX;
}

// ---- fused crop + L1 loss + gradient (gsasr_model.py:212-234) ---------------------------------------------------
constexpr int GSR_LOSS_MAX_GRID = 148 * 8;
extern "C" size_t gsr_l1_crop_workspace_bytes(void) { return 256 + (size_t)GSR_LOSS_MAX_GRID * sizeof(double); }

extern "C" int gsr_l1_crop_loss(const float* sr, const long long* sr_strides, const float* gt,
                                const long long* gt_strides, float* grad, float* loss, int batch, int hmax,
                                int wmax, const int* hw_host, float weight, int accumulate, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!sr || !gt || !grad || !loss || !sr_strides || !gt_strides || !hw_host) return GSR_ERR_NULL_POINTER;
  if (batch < 0 || hmax < 1 || wmax < 1 || hmax > GSR_MAX_DIM || wmax > GSR_MAX_DIM) return GSR_ERR_BAD_SHAPE;
  int rc = gsr_check_ws(workspace, workspace_bytes, gsr_l1_crop_workspace_bytes());
  if (rc) return rc;
  for (int b = 0; b < batch; ++b)
    if (hw_host[2 * b] < 1 || hw_host[2 * b] > hmax || hw_host[2 * b + 1] < 1 || hw_host[2 * b + 1] > wmax)
      return GSR_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  GSR_CUDA(cudaMemsetAsync(workspace, 0, 256, st));  // the finished-CTA counter
  if (batch == 0 && !accumulate) GSR_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  for (int b0 = 0; b0 < batch; b0 += GSR_LOSS_MAX_BATCH) {
    const int nb = batch - b0 < GSR_LOSS_MAX_BATCH ? batch - b0 : GSR_LOSS_MAX_BATCH;
    GsrLossArgs a;
    a.sr = sr + (long long)b0 * sr_strides[0];
    a.gt = gt + (long long)b0 * gt_strides[0];
    a.grad = grad + (long long)b0 * sr_strides[0];
    a.sr_n = sr_strides[0], a.sr_c = sr_strides[1], a.sr_h = sr_strides[2], a.sr_w = sr_strides[3];
    a.gt_n = gt_strides[0], a.gt_c = gt_strides[1], a.gt_h = gt_strides[2], a.gt_w = gt_strides[3];
    a.batch = nb;
    a.hmax = hmax;
    a.wmax = wmax;
    a.weight = weight;
    a.counter = (unsigned int*)workspace;
    a.partial = (double*)((char*)workspace + 256);
    a.loss = loss;
    a.accumulate = accumulate || b0 > 0;
    for (int b = 0; b < nb; ++b) {
      a.hw[b][0] = (short)hw_host[2 * (b0 + b)];
      a.hw[b][1] = (short)hw_host[2 * (b0 + b) + 1];
    }
    const long long total = (long long)nb * hmax * wmax;
    const long long want = (total + GSR_LOSS_THREADS - 1) / GSR_LOSS_THREADS;
    const int grid = (int)(want < GSR_LOSS_MAX_GRID ? want : GSR_LOSS_MAX_GRID);
    gsr_l1_crop_kernel<<<grid, GSR_LOSS_THREADS, 0, st>>>(a);
  }
  GSR_CUDA(cudaGetLastError());
  return GSR_OK;
}
