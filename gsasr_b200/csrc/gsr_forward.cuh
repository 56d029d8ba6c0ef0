// gsr_forward.cuh -- forward raster kernels (sm_100a).
//
// Replace _gs_render_cuda (utils/gs_cuda_dmax/gs.cu:7-64; utils/gs_cuda/gs.cu:9-61).
// The reference scatters: one thread per Gaussian, three global atomics per (Gaussian,pixel).
// Here pixels are gathered: accumulators live in registers and every pixel is written once.
//
//   gsr_forward_region_kernel  (the fast path, second half of this file)  streams the per-region
//       buckets built by gsr_region_build2_kernel: a warp per 16x8-pixel region, a 2x2 pixel block per lane,
//       every 4x4-pixel cell (four lanes) evaluating only the entries whose cell mask names it, persistent
//       independent warps, two-deep cp.async prefetch.  See its own comment.
//   gsr_forward_bins_kernel    (the fallback, runs when a bucket overflowed)  finds a tile's candidates
//       at run time.  The image is cut into 32x16 tiles, one CTA per tile (persistent over tiles), 8 warps
//       each owning an 8x8 region (two horizontally adjacent pixels per lane).  Per tile, in rounds of
//       at most GSR_FWD_CAP candidates:
//         stage A  (warp-private, no CTA barrier inside)
//           pass 1  every warp takes a slice of the round's candidates -- Gaussians of the home bins
//                   within reach of the tile (contiguous runs of the sorted arrays, one run per bin
//                   row) plus the "large" list -- tests cull box vs tile and streams the 32 B records
//                   of the hits into its private shared-memory segment with cp.async;
//           pass 2  one lane per hit: ellipse-vs-region mask (gsr_region_mask); hits are appended to
//                   the per-(producer warp, region) lists (ranks from ballots, no atomics);
//         stage C  (one warp per region)  walks the lists left for its region: per Gaussian 2 LDS.128
//                  (broadcast), FADD + 3 FMUL, FADD2, 2 FFMA2, 2 MUFU.EX2, 3 FFMA2.  Gaussians whose
//                  dmax window binds take the predicated evaluation.
// Both are bound by the MUFU pipe (16 ex2/clk/SM): see DESIGN.md.
#pragma once
#include "gsr_prepass.cuh"

constexpr int GSR_FWD_THREADS = 32 * GSR_NRX * GSR_NRY;
constexpr int GSR_FWD_WARPS = GSR_FWD_THREADS / 32;
#ifndef GSR_CFG_MIN_CTAS
#define GSR_CFG_MIN_CTAS 3
#endif
#ifndef GSR_CFG_PER_LANE
#define GSR_CFG_PER_LANE 6
#endif
constexpr int GSR_FWD_CAP = 32 * GSR_CFG_PER_LANE * GSR_NRX * GSR_NRY;                        // candidate slots per round
constexpr int GSR_FWD_SEG = GSR_FWD_CAP / GSR_FWD_WARPS;   // slots per warp segment (128)
constexpr int GSR_FWD_PER_LANE = GSR_FWD_SEG / 32;         // candidates per lane per round (4)
constexpr int GSR_FWD_LCAP = 32;   // entries per (producer warp, region) list per round; the rest
                                   // (and window-binding Gaussians) go to the scan list
constexpr int GSR_FWD_MAXRUNS = 2 * ((GSR_LARGE_PX + GSR_BIN - 1) / GSR_BIN) + GSR_TILE_H / GSR_BIN + 2;
static_assert(GSR_FWD_WARPS == GSR_NRX * GSR_NRY, "one warp per region");
static_assert(GSR_FWD_MAXRUNS <= 32, "run table is built by one warp");

struct GsrFwdSmem {
  float4 rec[GSR_FWD_CAP * 2];   // 64 KB  records, addressed by slot
  uint2 box[GSR_FWD_CAP];        // 16 KB  packed cull boxes, addressed by slot
  uint32_t list[GSR_FWD_CAP];    //  8 KB  overflow scan list: region mask | binds<<16 | slot<<17
  uint16_t lfast[GSR_FWD_WARPS][GSR_FWD_WARPS][GSR_FWD_LCAP];  // [producer warp][region] slot lists
  uint8_t lcnt[GSR_FWD_WARPS][GSR_FWD_WARPS];                  // their lengths, rewritten every round
  int run_start[GSR_FWD_MAXRUNS];
  int run_prefix[GSR_FWD_MAXRUNS + 1];
  int nruns;
  int nlist[2];  // alternates between rounds
};

__device__ __forceinline__ float gsr_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed FP32x2 arithmetic (sm_100: FFMA2 / FADD2, one issue slot for two lanes of work; ptxas uses
// the scalar-broadcast operand form when both halves of a pair are the same register).
typedef unsigned long long gsr_f2;
__device__ __forceinline__ gsr_f2 gsr_pk(float lo, float hi) {
  gsr_f2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void gsr_upk(gsr_f2 v, float& lo, float& hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ gsr_f2 gsr_fma2(gsr_f2 a, gsr_f2 b, gsr_f2 c) {
  gsr_f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ gsr_f2 gsr_add2(gsr_f2 a, gsr_f2 b) {
  gsr_f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// One Gaussian against this lane's two pixels: 1 FADD + 3 FMUL (row terms), FADD2 (dx pair),
// 2 FFMA2 (exponent pair), 2 MUFU.EX2, 3 FFMA2 (colour accumulate).  px2 = {px0, px1}.
__device__ __forceinline__ void gsr_eval_pair(uint32_t addr, gsr_f2 px2, float py, bool in0, bool in1,
                                              gsr_f2& accr, gsr_f2& accg, gsr_f2& accb);

__device__ __forceinline__ uint32_t gsr_smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 gsr_lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t gsr_lds_u16(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void gsr_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void gsr_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void gsr_eval_pair(uint32_t addr, gsr_f2 px2, float py, bool in0, bool in1,
                                              gsr_f2& accr, gsr_f2& accg, gsr_f2& accb) {
  const float4 a0 = gsr_lds128(addr);       // x, y, a, b
  const float4 a1 = gsr_lds128(addr + 16);  // c, r, g, bl
  const float dy = py - a0.y;
  const float t1 = a0.w * dy;
  const float t0 = a1.x * dy * dy;
  const gsr_f2 dx2 = gsr_add2(px2, gsr_pk(-a0.x, -a0.x));
  const gsr_f2 e2 = gsr_fma2(dx2, gsr_fma2(gsr_pk(a0.z, a0.z), dx2, gsr_pk(t1, t1)), gsr_pk(t0, t0));
  float e0, e1;
  gsr_upk(e2, e0, e1);
  const float v0 = in0 ? gsr_ex2(e0) : 0.f;
  const float v1 = in1 ? gsr_ex2(e1) : 0.f;
  const gsr_f2 v2 = gsr_pk(v0, v1);
  accr = gsr_fma2(v2, gsr_pk(a1.y, a1.y), accr);
  accg = gsr_fma2(v2, gsr_pk(a1.z, a1.z), accg);
  accb = gsr_fma2(v2, gsr_pk(a1.w, a1.w), accb);
}

#ifndef GSR_MAX_CLIP
#define GSR_MAX_CLIP 8  // as in include/gsraster.h
#endif

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
  const int* guard;  // run only if *guard == want (nullptr: always)
  int want;
  // region-bucket path
  const int* reg_count;
  int reg_cap, nrx, nry;
  const uint32_t* entries;
  const GsrRec* rec_in;
  const uint2* box_in;
  int* sched;    // {next unit, finished warps}: dynamic work distribution, left at zero by the kernel
  int hf, row0;  // row-band view (see gsr_setup)
  int bhs;       // uniform batch: rows per sample of the stacked image (0: single image)
  // destination addressing, in floats: pixel (hi, wi), channel ch lives at
  //   img + hi * row_stride + wi * pix_stride + ch * chan_stride
  // (HWC: 3w, 3, 1; CHW: w, 1, h*w; a window of a larger canvas: the canvas' strides, img = the address of
  // the render's pixel (0,0)).  nclip > 0: only pixels inside one of the clip rectangles are written.
  long long row_stride, pix_stride, chan_stride;
  int nclip;
  int clip[GSR_MAX_CLIP][4];  // x0, y0, x1, y1 (inclusive), render coordinates
};

__device__ __forceinline__ bool gsr_fwd_writable(const GsrFwdArgs& p, int hi, int wi) {
  if (hi >= p.h || wi >= p.w) return false;
  if (p.nclip == 0) return true;
  bool in = false;
#pragma unroll
  for (int k = 0; k < GSR_MAX_CLIP; ++k)
    in = in || (k < p.nclip && wi >= p.clip[k][0] && hi >= p.clip[k][1] && wi <= p.clip[k][2] && hi <= p.clip[k][3]);
  return in;
}
// inference_paper.py:136-138: clamp_(0, 1), * 255.0, round() (half to even), astype(uint8); NaN -> 0
__device__ __forceinline__ unsigned char gsr_to_u8(float v) {
  return (unsigned char)__float2uint_rn(__saturatef(v) * 255.0f);
}
__device__ __forceinline__ float* gsr_fwd_pixel(const GsrFwdArgs& p, int hi, int wi) {
  return p.img + (long long)hi * p.row_stride + (long long)wi * p.pix_stride;
}

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

// One write (or read-modify-write: the reference accumulates into rendered_img) per pixel.
__device__ __forceinline__ void gsr_fwd_writeout(const GsrFwdArgs& p, int hi, int wi0, float r0,
                                                 float g0, float b0, float r1, float g1, float b1,
                                                 bool force_over = false) {
  const bool over = force_over || (p.flags & 1u) != 0;
  const float v[2][3] = {{r0, g0, b0}, {r1, g1, b1}};
#pragma unroll
  for (int xx = 0; xx < 2; ++xx) {
    if (!gsr_fwd_writable(p, hi, wi0 + xx)) continue;
    if (p.flags & 4u) {  // uint8 (h,w,3) output
      unsigned char* o8 = reinterpret_cast<unsigned char*>(p.img) + ((size_t)hi * p.w + wi0 + xx) * 3;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) o8[(p.flags & 8u) ? 2 - ch : ch] = gsr_to_u8(v[xx][ch]);
      continue;
    }
    float* o = gsr_fwd_pixel(p, hi, wi0 + xx);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) o[ch * p.chan_stride] = over ? v[xx][ch] : o[ch * p.chan_stride] + v[xx][ch];
  }
}

__device__ __forceinline__ void gsr_forward_bins_body(const GsrFwdArgs& p, GsrFwdSmem& sm) {
  constexpr int NR = GSR_NRX;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  // persistent over tiles: the grid is capped so that a guarded no-op launch stays cheap
  const int ntx_ = (p.w + GSR_TILE_W - 1) / GSR_TILE_W, nty_ = (p.h + GSR_TILE_H - 1) / GSR_TILE_H;
  for (int tile = blockIdx.x; tile < ntx_ * nty_; tile += gridDim.x) {
  const int tx0 = (tile % ntx_) * GSR_TILE_W, ty0 = (tile / ntx_) * GSR_TILE_H;
  __syncthreads();  // shared memory of the previous tile is free

  // this thread's two pixels
  const int wi0 = tx0 + (warp % NR) * GSR_REGION + (lane & 3) * 2;
  const int hi = ty0 + (warp / NR) * GSR_REGION + (lane >> 2);
  const float px0 = __ldg(p.px_tab + min(wi0, p.w - 1));
  const float px1 = __ldg(p.px_tab + min(wi0 + 1, p.w - 1));
  const float py = __ldg(p.py_tab + min(hi, p.h - 1));
  gsr_f2 accr = gsr_pk(0.f, 0.f), accg = gsr_pk(0.f, 0.f), accb = gsr_pk(0.f, 0.f);
  const gsr_f2 px2 = gsr_pk(px0, px1);

  if (warp == 0)
    gsr_build_runs(p.bin_off, p.stats, p.nbx, p.nby, p.nb, tx0, tx0 + GSR_TILE_W - 1, ty0,
                   ty0 + GSR_TILE_H - 1, lane, sm.run_start, sm.run_prefix, &sm.nruns);
  if (tid == 0) sm.nlist[0] = sm.nlist[1] = 0;
  __syncthreads();
  const int nruns = sm.nruns;
  const int total = sm.run_prefix[nruns];
  // balanced rounds: each at most GSR_FWD_CAP candidates, split evenly over the warps
  const int nrounds = (total + GSR_FWD_CAP - 1) / GSR_FWD_CAP;
  const int per_warp = nrounds ? ((total + nrounds - 1) / nrounds + GSR_FWD_WARPS - 1) / GSR_FWD_WARPS : 0;
  const int per_round = per_warp * GSR_FWD_WARPS;

  const uint32_t rec_s = gsr_smem_addr(sm.rec);
  const int seg0 = warp * GSR_FWD_SEG;

  for (int round = 0; round < nrounds; ++round) {
    // ---------------- stage A, pass 1: cull box vs tile, stream hits into the segment ----------
    // 32-candidate groups are dealt round-robin to the warps, so that the hits (and with them the
    // expensive pass 2) spread evenly however the runs are laid out.
    const int rbase = round * per_round;
    const int rend = min(rbase + per_round, total);
    uint2 pb[GSR_FWD_PER_LANE];
    int gidx[GSR_FWD_PER_LANE];
    int r = 0;  // run of the current candidate: candidates grow with k, so the search resumes
#pragma unroll
    for (int k = 0; k < GSR_FWD_PER_LANE; ++k) {
      const int cnd = rbase + (k * GSR_FWD_WARPS + warp) * 32 + lane;
      gidx[k] = -1;
      pb[k] = make_uint2(0, 0);
      if (cnd < rend) {
        while (cnd >= sm.run_prefix[r + 1]) ++r;
        gidx[k] = sm.run_start[r] + (cnd - sm.run_prefix[r]);
        pb[k] = __ldg(p.box + gidx[k]);
      }
    }
    int n1 = 0;
#pragma unroll
    for (int k = 0; k < GSR_FWD_PER_LANE; ++k) {
      int bx0, bx1, by0, by1;
      bool binds;
      gsr_box_unpack(pb[k], bx0, bx1, by0, by1, binds);
      const bool hit = gidx[k] >= 0 && bx1 >= tx0 && bx0 < tx0 + GSR_TILE_W && by1 >= ty0 &&
                       by0 < ty0 + GSR_TILE_H;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int slot = seg0 + n1 + __popc(bal & lt_mask);
        const float4* src = reinterpret_cast<const float4*>(p.rec + gidx[k]);
        gsr_cp_async16(rec_s + slot * 32, src);
        gsr_cp_async16(rec_s + slot * 32 + 16, src + 1);
        sm.box[slot] = pb[k];
      }
      n1 += __popc(bal);
    }
    gsr_cp_async_wait_all();
    __syncwarp();
    // ---------------- stage A, pass 2: region masks, append to the region lists ---------------
    int cnt[GSR_FWD_WARPS];
#pragma unroll
    for (int rg = 0; rg < GSR_FWD_WARPS; ++rg) cnt[rg] = 0;
    for (int j = 0; j < n1; j += 32) {
      const int slot = seg0 + j + lane;
      uint32_t m = 0;
      bool binds = false;
      if (j + lane < n1) {
        const float4 q0 = sm.rec[2 * slot], q1 = sm.rec[2 * slot + 1];
        GsrRec g;
        g.x = q0.x; g.y = q0.y; g.a = q0.z; g.b = q0.w;
        g.c = q1.x; g.r = q1.y; g.g = q1.z; g.bl = q1.w;
        int bx0, bx1, by0, by1;
        gsr_box_unpack(sm.box[slot], bx0, bx1, by0, by1, binds);
        m = gsr_region_mask(g, bx0, bx1, by0, by1, tx0, ty0, p.h, p.w, p.ecut, p.hf, p.row0, p.bhs);
      }
      // Append to this warp's private list of every region touched (no atomics: the ranks come
      // from ballots).  Window-binding Gaussians and entries that do not fit go to the CTA-wide
      // scan list instead.
      uint32_t ovf = binds ? m : 0u;
#pragma unroll
      for (int rg = 0; rg < GSR_FWD_WARPS; ++rg) {
        const bool bit = ((m >> rg) & 1u) && !binds;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        if (bit) {
          const int pos = cnt[rg] + __popc(bal & lt_mask);
          if (pos < GSR_FWD_LCAP) sm.lfast[warp][rg][pos] = (uint16_t)slot; else ovf |= 1u << rg;
        }
        cnt[rg] += __popc(bal);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ovf != 0);
      if (bal) {
        int base = 0;
        const int leader = __ffs(bal) - 1;
        if (lane == leader) base = atomicAdd(&sm.nlist[round & 1], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (ovf) sm.list[base + __popc(bal & lt_mask)] = ovf | (binds ? 0x10000u : 0u) | ((uint32_t)slot << 17);
      }
    }
#pragma unroll
    for (int rg = 0; rg < GSR_FWD_WARPS; ++rg)
      if (lane == rg) sm.lcnt[warp][rg] = (uint8_t)min(cnt[rg], GSR_FWD_LCAP);
    __syncthreads();

    // ---------------- stage C: every warp walks its region's lists -----------------------------
    auto eval = [&](uint32_t addr, const bool in0, const bool in1) {
      gsr_eval_pair(addr, px2, py, in0, in1, accr, accg, accb);
    };
    auto eval_slow = [&](int slot) {  // dmax window cuts this Gaussian: exact inclusion per pixel
      int bx0, bx1, by0, by1;
      bool binds;
      gsr_box_unpack(sm.box[slot], bx0, bx1, by0, by1, binds);
      const bool iny = hi >= by0 && hi <= by1;
      eval(rec_s + (slot << 5), iny && wi0 >= bx0 && wi0 <= bx1, iny && wi0 + 1 >= bx0 && wi0 + 1 <= bx1);
    };
    for (int pw = 0; pw < GSR_FWD_WARPS; ++pw) {  // the list each producer warp left for this region
      const int nf = sm.lcnt[pw][warp];
      const uint16_t* lf = sm.lfast[pw][warp];
#pragma unroll 2
      for (int i = 0; i < nf; ++i) eval(rec_s + ((uint32_t)lf[i] << 5), true, true);
    }
    const int nlist = sm.nlist[round & 1];  // overflow entries (normally none)
    if (tid == 0) sm.nlist[(round + 1) & 1] = 0;  // free since the end of the previous round
    for (int j = 0; j < nlist; j += 32) {
      const uint32_t e = (j + lane < nlist) ? sm.list[j + lane] : 0u;
      unsigned mine = __ballot_sync(0xffffffffu, (e >> warp) & 1u);
      while (mine) {
        const int src = __ffs(mine) - 1;
        mine &= mine - 1;
        const uint32_t es = __shfl_sync(0xffffffffu, e, src);
        if (es & 0x10000u) eval_slow((int)(es >> 17)); else eval(rec_s + ((es >> 17) << 5), true, true);
      }
    }
    __syncthreads();  // segments and list are reused by the next round
  }

  {
    float r0, r1, g0, g1, b0, b1;
    gsr_upk(accr, r0, r1);
    gsr_upk(accg, g0, g1);
    gsr_upk(accb, b0, b1);
    gsr_fwd_writeout(p, hi, wi0, r0, g0, b0, r1, g1, b1);
  }
  }  // tile loop
}

__global__ void __launch_bounds__(GSR_FWD_THREADS, GSR_CFG_MIN_CTAS) gsr_forward_bins_kernel(GsrFwdArgs p) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  extern __shared__ __align__(16) unsigned char gsr_smem_raw[];
  gsr_forward_bins_body(p, *reinterpret_cast<GsrFwdSmem*>(gsr_smem_raw));
}

// The forward's fallback as ONE launch: when a region bucket overflowed, the home-bin set-up (gsr_bin_one, scan,
// gsr_scatter_one -- the kernels K1..K3 as phases separated by grid barriers) and then the raster over the home
// bins; when none did -- the normal case -- every CTA returns at once, and the call has paid for one idle launch
// instead of four (3.4 us instead of 13 at HL).  The grid is one resident wave (sized by occupancy).
template <bool RAGGED>
__global__ void __launch_bounds__(GSR_FWD_THREADS, GSR_CFG_MIN_CTAS)
gsr_forward_fallback_kernel(GsrFwdArgs p, const float* __restrict__ sigmas, const float* __restrict__ coords,
                            const float* __restrict__ colors, int s, float dmax, float ksigma, GsrWorkspace ws) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  extern __shared__ __align__(16) unsigned char gsr_smem_raw[];
  int* const barrier = ws.stats + GSR_STAT_BARRIER;
  int phase = 0;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
  if (gtid == 0) ws.stats[GSR_STAT_KSIGMA] = __float_as_int(ksigma);
  for (int i0 = blockIdx.x * blockDim.x; i0 < s; i0 += gsize)  // (whole warps enter gsr_bin_one: it votes)
    if (i0 + (int)threadIdx.x < s) gsr_bin_one<RAGGED>(sigmas, coords, colors, i0 + threadIdx.x, p.h, p.w, dmax, ksigma, ws);
  gsr_grid_barrier(barrier, (++phase) * gridDim.x);
  gsr_grid_scan(ws.bin_count, ws.bin_off, ws.nb + 1, ws.scan_state, barrier, phase);
  gsr_grid_barrier(barrier, (++phase) * gridDim.x);
  for (int i = gtid; i < s; i += gsize) gsr_scatter_one(sigmas, coords, colors, i, ws);
  gsr_grid_barrier(barrier, (++phase) * gridDim.x);
  gsr_forward_bins_body(p, *reinterpret_cast<GsrFwdSmem*>(gsr_smem_raw));
}

// ---- region-bucket forward kernel (the fast path) ------------------------------------------------
// One warp per 16x8-pixel region, a 2x2 pixel block per lane (accumulators in registers: six FP32x2 pairs);
// the four lanes of a 4x4-pixel CELL move together.  A bucket entry (written by gsr_region_build2_kernel)
// carries the Gaussian's index and an 8-bit mask of the cells its k-sigma ellipse reaches, and every cell
// evaluates ONLY the entries that name it: of the 128 pixels of a region a Gaussian of the x4 head reaches
// ~65 (half the cells), so half of the exponentials a whole-region evaluation would issue never are.
//
// Per chunk of up to 64 entries the warp (i) stages the 32-byte records with cp.async (global -> shared, no
// register staging), (ii) turns the masks into eight per-cell slot lists in shared memory -- 16 ballots, the
// ranks are popcounts, no atomics -- padded with the slot of a null record to a multiple of four, and
// (iii) runs ONE loop of T = max over the cells of the list length: each lane reads its own cell's next
// four slots (one LDS.32) and evaluates those records, so lanes of different cells work on different
// records in the same instruction (LDS.128 with one address per cell: conflict-free or two-way).
// Per record and lane: 2 LDS.128, 2 FADD2 (dx, dy pairs; the pixel coordinates are held negated), 3 FMUL2
// (row terms), 4 FFMA2 (exponents), 4 MUFU.EX2, 6 FFMA2 (colours) + 2 integer instructions for the address.
// A chunk takes every nch-th entry of the bucket (nch = chunks of the bucket), so that each is a stratified
// sample of the bucket's Gaussians -- which arrive sorted by position for a fea2gs field -- and the eight
// lists of a chunk come out nearly equally long (the loop runs as long as the longest).
//
// Warps are persistent and independent (no CTA barrier, nothing shared between warps) and claim regions
// from a global counter, two at a time and well ahead of need (one at a time near the end of the image).  The
// chunk stream is flat over a warp's regions, two-deep: while chunk k is evaluated the records of chunk k+1
// are in flight as cp.async copies and the entries of chunk k+2 as loads whose values are not touched
// before the next iteration.  Window-binding Gaussians (entry bit 31) also bring their cull box and are
// evaluated with the exact per-pixel inclusion test.
constexpr int GSR_FR_WARPS = 4;
constexpr int GSR_FR_THREADS = 32 * GSR_FR_WARPS;
constexpr int GSR_FR_CHUNK = 64;                          // entries per stage (two per lane)
constexpr int GSR_FR_SLOTS = GSR_FR_CHUNK + 1;            // + the null record (slot GSR_FR_CHUNK)
constexpr int GSR_FR_HI = GSR_FR_SLOTS * 16;              // byte offset of the second float4 of a record
constexpr int GSR_FR_STAGE_BYTES = 2 * GSR_FR_HI;
#ifndef GSR_CFG_FR_LW
#define GSR_CFG_FR_LW 2
#endif
constexpr int GSR_FR_LW = GSR_CFG_FR_LW;                  // bytes per list entry: a 16- or 32-bit shared-memory address
constexpr int GSR_FR_LIST = GSR_FR_LW * (GSR_FR_CHUNK + 4);  // bytes per cell list: 34 (68) words, so the eight cells'
                                                          // LDS.64 (LDS.128) fall into different banks
constexpr int GSR_FR_LIST_STAGE = 8 * GSR_FR_LIST;        // 1088 = 68 x 16 (2176)
static_assert(GSR_FR_LW == 2 || GSR_FR_LW == 4, "list entries are 16- or 32-bit addresses");
#ifndef GSR_CFG_FR_TAIL
#define GSR_CFG_FR_TAIL 1   // trip counts are not rounded up to the unroll factor: a chunk's last 1-3 list positions run singly
#endif
#ifndef GSR_CFG_FR_MIN_CTAS
#define GSR_CFG_FR_MIN_CTAS 5
#endif
#ifndef GSR_CFG_FR_UNROLL8
#define GSR_CFG_FR_UNROLL8 1   // eight list positions per iteration while they last, then four, then singly (-1.7 %)
#endif
static_assert(GSR_RGW == 16 && GSR_RGH == 8 && GSR_CELL == 4, "a warp of 2x2 blocks covers a 16x8 region, four lanes a cell");

struct GsrFwdRegionSmem {
  float4 rec[GSR_FR_WARPS][2][2 * GSR_FR_SLOTS];   // per warp, 2 stages x { first float4 x 65, second float4 x 65 }
  uint2 box[GSR_FR_WARPS][2][GSR_FR_CHUNK];
  uint32_t list[GSR_FR_WARPS][2][GSR_FR_LIST_STAGE / 4];
  float4 tile[GSR_FR_WARPS][GSR_RGW * GSR_RGH * 3 / 4];  // write-out staging (gsr_fr_write_unit)
};
static_assert(GSR_FR_LW == 4 || sizeof(GsrFwdRegionSmem) + 1024 < 65536, "cell lists hold 16-bit shared-memory addresses");

__device__ __forceinline__ gsr_f2 gsr_mul2(gsr_f2 a, gsr_f2 b) {
  gsr_f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void gsr_cp_async16ca(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void gsr_cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t gsr_lds32u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void gsr_sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%1], %0;" ::"h"((unsigned short)v), "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t gsr_lds16u(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 gsr_lds64u(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void gsr_sts128u(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void gsr_sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// Four consecutive slots (positions t .. t+3, t a multiple of 4) of a cell list -> four record addresses.
__device__ __forceinline__ void gsr_fr_load4(uint32_t lb, int t, uint32_t a[4]) {
  if (GSR_FR_LW == 2) {
    const uint2 s4 = gsr_lds64u(lb + 2 * t);
    a[0] = s4.x & 0xffffu, a[1] = s4.x >> 16, a[2] = s4.y & 0xffffu, a[3] = s4.y >> 16;
  } else {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(lb + 4 * t));
  }
}

// One Gaussian against this lane's 2x2 pixel block.  nx2/ny2 hold the NEGATED pixel coordinates, so
// d = x + (-px) = -(px - x): the reference's fp32 subtraction with the sign flipped, which the quadratic
// form does not see (every term is a product of two d's) -- bit-identical exponents, no negation.
// acc = {R,G,B} x {row 0, row 1}.
template <bool MASKED>
__device__ __forceinline__ void gsr_eval_quad(uint32_t addr0, uint32_t addr1, gsr_f2 nx2, gsr_f2 ny2,
                                              bool m00, bool m01, bool m10, bool m11, gsr_f2& r0,
                                              gsr_f2& g0, gsr_f2& b0, gsr_f2& r1, gsr_f2& g1, gsr_f2& b1) {
  const float4 a0 = gsr_lds128(addr0);  // x, y, a, b
  const float4 a1 = gsr_lds128(addr1);  // c, r, g, bl
  const gsr_f2 dx2 = gsr_add2(nx2, gsr_pk(a0.x, a0.x));
  const gsr_f2 dy2 = gsr_add2(ny2, gsr_pk(a0.y, a0.y));
  const gsr_f2 t1 = gsr_mul2(gsr_pk(a0.w, a0.w), dy2);
  const gsr_f2 t0 = gsr_mul2(gsr_mul2(gsr_pk(a1.x, a1.x), dy2), dy2);
  float t1a, t1b, t0a, t0b;
  gsr_upk(t1, t1a, t1b);
  gsr_upk(t0, t0a, t0b);
  const gsr_f2 a2 = gsr_pk(a0.z, a0.z);
  const gsr_f2 ea = gsr_fma2(dx2, gsr_fma2(a2, dx2, gsr_pk(t1a, t1a)), gsr_pk(t0a, t0a));
  const gsr_f2 eb = gsr_fma2(dx2, gsr_fma2(a2, dx2, gsr_pk(t1b, t1b)), gsr_pk(t0b, t0b));
  float e00, e01, e10, e11;
  gsr_upk(ea, e00, e01);
  gsr_upk(eb, e10, e11);
  float v00 = gsr_ex2(e00), v01 = gsr_ex2(e01), v10 = gsr_ex2(e10), v11 = gsr_ex2(e11);
  if (MASKED) {
    v00 = m00 ? v00 : 0.f;
    v01 = m01 ? v01 : 0.f;
    v10 = m10 ? v10 : 0.f;
    v11 = m11 ? v11 : 0.f;
  }
  const gsr_f2 va = gsr_pk(v00, v01), vb = gsr_pk(v10, v11);
  const gsr_f2 cr = gsr_pk(a1.y, a1.y), cg = gsr_pk(a1.z, a1.z), cb = gsr_pk(a1.w, a1.w);
  r0 = gsr_fma2(va, cr, r0);
  g0 = gsr_fma2(va, cg, g0);
  b0 = gsr_fma2(va, cb, b0);
  r1 = gsr_fma2(vb, cr, r1);
  g1 = gsr_fma2(vb, cg, g1);
  b1 = gsr_fma2(vb, cb, b1);
}

// Per-cell lists of 16-bit shared-memory addresses (the CTA's shared window is < 64 KB) for one staged chunk: every
// slot starts as the null record's; the rank of an entry in the list of cell q = number of earlier entries that
// name q, in the order (lane 0: first, second entry), (lane 1: ...).  The eight ranks of an entry come from ONE warp
// scan: the masks are spread to a byte per cell (two registers), the bytes are prefix-summed across the lanes with
// shuffles (counts stay below 256: at most 64 entries) -- no popcounts (they share the MUFU pipe).
// lw: the stage's lists; rb: the stage's records; (v1a, e1a), (v1b, e1b): this lane's two entries.  Returns the
// length of the list of this lane's cell.  Warp-collective.
// RANKS: also hand back {masks (first | second << 8), ranks of the first entry (cells 0-3, 4-7), of the second}.
template <bool RANKS = false>
__device__ __forceinline__ int gsr_fr_build_lists(uint32_t lw, uint32_t rb, int lane, int cell, bool v1a, uint32_t e1a,
                                                  bool v1b, uint32_t e1b, uint32_t* rk = nullptr) {
  constexpr int CH = GSR_FR_CHUNK;
  const unsigned full = 0xffffffffu;
  const uint32_t null2 = (rb + CH * 16u) * (GSR_FR_LW == 2 ? 0x00010001u : 1u);
#pragma unroll
  for (int o = 0; o < GSR_FR_LIST_STAGE; o += 512)
    if (o + 512 <= GSR_FR_LIST_STAGE || lane < (GSR_FR_LIST_STAGE - o) / 16) gsr_sts128u(lw + o + lane * 16, null2);
  const uint32_t ma = v1a ? (e1a >> GSR_ENT_MASK_SHIFT) & 0xffu : 0u, mb = v1b ? (e1b >> GSR_ENT_MASK_SHIFT) & 0xffu : 0u;
  const uint32_t a_lo = ((ma & 15u) * 0x00204081u) & 0x01010101u, a_hi = ((ma >> 4) * 0x00204081u) & 0x01010101u;
  const uint32_t b_lo = ((mb & 15u) * 0x00204081u) & 0x01010101u, b_hi = ((mb >> 4) * 0x00204081u) & 0x01010101u;
  uint32_t s_lo = a_lo + b_lo, s_hi = a_hi + b_hi;  // inclusive prefix sums, a byte per cell
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t0 = __shfl_up_sync(full, s_lo, d), t1 = __shfl_up_sync(full, s_hi, d);
    if (lane >= d) {
      s_lo += t0;
      s_hi += t1;
    }
  }
  const uint32_t t_lo = __shfl_sync(full, s_lo, 31), t_hi = __shfl_sync(full, s_hi, 31);
  const uint32_t ra_lo = s_lo - a_lo - b_lo, ra_hi = s_hi - a_hi - b_hi;  // exclusive: rank of the first entry
  const uint32_t rb_lo = ra_lo + a_lo, rb_hi = ra_hi + a_hi;              // the second entry follows the first
  __syncwarp();  // the null fill is complete before the slots are written
  const uint32_t adr_a = rb + lane * 16u, adr_b = adr_a + 32 * 16u;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t ra = ((q < 4 ? ra_lo : ra_hi) >> (8 * (q & 3))) & 0xffu;
    const uint32_t rbq = ((q < 4 ? rb_lo : rb_hi) >> (8 * (q & 3))) & 0xffu;
    if (GSR_FR_LW == 2) {
      if ((ma >> q) & 1u) gsr_sts16(lw + q * GSR_FR_LIST + 2 * ra, adr_a);
      if ((mb >> q) & 1u) gsr_sts16(lw + q * GSR_FR_LIST + 2 * rbq, adr_b);
    } else {
      if ((ma >> q) & 1u) gsr_sts32(lw + q * GSR_FR_LIST + 4 * ra, adr_a);
      if ((mb >> q) & 1u) gsr_sts32(lw + q * GSR_FR_LIST + 4 * rbq, adr_b);
    }
  }
  if (RANKS) {
    rk[0] = ma | (mb << 8);
    rk[1] = ra_lo;
    rk[2] = ra_hi;
    rk[3] = rb_lo;
    rk[4] = rb_hi;
  }
  const uint32_t tot = cell < 4 ? t_lo : t_hi;
  const int mine = (int)((tot >> (8 * (cell & 3))) & 0xffu);
  return mine;
}

// ---- write-out of a finished 16x8 region (both raster kernels) ----------------------------------------------
// v[yy][xx][ch]: this lane's 2x2 block at region offset (bx, by); (ux, uy): the region.  `tile` is 1536 bytes of
// warp-private shared memory.  Plain stores when the image is overwritten, fire-and-forget reductions (RED) when
// the call accumulates into the caller's image (the reference's contract): no read, no latency.
// Writing the uint8 image (w % 16 == 0), or overwriting an (h,w,3) fp32 image (w % 4 == 0) under
// GSR_FLAG_ROW_STORES, a region inside the image goes through the tile and leaves as 128-bit stores of whole
// region rows (48 / 192 contiguous bytes): what matters when the image lives on ANOTHER GPU
// (render_image_bands_peer) -- NVLink packets of 16 bytes per lane instead of 4; on a local fp32 image the
// detour costs 1 % (HL 279.6 vs 276.7 us), hence the flag.
#ifndef GSR_CFG_FR_STAGE_OUT
#define GSR_CFG_FR_STAGE_OUT 1
#endif
constexpr int GSR_FR_TILE_BYTES = GSR_RGW * GSR_RGH * 3 * 4;
template <bool WINDOW>
__device__ __forceinline__ void gsr_fr_write_unit(const GsrFwdArgs& p, uint32_t tile, int lane, int ux, int uy, int bx,
                                                  int by, const float (&v)[2][2][3]) {
  const bool over = (p.flags & 1u) != 0, chw = (p.flags & 2u) != 0, u8 = (p.flags & 4u) != 0, bgr = (p.flags & 8u) != 0;
  const int wi0 = ux * GSR_RGW + bx, hi0 = uy * GSR_RGH + by;
  const size_t plane = (size_t)p.h * p.w;
  if (GSR_CFG_FR_STAGE_OUT && !WINDOW && (ux + 1) * GSR_RGW <= p.w && (uy + 1) * GSR_RGH <= p.h &&
      (reinterpret_cast<uintptr_t>(p.img) & 15u) == 0) {
    if (u8 && (p.w & 15) == 0) {
      // bytes of the tile: [row][pixel][channel], 48 per row; a lane's row piece is 6 bytes at an even offset
#pragma unroll
      for (int yy = 0; yy < 2; ++yy) {
        uint32_t b[6];
#pragma unroll
        for (int xx = 0; xx < 2; ++xx)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) b[xx * 3 + (bgr ? 2 - ch : ch)] = gsr_to_u8(v[yy][xx][ch]);
        const uint32_t a = tile + ((by + yy) * GSR_RGW + bx) * 3;
        gsr_sts16(a, b[0] | (b[1] << 8));
        gsr_sts16(a + 2, b[2] | (b[3] << 8));
        gsr_sts16(a + 4, b[4] | (b[5] << 8));
      }
      __syncwarp();
      if (lane < GSR_RGH * 3) {  // 24 x 16 bytes
        const int row = lane / 3, part = lane - row * 3;
        uint4 q;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(tile + lane * 16) : "memory");
        unsigned char* o8 = reinterpret_cast<unsigned char*>(p.img) + ((size_t)(uy * GSR_RGH + row) * p.w + ux * GSR_RGW) * 3;
        *reinterpret_cast<uint4*>(o8 + part * 16) = q;
      }
      __syncwarp();
      return;
    }
    if (over && !chw && !u8 && (p.flags & 16u) != 0 && (p.w & 3) == 0) {  // GSR_FLAG_ROW_STORES
      // floats of the tile: [row][pixel][channel], 48 per row; a lane's row piece is 6 floats, 8-byte aligned
#pragma unroll
      for (int yy = 0; yy < 2; ++yy) {
        const uint32_t a = tile + ((by + yy) * GSR_RGW + bx) * 12;
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v[yy][0][0]), "f"(v[yy][0][1]) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a + 8), "f"(v[yy][0][2]), "f"(v[yy][1][0]) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a + 16), "f"(v[yy][1][1]), "f"(v[yy][1][2]) : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 3; ++j) {  // 96 x 16 bytes, 12 per region row
        const int i = lane + 32 * j, row = i / 12, part = i - row * 12;
        float4 q;  // (volatile + memory clobber: ordered after the staging stores, unlike gsr_lds128)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(tile + i * 16) : "memory");
        float* o = p.img + ((size_t)(uy * GSR_RGH + row) * p.w + ux * GSR_RGW) * 3;
        *reinterpret_cast<float4*>(o + part * 4) = q;
      }
      __syncwarp();
      return;
    }
  }
#pragma unroll
  for (int yy = 0; yy < 2; ++yy) {
    if (!WINDOW && !over && !chw && !u8 && (p.w & 1) == 0) {
      // accumulate into an (h,w,3) image of even width: the lane's two pixels of this row are six contiguous
      // floats at an 8-byte aligned address (wi0 is even) -- three vector reductions instead of six scalar ones
      if (hi0 + yy < p.h && wi0 < p.w) {  // (w even, wi0 even: both pixels of the pair are inside)
        float* o = p.img + ((size_t)(hi0 + yy) * p.w + wi0) * 3;
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o), "f"(v[yy][0][0]), "f"(v[yy][0][1]) : "memory");
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o + 2), "f"(v[yy][0][2]), "f"(v[yy][1][0]) : "memory");
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o + 4), "f"(v[yy][1][1]), "f"(v[yy][1][2]) : "memory");
      }
      continue;
    }
#pragma unroll
    for (int xx = 0; xx < 2; ++xx) {
      const int hi = hi0 + yy, wi = wi0 + xx;
      if (WINDOW) {
        if (gsr_fwd_writable(p, hi, wi)) {
          float* o = gsr_fwd_pixel(p, hi, wi);
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            if (over) o[ch * p.chan_stride] = v[yy][xx][ch];
            else atomicAdd(o + ch * p.chan_stride, v[yy][xx][ch]);
          }
        }
      } else if (hi < p.h && wi < p.w) {
        const size_t pix = (size_t)hi * p.w + wi;
        if (u8) {  // fused post-processing: clamp, x255, round-half-even, uint8 (h,w,3)
          unsigned char* o8 = reinterpret_cast<unsigned char*>(p.img) + pix * 3;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) o8[bgr ? 2 - ch : ch] = gsr_to_u8(v[yy][xx][ch]);
        } else {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            float* o = chw ? p.img + ch * plane + pix : p.img + pix * 3 + ch;
            if (over) *o = v[yy][xx][ch];
            else atomicAdd(o, v[yy][xx][ch]);
          }
        }
      }
    }
  }
}

// WINDOW = false: the plain (h,w,3) / (3,h,w) image, addressed with compile-time-simple arithmetic;
// WINDOW = true: the general strided destination with clip rectangles (gsr_forward_window).
template <bool WINDOW>
__global__ void __launch_bounds__(GSR_FR_THREADS, GSR_CFG_FR_MIN_CTAS) gsr_forward_region_kernel(GsrFwdArgs p) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  __shared__ GsrFwdRegionSmem sm;
  constexpr int CH = GSR_FR_CHUNK;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cell = lane >> 2;
  const unsigned lt = (1u << lane) - 1u;
  const int nunits = p.nrx * p.nry;
  const uint32_t rec_s = gsr_smem_addr(&sm.rec[warp][0][0]);
  const uint32_t list_w = gsr_smem_addr(&sm.list[warp][0][0]);   // the warp's lists, stage 0
  const uint32_t list_c = list_w + cell * GSR_FR_LIST;           // this lane's cell
  const uint32_t box_s = gsr_smem_addr(&sm.box[warp][0][0]);
  const uint2* box_w = &sm.box[warp][0][0];
  const uint32_t tile_s = gsr_smem_addr(&sm.tile[warp][0]);
  const int total_warps = gridDim.x * GSR_FR_WARPS;

  // the null record of both stages (slot CH): zero conic and colour, adds exactly 0
  if (lane < 4) sm.rec[warp][lane >> 1][(lane & 1) * GSR_FR_SLOTS + CH] = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- work distribution: units are claimed from a global counter, two at a time while plenty are left (one hot
  // address serves every warp of the GPU: one atomic per unit made it the kernel's largest single stall) and one
  // at a time near the end; a claim is requested a whole claim's worth of units before it is needed.
  // claims of two while at least eight units per warp remain ahead: the tail (and small images) balance unit by unit
  auto claim_size = [&](int progress) { return progress + 8 * total_warps < nunits ? 2 : 1; };
  int uA, uB, uC;
  int qn, qe;              // [qn, qe): units claimed and not yet handed out
  int pend, pend_n;        // lane 0's counter value of the claim in flight, and its size
  {
    pend_n = claim_size(3 * total_warps);
    int base = 0;
    if (lane == 0) base = atomicAdd(p.sched, 3 + pend_n);
    base = __shfl_sync(full, base, 0);
    uA = base, uB = base + 1, uC = base + 2;
    qn = qe = base + 3;    // nothing queued: the first refill takes the claim made here
    pend = base + 3;
  }
  auto take_unit = [&]() {  // next unit of this warp; refills from the claim in flight and requests another
    if (qn == qe) {
      qn = __shfl_sync(full, pend, 0);
      qe = qn + pend_n;
      pend_n = claim_size(qn);
      // lane 0 only, predicated inside the asm (no divergent region): the result is not waited for before the
      // claim is needed, several units from now
      asm volatile("{\n\t.reg .pred pl0;\n\tsetp.eq.s32 pl0, %2, 0;\n\t@pl0 atom.global.add.u32 %0, [%1], %3;\n\t}"
                   : "+r"(pend) : "l"(p.sched), "r"(lane), "r"(pend_n) : "memory");
    }
    return qn++;
  };
  auto finish = [&]() {  // the last warp to leave resets the counters for the next launch
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(p.sched + 1, 1) == total_warps - 1) {
        p.sched[0] = 0;
        p.sched[1] = 0;
      }
    }
  };
  if (uA >= nunits) {
    finish();
    return;
  }

  auto count_of = [&](int u) { return u < nunits ? __ldg(p.reg_count + u) : 0; };  // clamp with reg_cap on use
  auto chunks_of = [&](int n) { return n > CH ? (n + CH - 1) / CH : 1; };

  // Entries lane and lane + 32 of chunk ci of unit u: positions ci + k * nch.  Loaded values are not touched
  // before they are consumed one chunk later; validity is decided from the indices alone.
  uint32_t e1a = 0, e1b = 0;
  bool v1a = false, v1b = false;
  auto request_entries = [&](int u, int ci, int n, int nch) {
    const int i0 = ci + lane * nch, i1 = i0 + 32 * nch;
    v1a = i0 < n;
    v1b = i1 < n;
    const uint32_t* src = p.entries + (size_t)(u < nunits ? u : 0) * p.reg_cap;
    e1a = v1a ? __ldg(src + i0) : 0u;
    e1b = v1b ? __ldg(src + i1) : 0u;
  };
  // Records of the requested entries -> stage `st` (cp.async), their cell lists -> list stage `st`.
  // Returns the trip count of the chunk (longest list, rounded up to 4) and the binds ballots.
  unsigned slow_a = 0, slow_b = 0;
  auto stage_chunk = [&](int st) -> int {
    const uint32_t rb = rec_s + st * GSR_FR_STAGE_BYTES;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t en = t ? e1b : e1a;
      const bool v = t ? v1b : v1a;
      if (v) {
        const uint32_t gi = en & GSR_ENT_INDEX;
        const int k = lane + 32 * t;
        const char* src = reinterpret_cast<const char*>(p.rec_in + gi);
        gsr_cp_async16ca(rb + k * 16, src);
        gsr_cp_async16ca(rb + GSR_FR_HI + k * 16, src + 16);
        if (en >> 31) gsr_cp_async8(box_s + (st * CH + k) * 8, p.box_in + gi);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    slow_a = __ballot_sync(full, v1a && (e1a >> 31));
    slow_b = __ballot_sync(full, v1b && (e1b >> 31));
    const int mine = gsr_fr_build_lists(list_w + st * GSR_FR_LIST_STAGE, rb, lane, cell, v1a, e1a, v1b, e1b);
#if GSR_CFG_FR_TAIL
    return __reduce_max_sync(full, mine);
#else
    return (__reduce_max_sync(full, mine) + 3) & ~3;
#endif
  };

  // Units in flight: A is evaluated, B and C are known far enough ahead for the two-deep prefetch to run
  // across unit boundaries; further units wait in the claimed range [qn, qe) and in the claim in flight.
  int nA = min(count_of(uA), p.reg_cap), nB = min(count_of(uB), p.reg_cap), nC = count_of(uC);
  int nchA = chunks_of(nA), nchB = chunks_of(nB);

  request_entries(uA, 0, nA, nchA);
  int trip = stage_chunk(0);
  unsigned slow_ac = slow_a, slow_bc = slow_b;
  if (1 < nchA) request_entries(uA, 1, nA, nchA); else request_entries(uB, 0, nB, nchB);

  int cur = 0, ci = 0;
  // pixel block of this lane in unit u: cell (cell & 3, cell >> 2), block (lane & 1, (lane >> 1) & 1) of the cell
  const int bx = (cell & 3) * GSR_CELL + (lane & 1) * 2, by = (cell >> 2) * GSR_CELL + ((lane >> 1) & 1) * 2;
  auto coords_of = [&](int u, gsr_f2& nx, gsr_f2& ny) {
    const int uy = (u < nunits ? u : 0) / p.nrx, ux = (u < nunits ? u : 0) - uy * p.nrx;
    const int wi = ux * GSR_RGW + bx, hi = uy * GSR_RGH + by;
    nx = gsr_pk(-__ldg(p.px_tab + min(wi, p.w - 1)), -__ldg(p.px_tab + min(wi + 1, p.w - 1)));
    ny = gsr_pk(-__ldg(p.py_tab + min(hi, p.h - 1)), -__ldg(p.py_tab + min(hi + 1, p.h - 1)));
  };
  gsr_f2 nx2, ny2, nx2B, ny2B;
  coords_of(uA, nx2, ny2);
  coords_of(uB, nx2B, ny2B);
  gsr_f2 r0 = gsr_pk(0.f, 0.f), g0 = r0, b0 = r0, r1 = r0, g1 = r0, b1 = r0;

  for (;;) {  // one chunk per iteration, flat over the warp's units
    const uint32_t rb = rec_s + cur * GSR_FR_STAGE_BYTES;
    const uint32_t lb = list_c + cur * GSR_FR_LIST_STAGE;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();  // this chunk's records and lists are visible; the other stage is free
    // records and lists of the next chunk (its entries were requested one chunk ago) ...
    const int trip_n = stage_chunk(cur ^ 1);
    const unsigned slow_an = slow_a, slow_bn = slow_b;
    // ... and the entries of the chunk after it: (A, ci + 2), (B, 0), (B, 1) or (C, 0)
    const bool last = ci + 1 >= nchA;
    if (!last) {
      if (ci + 2 < nchA) request_entries(uA, ci + 2, nA, nchA); else request_entries(uB, 0, nB, nchB);
    } else {
      if (1 < nchB) request_entries(uB, 1, nB, nchB);
      else { const int n = min(nC, p.reg_cap); request_entries(uC, 0, n, chunks_of(n)); }
    }

    if ((slow_ac | slow_bc) == 0) {
#if GSR_CFG_FR_TAIL
      int t = 0;
#if GSR_CFG_FR_UNROLL8
      for (; t + 8 <= trip; t += 8) {
        uint32_t a4[4], b4[4];
        gsr_fr_load4(lb, t, a4);
        gsr_fr_load4(lb, t + 4, b4);
#pragma unroll
        for (int k = 0; k < 4; ++k) gsr_eval_quad<false>(a4[k], a4[k] + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
#pragma unroll
        for (int k = 0; k < 4; ++k) gsr_eval_quad<false>(b4[k], b4[k] + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
      }
#endif
      for (; t + 4 <= trip; t += 4) {
        uint32_t a4[4];
        gsr_fr_load4(lb, t, a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = a4[k];
          gsr_eval_quad<false>(a, a + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
        }
      }
      if (t < trip) {  // one to three left: two at once, then one
        uint32_t a4[4];
        gsr_fr_load4(lb, t, a4);
        gsr_eval_quad<false>(a4[0], a4[0] + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
        if (t + 1 < trip) gsr_eval_quad<false>(a4[1], a4[1] + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
        if (t + 2 < trip) gsr_eval_quad<false>(a4[2], a4[2] + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
      }
#else
      for (int t = 0; t < trip; t += 4) {
        uint32_t a4[4];
        gsr_fr_load4(lb, t, a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = a4[k];
          gsr_eval_quad<false>(a, a + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
        }
      }
#endif
    } else {
      const int uy = uA / p.nrx, ux = uA - uy * p.nrx;
      const int wi0 = ux * GSR_RGW + bx, hi0 = uy * GSR_RGH + by;
      for (int t = 0; t < trip; t += 4) {
        uint32_t a4[4];
        gsr_fr_load4(lb, t, a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = a4[k];
          const uint32_t slot = (a - rb) >> 4;
          const bool binds = slot < 32 ? ((slow_ac >> slot) & 1u) : (slot < 64 ? ((slow_bc >> (slot - 32)) & 1u) : false);
          bool m00 = true, m01 = true, m10 = true, m11 = true;
          if (binds) {  // exact inclusion
            int bx0, bx1, by0, by1;
            bool bd;
            gsr_box_unpack(box_w[cur * CH + slot], bx0, bx1, by0, by1, bd);
            const bool y0in = hi0 >= by0 && hi0 <= by1, y1in = hi0 + 1 >= by0 && hi0 + 1 <= by1;
            const bool x0in = wi0 >= bx0 && wi0 <= bx1, x1in = wi0 + 1 >= bx0 && wi0 + 1 <= bx1;
            m00 = y0in && x0in, m01 = y0in && x1in, m10 = y1in && x0in, m11 = y1in && x1in;
          }
          gsr_eval_quad<true>(a, a + GSR_FR_HI, nx2, ny2, m00, m01, m10, m11, r0, g0, b0, r1, g1, b1);
        }
      }
    }
    cur ^= 1;
    ++ci;
    trip = trip_n;
    slow_ac = slow_an;
    slow_bc = slow_bn;
    if (!last) continue;

    // ---- unit finished: write out
    {
      const int uy = uA / p.nrx, ux = uA - uy * p.nrx;
      float v[2][2][3];
      gsr_upk(r0, v[0][0][0], v[0][1][0]);
      gsr_upk(g0, v[0][0][1], v[0][1][1]);
      gsr_upk(b0, v[0][0][2], v[0][1][2]);
      gsr_upk(r1, v[1][0][0], v[1][1][0]);
      gsr_upk(g1, v[1][0][1], v[1][1][1]);
      gsr_upk(b1, v[1][0][2], v[1][1][2]);
      gsr_fr_write_unit<WINDOW>(p, tile_s, lane, ux, uy, bx, by, v);
    }
    // ---- advance: B becomes A, C becomes B, the next claimed unit becomes C
    uA = uB;
    if (uA >= nunits) break;
    nA = nB;
    nchA = nchB;
    uB = uC;
    nB = min(nC, p.reg_cap);  // requested one unit ago
    nchB = chunks_of(nB);
    uC = take_unit();
    nC = count_of(uC);
    ci = 0;
    nx2 = nx2B;
    ny2 = ny2B;
    coords_of(uB, nx2B, ny2B);
    r0 = g0 = b0 = r1 = g1 = b1 = gsr_pk(0.f, 0.f);
  }
  finish();
}
