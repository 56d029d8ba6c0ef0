// gsr_backward_region.cuh -- backward over the forward's region buckets (sm_100a).
//
// The Gaussian-centric backward (gsr_backward.cuh) sweeps every Gaussian's cull BOX one pixel per lane and step and
// spends half of its instructions outside the sweeps (staging, indexing, reductions per Gaussian).  This kernel walks
// the structure the forward walks instead: a warp per 16x8-pixel region, the bucket's entries staged in chunks of 64
// records (the same claim / request / cp.async pipeline as gsr_forward_region_kernel), every entry carrying the 8-bit
// mask of the region's 4x4-pixel cells its ellipse reaches.  A chunk's (cell, entry) pairs are laid out as ONE flat
// list -- cell after cell, each cell's entries in staging order -- and dealt to the lanes round-robin: a lane
// evaluates a pair's whole 4x4 cell by itself and gets the EIGHT sums the gradients are linear in
//     sum v g_r, sum v g_g, sum v g_b, Sx, Sy, Sxx, Sxy, Syy      (u = v sum_c g_c col_c; S* = sum u {dx, dy, dx^2, dx dy, dy^2})
// complete in its registers -- no cross-lane reduction, no idle lanes behind the longest cell list -- and adds them to
// the Gaussian's row of a moment array in global memory (2 vector REDs per pair); gsr_bwd_chain_kernel then applies
// the chain rule of gs.cu:139-159 once per Gaussian.  No home-bin sort.
// History (HL, kernel + chain + clearing): 2x2 block per lane, sums reduced over a cell's four lanes and added to
// per-slot accumulators with shared-memory float atomics (compare-and-swap loops) 1.58 ms; parked per (list position,
// cell) and gathered per entry 1.10 ms; a whole cell per lane (the reduction was 44 % of the instructions) 0.87 ms;
// three CTAs per SM, record prefetch 0.77 ms; the flat list with direct REDs (this version) 0.66 ms.
// Summation order follows the atomics: not bit-reproducible -- GSR_FLAG_DETERMINISTIC keeps the Gaussian-centric kernel.
#pragma once
#include "gsr_forward_ws.cuh"

#ifndef GSR_CFG_BR_MIN_CTAS
#define GSR_CFG_BR_MIN_CTAS 3
#endif
constexpr int GSR_BR_FLAT = 8 * GSR_FR_CHUNK + 32;  // items of a chunk's flat list: every cell of every entry, + padding

struct GsrBwdRegionSmem {
  float4 rec[GSR_FR_WARPS][2][2 * GSR_FR_SLOTS];   // per warp, 2 stages x { first float4 x 65, second float4 x 65 }
  uint2 box[GSR_FR_WARPS][2][GSR_FR_CHUNK];
  uint32_t flat[GSR_FR_WARPS][2][GSR_BR_FLAT];     // record address (16 bits: first 64 KB of the window) | cell << 16
  int gidx[GSR_FR_WARPS][2][GSR_FR_CHUNK];         // Gaussian index of every staged slot
  float4 gtile[GSR_FR_WARPS][8][4 * 3];            // dL/dimg of the unit being evaluated: [cell][row][channel] x 4 columns
  float4 cx[GSR_FR_WARPS][4], cy[GSR_FR_WARPS][2]; // its NEGATED pixel coordinates: 16 columns, 8 rows
};
struct GsrBwdRegionArgs {
  const float* grads;  // dL/dimg, (h,w,3) or (3,h,w) (flag 2)
  float* mom;          // (s, 8) moment rows, zero on entry
};

// One record against a whole 4x4 CELL, by one lane.  a0 = {x, y, a, b}, a1 = {c, r, g, bl}: the staged record;
// X / Y: NEGATED pixel coordinates of the cell's four columns / rows, so d = x - px: the odd moments come out with the
// opposite sign of the reference's dx = px - x (gsr_bwd_chain_kernel accounts for it).  gt: the cell's dL/dimg in
// shared memory, [row][channel][column]; mo: the Gaussian's moment row (nullptr: the null record that pads the list).
template <bool MASKED>
__device__ __forceinline__ void gsr_bwd_eval_cell(const float4 a0, const float4 a1, const float4 X, const float4 Y,
                                                  const float4* __restrict__ gt, unsigned xin, unsigned yin, float* mo) {
  const gsr_f2 x2 = gsr_pk(a0.x, a0.x), a2 = gsr_pk(a0.z, a0.z);
  const gsr_f2 dx0 = gsr_add2(gsr_pk(X.x, X.y), x2), dx1 = gsr_add2(gsr_pk(X.z, X.w), x2);
  const gsr_f2 ad0 = gsr_mul2(a2, dx0), ad1 = gsr_mul2(a2, dx1);
  const gsr_f2 cr = gsr_pk(a1.y, a1.y), cg = gsr_pk(a1.z, a1.z), cb = gsr_pk(a1.w, a1.w);
  const gsr_f2 b2 = gsr_pk(a0.w, a0.w), c2 = gsr_pk(a1.x, a1.x);
  const float ny[4] = {Y.x, Y.y, Y.z, Y.w};
  // every sum is kept as a pair over the two column halves until the end (rows enter as broadcast pairs)
  gsr_f2 Cr = gsr_pk(0.f, 0.f), Cg = Cr, Cb = Cr, Sx2 = Cr, Sy2 = Cr, Sxx2 = Cr, Sxy2 = Cr, Syy2 = Cr;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float dy = ny[r] + a0.y;
    const gsr_f2 dyb = gsr_pk(dy, dy);
    const gsr_f2 t12 = gsr_mul2(b2, dyb), t02 = gsr_mul2(gsr_mul2(c2, dyb), dyb);  // (scalar forms: same time)
    const gsr_f2 e0 = gsr_fma2(dx0, gsr_add2(ad0, t12), t02), e1 = gsr_fma2(dx1, gsr_add2(ad1, t12), t02);
    float ea, eb, ec, ed;
    gsr_upk(e0, ea, eb);
    gsr_upk(e1, ec, ed);
    const float va = gsr_ex2(ea), vb = gsr_ex2(eb), vc = gsr_ex2(ec), vd = gsr_ex2(ed);
    const gsr_f2 v0 = gsr_pk(va, vb), v1 = gsr_pk(vc, vd);
    float4 gr = gt[r * 3 + 0], gg = gt[r * 3 + 1], gb = gt[r * 3 + 2];
    if (MASKED) {  // pixels outside the dmax window: their gradient is not seen at all (v stays finite: no 0 * inf)
      const bool yr = (yin >> r) & 1u;
      const bool m0 = yr && (xin & 1u), m1 = yr && (xin & 2u), m2 = yr && (xin & 4u), m3 = yr && (xin & 8u);
      gr.x = m0 ? gr.x : 0.f, gg.x = m0 ? gg.x : 0.f, gb.x = m0 ? gb.x : 0.f;
      gr.y = m1 ? gr.y : 0.f, gg.y = m1 ? gg.y : 0.f, gb.y = m1 ? gb.y : 0.f;
      gr.z = m2 ? gr.z : 0.f, gg.z = m2 ? gg.z : 0.f, gb.z = m2 ? gb.z : 0.f;
      gr.w = m3 ? gr.w : 0.f, gg.w = m3 ? gg.w : 0.f, gb.w = m3 ? gb.w : 0.f;
    }
    const gsr_f2 gr0 = gsr_pk(gr.x, gr.y), gr1 = gsr_pk(gr.z, gr.w), gg0 = gsr_pk(gg.x, gg.y), gg1 = gsr_pk(gg.z, gg.w);
    const gsr_f2 gb0 = gsr_pk(gb.x, gb.y), gb1 = gsr_pk(gb.z, gb.w);
    const gsr_f2 u0 = gsr_mul2(v0, gsr_fma2(gr0, cr, gsr_fma2(gg0, cg, gsr_mul2(gb0, cb))));
    const gsr_f2 u1 = gsr_mul2(v1, gsr_fma2(gr1, cr, gsr_fma2(gg1, cg, gsr_mul2(gb1, cb))));
    Cr = gsr_fma2(v0, gr0, gsr_fma2(v1, gr1, Cr));
    Cg = gsr_fma2(v0, gg0, gsr_fma2(v1, gg1, Cg));
    Cb = gsr_fma2(v0, gb0, gsr_fma2(v1, gb1, Cb));
    const gsr_f2 ux0 = gsr_mul2(u0, dx0), ux1 = gsr_mul2(u1, dx1);
    Sxx2 = gsr_fma2(ux0, dx0, gsr_fma2(ux1, dx1, Sxx2));
    const gsr_f2 rx = gsr_add2(ux0, ux1), ry = gsr_mul2(gsr_add2(u0, u1), dyb);  // the row's u dx and u dy, by column half
    Sx2 = gsr_add2(Sx2, rx);
    Sxy2 = gsr_fma2(rx, dyb, Sxy2);
    Sy2 = gsr_add2(Sy2, ry);
    Syy2 = gsr_fma2(ry, dyb, Syy2);
  }
  float s[8];
  {
    const gsr_f2 all[8] = {Cr, Cg, Cb, Sx2, Sy2, Sxx2, Sxy2, Syy2};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float lo, hi;
      gsr_upk(all[i], lo, hi);
      s[i] = lo + hi;
    }
  }
  if (mo) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mo), "f"(s[0]), "f"(s[1]), "f"(s[2]), "f"(s[3]) : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mo + 4), "f"(s[4]), "f"(s[5]), "f"(s[6]), "f"(s[7]) : "memory");
  }
}

// The flat list of one staged chunk: item = record address | cell << 16, the items of cell 0 first, then cell 1, ...,
// within a cell in the order (lane 0: first, second entry), (lane 1: ...).  The eight ranks of an entry come from ONE
// warp scan (the masks spread to a byte per cell, prefix-summed across the lanes; counts stay below 256), the cells'
// offsets from the totals.  32 null items follow the last one (lanes read one item ahead).  Returns the item count.
// lw: the stage's list; rb: the stage's records; (v1a, e1a), (v1b, e1b): this lane's two entries.  Warp-collective.
__device__ __forceinline__ int gsr_br_build_flat(uint32_t lw, uint32_t rb, int lane, bool v1a, uint32_t e1a, bool v1b,
                                                 uint32_t e1b) {
  const unsigned full = 0xffffffffu;
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
  const uint32_t adr_a = rb + lane * 16u, adr_b = adr_a + 32 * 16u;
  uint32_t off = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t ra = ((q < 4 ? ra_lo : ra_hi) >> (8 * (q & 3))) & 0xffu;
    const uint32_t rbq = ((q < 4 ? rb_lo : rb_hi) >> (8 * (q & 3))) & 0xffu;
    if ((ma >> q) & 1u) gsr_sts32(lw + 4 * (off + ra), adr_a | ((uint32_t)q << 16));
    if ((mb >> q) & 1u) gsr_sts32(lw + 4 * (off + rbq), adr_b | ((uint32_t)q << 16));
    off += ((q < 4 ? t_lo : t_hi) >> (8 * (q & 3))) & 0xffu;
  }
  gsr_sts32(lw + 4 * (off + lane), rb + GSR_FR_CHUNK * 16u);  // null items
  return (int)off;
}

__global__ void __launch_bounds__(GSR_FR_THREADS, GSR_CFG_BR_MIN_CTAS) gsr_backward_region_kernel(GsrFwdArgs p, GsrBwdRegionArgs q) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  extern __shared__ __align__(16) unsigned char gsr_br_smem_raw[];
  GsrBwdRegionSmem& sm = *reinterpret_cast<GsrBwdRegionSmem*>(gsr_br_smem_raw);
  static_assert(offsetof(GsrBwdRegionSmem, box) < 65536, "16-bit record addresses: the records sit in the first 64 KB");
  constexpr int CH = GSR_FR_CHUNK;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nunits = p.nrx * p.nry;
  const uint32_t rec_s = gsr_smem_addr(&sm.rec[warp][0][0]);
  const uint32_t flat_w = gsr_smem_addr(&sm.flat[warp][0][0]);   // the warp's flat list, stage 0
  const uint32_t box_s = gsr_smem_addr(&sm.box[warp][0][0]);
  const uint2* box_w = &sm.box[warp][0][0];
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
  // Records of the requested entries -> stage `st` (cp.async), their (cell, entry) pairs -> flat list of stage `st`.
  // Returns the number of pairs, and the binds ballots.
  unsigned slow_a = 0, slow_b = 0;
  auto stage_chunk = [&](int st) -> int {
    const uint32_t rb = rec_s + st * GSR_FR_STAGE_BYTES;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t en = t ? e1b : e1a;
      const bool v = t ? v1b : v1a;
      const int k = lane + 32 * t;
      if (v) {
        const uint32_t gi = en & GSR_ENT_INDEX;
        const char* src = reinterpret_cast<const char*>(p.rec_in + gi);
        gsr_cp_async16ca(rb + k * 16, src);
        gsr_cp_async16ca(rb + GSR_FR_HI + k * 16, src + 16);
        if (en >> 31) gsr_cp_async8(box_s + (st * CH + k) * 8, p.box_in + gi);
        sm.gidx[warp][st][k] = (int)gi;  // whom the slot's sums belong to
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    slow_a = __ballot_sync(full, v1a && (e1a >> 31));
    slow_b = __ballot_sync(full, v1b && (e1b >> 31));
    return gsr_br_build_flat(flat_w + st * (GSR_BR_FLAT * 4), rb, lane, v1a, e1a, v1b, e1b);
  };

  // Units in flight: A is evaluated, B and C are known far enough ahead for the two-deep prefetch to run
  // across unit boundaries; further units wait in the claimed range [qn, qe) and in the claim in flight.
  int nA = min(count_of(uA), p.reg_cap), nB = min(count_of(uB), p.reg_cap), nC = count_of(uC);
  int nchA = chunks_of(nA), nchB = chunks_of(nB);

  request_entries(uA, 0, nA, nchA);
  int total = stage_chunk(0);
  unsigned slow_ac = slow_a, slow_bc = slow_b;
  if (1 < nchA) request_entries(uA, 1, nA, nchA); else request_entries(uB, 0, nB, nchB);

  int cur = 0, ci = 0;
  // What a unit needs besides its bucket: dL/dimg of its 16x8 pixels and their coordinates.  Fetched a unit ahead
  // into registers -- a 2x2 block of dL/dimg per lane ({row 0, row 1} x {r, g, b}, each a pair over the two columns;
  // pixels outside the image: 0) and one coordinate (lanes 0-15: columns, 16-23: rows) -- and laid down in the warp's
  // shared memory when the unit's turn comes.
  const int cell = lane >> 2;
  const int bx = (cell & 3) * GSR_CELL + (lane & 1) * 2, by = (cell >> 2) * GSR_CELL + ((lane >> 1) & 1) * 2;
  const bool chw = (p.flags & 2u) != 0;
  const size_t plane = (size_t)p.h * p.w;
  auto fetch_unit = [&](int u, gsr_f2 (&g)[2][3], float& coord) {
    const int uy = (u < nunits ? u : 0) / p.nrx, ux = (u < nunits ? u : 0) - uy * p.nrx;
    const int wi = ux * GSR_RGW + bx, hi = uy * GSR_RGH + by;
#pragma unroll
    for (int yy = 0; yy < 2; ++yy)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float v[2];
#pragma unroll
        for (int xx = 0; xx < 2; ++xx) {
          const bool in = u < nunits && hi + yy < p.h && wi + xx < p.w;
          const size_t pix = (size_t)min(hi + yy, p.h - 1) * p.w + min(wi + xx, p.w - 1);
          v[xx] = in ? __ldg(q.grads + (chw ? ch * plane + pix : pix * 3 + ch)) : 0.f;
        }
        g[yy][ch] = gsr_pk(v[0], v[1]);
      }
    coord = lane < 16 ? -__ldg(p.px_tab + min(ux * GSR_RGW + lane, p.w - 1))
                      : -__ldg(p.py_tab + min(uy * GSR_RGH + (lane & 7), p.h - 1));
  };
  auto store_unit = [&](const gsr_f2 (&g)[2][3], float coord) {  // tile: [cell][row][channel][column]
    float* base = reinterpret_cast<float*>(&sm.gtile[warp][cell][0]) + (((lane >> 1) & 1) * 2) * 12 + (lane & 1) * 2;
#pragma unroll
    for (int yy = 0; yy < 2; ++yy)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float lo, hi;
        gsr_upk(g[yy][ch], lo, hi);
        *reinterpret_cast<float2*>(base + yy * 12 + ch * 4) = make_float2(lo, hi);
      }
    if (lane < 16) reinterpret_cast<float*>(&sm.cx[warp][0])[lane] = coord;
    else if (lane < 24) reinterpret_cast<float*>(&sm.cy[warp][0])[lane - 16] = coord;
  };
  gsr_f2 gB[2][3];
  float coordB;
  fetch_unit(uA, gB, coordB);
  store_unit(gB, coordB);
  fetch_unit(uB, gB, coordB);
  __syncwarp();

  for (;;) {  // one chunk per iteration, flat over the warp's units
    const uint32_t rb = rec_s + cur * GSR_FR_STAGE_BYTES;
    const uint32_t fb = flat_w + cur * (GSR_BR_FLAT * 4);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();  // this chunk's records and list are visible; the other stage is free
    // records and list of the next chunk (its entries were requested one chunk ago) ...
    const int total_n = stage_chunk(cur ^ 1);
    const unsigned slow_an = slow_a, slow_bn = slow_b;
    // ... and the entries of the chunk after it: (A, ci + 2), (B, 0), (B, 1) or (C, 0)
    const bool last = ci + 1 >= nchA;
    if (!last) {
      if (ci + 2 < nchA) request_entries(uA, ci + 2, nA, nchA); else request_entries(uB, 0, nB, nchB);
    } else {
      if (1 < nchB) request_entries(uB, 1, nB, nchB);
      else { const int n = min(nC, p.reg_cap); request_entries(uC, 0, n, chunks_of(n)); }
    }

    // ---- evaluate: item lane, lane + 32, ... of the flat list; the item after next is fetched ahead (the list
    // ends in 32 null items)
    {
      const int uy = uA / p.nrx, ux = uA - uy * p.nrx;
      const bool slow = (slow_ac | slow_bc) != 0;
      int i = lane;
      uint32_t itn = gsr_lds32u(fb + 4 * i);
      float4 n0 = gsr_lds128(itn & 0xffffu), n1 = gsr_lds128((itn & 0xffffu) + GSR_FR_HI);
      for (; i < total; i += 32) {
        const float4 c0 = n0, c1 = n1;
        const uint32_t it = itn;
        itn = gsr_lds32u(fb + 4 * (i + 32));
        n0 = gsr_lds128(itn & 0xffffu), n1 = gsr_lds128((itn & 0xffffu) + GSR_FR_HI);
        const uint32_t cq = it >> 16, slot = ((it & 0xffffu) - rb) >> 4;  // (slot < CH: i < total)
        float* mo = q.mom + (size_t)sm.gidx[warp][cur][slot & (CH - 1)] * 8;
        const float4 X = sm.cx[warp][cq & 3], Y = sm.cy[warp][cq >> 2];
        const float4* gt = &sm.gtile[warp][cq][0];
        if (!slow) {
          gsr_bwd_eval_cell<false>(c0, c1, X, Y, gt, 15u, 15u, mo);
        } else {
          const bool binds = slot < 32 ? ((slow_ac >> slot) & 1u) : ((slow_bc >> (slot - 32)) & 1u);
          unsigned xin = 15u, yin = 15u;
          if (binds) {  // exact inclusion
            int bx0, bx1, by0, by1;
            bool bd;
            gsr_box_unpack(box_w[cur * CH + slot], bx0, bx1, by0, by1, bd);
            const int wi0 = ux * GSR_RGW + (int)(cq & 3) * GSR_CELL, hi0 = uy * GSR_RGH + (int)(cq >> 2) * GSR_CELL;
            xin = yin = 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              xin |= (wi0 + k >= bx0 && wi0 + k <= bx1) ? 1u << k : 0u;
              yin |= (hi0 + k >= by0 && hi0 + k <= by1) ? 1u << k : 0u;
            }
          }
          gsr_bwd_eval_cell<true>(c0, c1, X, Y, gt, xin, yin, mo);
        }
      }
    }
    cur ^= 1;
    ++ci;
    total = total_n;
    slow_ac = slow_an;
    slow_bc = slow_bn;
    if (!last) continue;

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
    __syncwarp();  // every lane is past its last read of the unit's tile and coordinates
    store_unit(gB, coordB);
    fetch_unit(uB, gB, coordB);
    __syncwarp();
  }
  finish();
}

// Chain rule of gs.cu:139-159 from a Gaussian's moment row (see gsr_bwd_chain in gsr_backward.cuh; here in INPUT order,
// with the odd moments negated: the region kernel measures d = x - px).  Outputs are accumulated into, like the
// reference's.  Rows that received nothing (skipped or invisible Gaussians) are left alone.
__global__ void __launch_bounds__(256)
gsr_bwd_chain_kernel(const float* __restrict__ mom, const GsrRec* __restrict__ rec_in, const float* __restrict__ sigmas,
                     float* __restrict__ g_sigmas, float* __restrict__ g_coords, float* __restrict__ g_colors, int s,
                     const int* guard, int want, const GsrBDesc* __restrict__ bdesc, int bn) {
  if (gsr_guard_skip(guard, want)) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s; i += gridDim.x * blockDim.x) {
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(mom) + 2 * (size_t)i);
    const float4 m1 = __ldg(reinterpret_cast<const float4*>(mom) + 2 * (size_t)i + 1);
    if (m0.x == 0.f && m0.y == 0.f && m0.z == 0.f && m0.w == 0.f && m1.x == 0.f && m1.y == 0.f && m1.z == 0.f && m1.w == 0.f)
      continue;
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(rec_in) + 2 * (size_t)i);
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(rec_in) + 2 * (size_t)i + 1);
    const float sgx = __ldg(sigmas + 3 * (size_t)i + 0), sgy = __ldg(sigmas + 3 * (size_t)i + 1), rho = __ldg(sigmas + 3 * (size_t)i + 2);
    const float Sx = -m0.w, Sy = -m1.x, Sxx = m1.y, Sxy = m1.z, Syy = m1.w;
    const float a = r0.z, b = r0.w, c = r1.x;
    const float iL = 0.6931471805599453f;  // 1 / log2(e)
    float gx = -(2.0f * a * Sx + b * Sy) * iL;
    float gy = -(2.0f * c * Sy + b * Sx) * iL;
    float sxy_own = Sxy;
    if (bdesc) {  // padded batch: see gsr_bwd_chain
      const GsrBDesc d = bdesc[i / bn];
      gx = (float)((double)gx / d.ax);
      gy = (float)((double)gy / d.ay);
      sxy_own = (float)((double)Sxy * d.ax * d.ay);
    }
    const float gsx = -(b * Sxy + 2.0f * a * Sxx) * iL / sgx;
    const float gsy = -(b * Sxy + 2.0f * c * Syy) * iL / sgy;
    const double Q = (double)a * Sxx + (double)b * Sxy + (double)c * Syy;
    const float grho = (float)((2.0 * (double)rho * Q * (double)iL + (double)sxy_own / ((double)sgx * sgy)) /
                               (1.0 - (double)rho * rho));
    float* os = g_sigmas + 3 * (size_t)i;
    float* oc = g_coords + 2 * (size_t)i;
    float* ok = g_colors + 3 * (size_t)i;
    os[0] += gsx;
    os[1] += gsy;
    os[2] += grho;
    oc[0] += gx;
    oc[1] += gy;
    ok[0] += m0.x;
    ok[1] += m0.y;
    ok[2] += m0.z;
  }
}

