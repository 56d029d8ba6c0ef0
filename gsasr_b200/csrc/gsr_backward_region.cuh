// gsr_backward_region.cuh -- backward over the forward's region buckets (sm_100a).
//
// The Gaussian-centric backward (gsr_backward.cuh) sweeps every Gaussian's cull BOX one pixel per lane and step and
// spends half of its instructions outside the sweeps (staging, indexing, reductions per Gaussian).  This kernel walks
// the structure the forward walks instead: a warp per 16x8-pixel region, a 2x2 pixel block per lane, per-cell lists
// of the bucket entries whose cell mask names the cell -- the same staging pipeline, list building and chunk loop as
// gsr_forward_region_kernel -- and per (entry, cell) accumulates the EIGHT sums the gradients are linear in
//     sum v g_r, sum v g_g, sum v g_b, Sx, Sy, Sxx, Sxy, Syy      (u = v sum_c g_c col_c; S* = sum u {dx, dy, dx^2, dx dy, dy^2})
// over the lane's four pixels, reduces them over the cell's four lanes (6 shuffles: recursive halving) and parks
// them in shared memory at (list position, cell) -- a (cell, entry) pair is evaluated once per chunk, so this is a
// plain 8-byte store per lane, no atomics (shared-memory float atomics are compare-and-swap loops: the first version
// of this kernel spent a third of its time in them).  After a block of list positions, lane k gathers the rows of
// ITS two staged entries (their cells and list ranks are what the list builder computed) into registers and, at the
// end of the chunk, adds them to the Gaussian's row of a moment array in global memory (2 vector REDs per entry and
// region); gsr_bwd_chain_kernel then applies the chain rule of gs.cu:139-159 once per Gaussian.  ~70 instructions per 128 (Gaussian, pixel) pairs
// against 26 per 32 in the box sweeps, none of the per-Gaussian overhead, and no home-bin sort.
// Summation order follows the atomics: not bit-reproducible -- GSR_FLAG_DETERMINISTIC keeps the Gaussian-centric kernel.
#pragma once
#include "gsr_forward_ws.cuh"

#ifndef GSR_CFG_BR_T
#define GSR_CFG_BR_T 32   // list positions parked before the entries gather their sums (16: one more CTA per SM)
#endif
#ifndef GSR_CFG_BR_MIN_CTAS
#define GSR_CFG_BR_MIN_CTAS (GSR_CFG_BR_T <= 16 ? 4 : 3)
#endif
constexpr int GSR_BR_T = GSR_CFG_BR_T;
constexpr int GSR_BR_ROW = 8 * 32 + 16;  // bytes per list position: 8 cells x 8 sums, padded (gather: conflict-free)

struct GsrBwdRegionSmem {
  float4 rec[GSR_FR_WARPS][2][2 * GSR_FR_SLOTS];
  uint2 box[GSR_FR_WARPS][2][GSR_FR_CHUNK];
  uint32_t list[GSR_FR_WARPS][2][GSR_FR_LIST_STAGE / 4];
  uint4 meta[GSR_FR_WARPS][2][2][32];  // per stage and lane: {masks, ranks a (2), index a}, {ranks b (2), index b, -}
  float4 park[GSR_FR_WARPS][GSR_BR_T * GSR_BR_ROW / 16];  // (list position, cell) -> the cell's eight sums
};
static_assert(offsetof(GsrBwdRegionSmem, box) < 65536, "16-bit list addresses: the records sit in the first 64 KB");
struct GsrBwdRegionArgs {
  const float* grads;  // dL/dimg, (h,w,3) or (3,h,w) (flag 2)
  float* mom;          // (s, 8) moment rows, zero on entry
};

// One record against this lane's 2x2 block: the eight sums over its four pixels, reduced over the cell's four lanes
// and parked at dst.  nx2 / ny2: negated pixel coordinates (d = x - px: the odd moments come out
// with the opposite sign of the reference's dx = px - x; gsr_bwd_chain_kernel accounts for it).
template <bool MASKED>
__device__ __forceinline__ void gsr_bwd_eval_quad(uint32_t addr0, uint32_t dst, gsr_f2 nx2, gsr_f2 ny2, bool m00,
                                                  bool m01, bool m10, bool m11, const gsr_f2 (&g)[2][3], int lane) {
  const float4 a0 = gsr_lds128(addr0);               // x, y, a, b
  const float4 a1 = gsr_lds128(addr0 + GSR_FR_HI);   // c, r, g, bl
  const gsr_f2 dx2 = gsr_add2(nx2, gsr_pk(a0.x, a0.x));
  const gsr_f2 dy2 = gsr_add2(ny2, gsr_pk(a0.y, a0.y));
  const gsr_f2 t1 = gsr_mul2(gsr_pk(a0.w, a0.w), dy2);
  const gsr_f2 t0 = gsr_mul2(gsr_mul2(gsr_pk(a1.x, a1.x), dy2), dy2);
  float t1a, t1b, t0a, t0b, dya, dyb;
  gsr_upk(t1, t1a, t1b);
  gsr_upk(t0, t0a, t0b);
  gsr_upk(dy2, dya, dyb);
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
  const gsr_f2 ua = gsr_mul2(va, gsr_fma2(g[0][0], cr, gsr_fma2(g[0][1], cg, gsr_mul2(g[0][2], cb))));
  const gsr_f2 ub = gsr_mul2(vb, gsr_fma2(g[1][0], cr, gsr_fma2(g[1][1], cg, gsr_mul2(g[1][2], cb))));
  const gsr_f2 dya2 = gsr_pk(dya, dya), dyb2 = gsr_pk(dyb, dyb);
  const gsr_f2 uxa = gsr_mul2(ua, dx2), uxb = gsr_mul2(ub, dx2), uya = gsr_mul2(ua, dya2), uyb = gsr_mul2(ub, dyb2);
  gsr_f2 pk8[8];  // each: {column 0, column 1} halves of one sum
  pk8[0] = gsr_fma2(va, g[0][0], gsr_mul2(vb, g[1][0]));
  pk8[1] = gsr_fma2(va, g[0][1], gsr_mul2(vb, g[1][1]));
  pk8[2] = gsr_fma2(va, g[0][2], gsr_mul2(vb, g[1][2]));
  pk8[3] = gsr_add2(uxa, uxb);
  pk8[4] = gsr_add2(uya, uyb);
  pk8[5] = gsr_fma2(uxa, dx2, gsr_mul2(uxb, dx2));
  pk8[6] = gsr_fma2(uxa, dya2, gsr_mul2(uxb, dyb2));
  pk8[7] = gsr_fma2(uya, dya2, gsr_mul2(uyb, dyb2));
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    gsr_upk(pk8[i], lo, hi);
    s[i] = lo + hi;
  }
  // recursive halving over the cell's four lanes: lane (bit 1, bit 0) ends with sums 4 * bit0 + 2 * bit1 + {0, 1}
  const bool b0 = lane & 1, b1 = lane & 2;
  float w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b0 ? s[i] : s[i + 4], keep = b0 ? s[i + 4] : s[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  float z[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b1 ? w[i] : w[i + 2], keep = b1 ? w[i + 2] : w[i];
    z[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  // dst: this lane's 8 bytes of the (list position, cell) row.  (The null record that pads the lists parks sums
  // nobody gathers.)
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(z[0]), "f"(z[1]) : "memory");
}

__global__ void __launch_bounds__(GSR_FR_THREADS, GSR_CFG_BR_MIN_CTAS) gsr_backward_region_kernel(GsrFwdArgs p, GsrBwdRegionArgs q) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  extern __shared__ __align__(16) unsigned char gsr_br_smem_raw[];
  GsrBwdRegionSmem& sm = *reinterpret_cast<GsrBwdRegionSmem*>(gsr_br_smem_raw);
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
  const uint32_t park_w = gsr_smem_addr(&sm.park[warp][0]);
  const uint32_t park_l = park_w + cell * 32 + ((lane & 1) ? 16 : 0) + ((lane & 2) ? 8 : 0);  // this lane's two sums of a row
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
    uint32_t rk[5];
    const int mine = gsr_fr_build_lists<true>(list_w + st * GSR_FR_LIST_STAGE, rb, lane, cell, v1a, e1a, v1b, e1b, rk);
    // whom the parked sums belong to, and where they are parked (read back by this lane only)
    sm.meta[warp][st][0][lane] = make_uint4(rk[0], rk[1], rk[2], v1a ? (e1a & GSR_ENT_INDEX) : 0xffffffffu);
    sm.meta[warp][st][1][lane] = make_uint4(rk[3], rk[4], v1b ? (e1b & GSR_ENT_INDEX) : 0xffffffffu, 0u);
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
  // dL/dimg of this lane's 2x2 block in unit u: {row 0, row 1} x {r, g, b}, each a pair over the two columns
  // (pixels outside the image: 0)
  const bool chw = (p.flags & 2u) != 0;
  const size_t plane = (size_t)p.h * p.w;
  auto grads_of = [&](int u, gsr_f2 (&g)[2][3]) {
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
  };
  gsr_f2 gA[2][3], gB[2][3];
  grads_of(uA, gA);
  grads_of(uB, gB);
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

    // ---- evaluate, a block of GSR_BR_T list positions at a time; after each block the lane gathers the sums of
    // its own two entries (entry k's row in cell q's list is its rank there)
    const uint4 mA = sm.meta[warp][cur][0][lane], mB = sm.meta[warp][cur][1][lane];
    float sa[8], sb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sa[i] = sb[i] = 0.f;
    const int uy = uA / p.nrx, ux = uA - uy * p.nrx;
    const int wi0 = ux * GSR_RGW + bx, hi0 = uy * GSR_RGH + by;
    for (int t0 = 0; t0 < trip; t0 += GSR_BR_T) {
      const int te = min(trip, t0 + GSR_BR_T);
      const uint32_t pk = park_l - t0 * GSR_BR_ROW;  // row of list position t: pk + t * GSR_BR_ROW
      if ((slow_ac | slow_bc) == 0) {
        int t = t0;
        for (; t + 4 <= te; t += 4) {
          uint32_t a4[4];
          gsr_fr_load4(lb, t, a4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            gsr_bwd_eval_quad<false>(a4[k], pk + (t + k) * GSR_BR_ROW, nx2, ny2, true, true, true, true, gA, lane);
        }
        if (t < te) {  // one to three left
          uint32_t a4[4];
          gsr_fr_load4(lb, t, a4);
          gsr_bwd_eval_quad<false>(a4[0], pk + t * GSR_BR_ROW, nx2, ny2, true, true, true, true, gA, lane);
          if (t + 1 < te) gsr_bwd_eval_quad<false>(a4[1], pk + (t + 1) * GSR_BR_ROW, nx2, ny2, true, true, true, true, gA, lane);
          if (t + 2 < te) gsr_bwd_eval_quad<false>(a4[2], pk + (t + 2) * GSR_BR_ROW, nx2, ny2, true, true, true, true, gA, lane);
        }
      } else {
        for (int t = t0; t < te; t += 4) {
          uint32_t a4[4];
          gsr_fr_load4(lb, t, a4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (t + k >= te) break;
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
            gsr_bwd_eval_quad<true>(a, pk + (t + k) * GSR_BR_ROW, nx2, ny2, m00, m01, m10, m11, gA, lane);
          }
        }
      }
      __syncwarp();  // the block's rows are parked
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t ra = (((c < 4 ? mA.y : mA.z) >> (8 * (c & 3))) & 0xffu) - (uint32_t)t0;
        const uint32_t rq = (((c < 4 ? mB.x : mB.y) >> (8 * (c & 3))) & 0xffu) - (uint32_t)t0;
        if (((mA.x >> c) & 1u) && ra < (uint32_t)GSR_BR_T) {
          const float4 r0 = gsr_lds128(park_w + ra * GSR_BR_ROW + c * 32), r1 = gsr_lds128(park_w + ra * GSR_BR_ROW + c * 32 + 16);
          sa[0] += r0.x, sa[1] += r0.y, sa[2] += r0.z, sa[3] += r0.w, sa[4] += r1.x, sa[5] += r1.y, sa[6] += r1.z, sa[7] += r1.w;
        }
        if (((mA.x >> (8 + c)) & 1u) && rq < (uint32_t)GSR_BR_T) {
          const float4 r0 = gsr_lds128(park_w + rq * GSR_BR_ROW + c * 32), r1 = gsr_lds128(park_w + rq * GSR_BR_ROW + c * 32 + 16);
          sb[0] += r0.x, sb[1] += r0.y, sb[2] += r0.z, sb[3] += r0.w, sb[4] += r1.x, sb[5] += r1.y, sb[6] += r1.z, sb[7] += r1.w;
        }
      }
      __syncwarp();  // gathered: the rows may be overwritten
    }
    // ---- the chunk's sums -> the Gaussians' moment rows
    if (mA.w != 0xffffffffu) {
      float* mo = q.mom + (size_t)mA.w * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mo), "f"(sa[0]), "f"(sa[1]), "f"(sa[2]), "f"(sa[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mo + 4), "f"(sa[4]), "f"(sa[5]), "f"(sa[6]), "f"(sa[7]) : "memory");
    }
    if (mB.z != 0xffffffffu) {
      float* mo = q.mom + (size_t)mB.z * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mo), "f"(sb[0]), "f"(sb[1]), "f"(sb[2]), "f"(sb[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mo + 4), "f"(sb[4]), "f"(sb[5]), "f"(sb[6]), "f"(sb[7]) : "memory");
    }
    cur ^= 1;
    ++ci;
    trip = trip_n;
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
    nx2 = nx2B;
    ny2 = ny2B;
    coords_of(uB, nx2B, ny2B);
#pragma unroll
    for (int yy = 0; yy < 2; ++yy)
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) gA[yy][ch] = gB[yy][ch];
    grads_of(uB, gB);
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

