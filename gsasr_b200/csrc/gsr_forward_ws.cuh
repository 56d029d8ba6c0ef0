// gsr_forward_ws.cuh -- warp-specialised form of the region-bucket forward kernel (sm_100a).
//
// gsr_forward_region_kernel lets every warp alternate between two kinds of work: turning a chunk of bucket
// entries into staged records and per-cell slot lists (integer / shuffle / shared-store work, latency bound) and
// evaluating the lists (MUFU + FP32 work, throughput bound).  With six warps per scheduler the phases of
// different warps only overlap by chance: the MUFU pipe idles while warps build lists and is oversubscribed when
// they evaluate (ncu: XU 54 %, issue 56 %, math-pipe throttle and not-selected stalls side by side).
//
// Here the two kinds of work live in DIFFERENT warps.  A CTA holds four pairs: warp i (consumer) and warp i + 4
// (producer) -- same scheduler, so every scheduler gets evaluating and staging warps.  The producer claims regions
// from the work counter and, per chunk of up to 64 entries, stages the records (cp.async), builds the eight
// per-cell slot lists (gsr_fr_build_lists) and a small header into one of THREE shared-memory stages of the pair;
// the consumer does nothing but wait for a stage, run the evaluation loop over its lists and, at the end of a
// region, write its 2x2 pixel block out.  Stages are handed over with two mbarriers each: `full` collects, per
// producer lane, one arrival for its shared-memory stores and one that the copy engine makes when the lane's
// cp.async copies have landed (cp.async.mbarrier.arrive.noinc) -- the producer never waits for its own copies;
// `empty` collects the 32 consumer lanes.  The evaluation loop, the slot lists and the semantics
// (cell masks, exact box test for window-binding Gaussians, write-out modes) are those of
// gsr_forward_region_kernel.
#pragma once
#include "gsr_forward.cuh"

constexpr int GSR_WS_PAIRS = 4;
constexpr int GSR_WS_THREADS = 64 * GSR_WS_PAIRS;
constexpr int GSR_WS_STAGES = 3;
#ifndef GSR_CFG_WS_MIN_CTAS
#define GSR_CFG_WS_MIN_CTAS 3
#endif
// Three CTAs per SM at 80 registers (twelve evaluating + twelve staging warps).  Measured alternative
// (-DGSR_CFG_WS_MIN_CTAS=4): the CTA is allocated 64 registers per thread, the warpgroup of producers gives
// registers back and the warpgroup of consumers takes them (setmaxnreg 48 / 80, 56 / 72, 40 / 88: sixteen + sixteen
// warps per SM) -- no faster at HL (302-306 us vs 302 us) and slower on small images (C2 75 vs 62 us).
#ifndef GSR_CFG_WS_SYNC_ARRIVE
#define GSR_CFG_WS_SYNC_ARRIVE 0
#endif
#ifndef GSR_CFG_WS_PROD_REGS
#define GSR_CFG_WS_PROD_REGS 48
#endif
#ifndef GSR_CFG_WS_CONS_REGS
#define GSR_CFG_WS_CONS_REGS 80
#endif
static_assert(GSR_CFG_WS_MIN_CTAS != 4 || GSR_CFG_WS_PROD_REGS + GSR_CFG_WS_CONS_REGS == 128, "register pool of a CTA");
#define GSR_STR2(x) #x
#define GSR_STR(x) GSR_STR2(x)

struct GsrWsStage {
  float4 rec[2 * GSR_FR_SLOTS];               // { first float4 x 65, second float4 x 65 }, slot 64 = null record
  uint32_t list[GSR_FR_LIST_STAGE / 4];       // eight cell lists of 16-bit shared addresses
  uint2 box[GSR_FR_CHUNK];                    // cull boxes of window-binding entries
  int hdr[8];                                 // trip, unit (-1: stop), flags (1 first, 2 last), slow_a, slow_b
};
struct GsrWsSmem {
  GsrWsStage st[GSR_WS_PAIRS][GSR_WS_STAGES];
  unsigned long long full[GSR_WS_PAIRS][GSR_WS_STAGES];
  unsigned long long empty[GSR_WS_PAIRS][GSR_WS_STAGES];
  float4 tile[GSR_WS_PAIRS][GSR_RGW * GSR_RGH * 3 / 4];  // write-out staging of the consumer warps
};
static_assert(GSR_FR_LW == 4 || sizeof(GsrWsSmem) + 1024 < 65536, "cell lists hold 16-bit shared-memory addresses");

__device__ __forceinline__ void gsr_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gsr_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gsr_mbar_wait(uint32_t bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}

template <bool WINDOW>
__global__ void __launch_bounds__(GSR_WS_THREADS, GSR_CFG_WS_MIN_CTAS) gsr_forward_region_ws_kernel(GsrFwdArgs p) {
  if (gsr_guard_skip(p.guard, p.want)) return;
  extern __shared__ __align__(16) unsigned char gsr_smem_raw[];
  GsrWsSmem& sm = *reinterpret_cast<GsrWsSmem*>(gsr_smem_raw);
  constexpr int CH = GSR_FR_CHUNK, NS = GSR_WS_STAGES;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp & (GSR_WS_PAIRS - 1);
  const bool producer = warp >= GSR_WS_PAIRS;
  const int cell = lane >> 2;
  const int nunits = p.nrx * p.nry;
  const uint32_t st_s = gsr_smem_addr(&sm.st[pair][0]);
  const uint32_t full_s = gsr_smem_addr(&sm.full[pair][0]), empty_s = gsr_smem_addr(&sm.empty[pair][0]);
  constexpr uint32_t STB = (uint32_t)sizeof(GsrWsStage);
  constexpr uint32_t LIST_OFF = (uint32_t)offsetof(GsrWsStage, list), BOX_OFF = (uint32_t)offsetof(GsrWsStage, box);

  if (warp < GSR_WS_PAIRS) {  // consumer warp i sets up pair i: barriers and the null records
    if (lane < NS) {
      gsr_mbar_init(full_s + lane * 8, 64);  // per producer lane: its copies have landed + its stores are done
      gsr_mbar_init(empty_s + lane * 8, 32);
      sm.st[pair][lane].rec[CH] = make_float4(0.f, 0.f, 0.f, 0.f);
      sm.st[pair][lane].rec[GSR_FR_SLOTS + CH] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // pixel block of this lane in a unit: cell (cell & 3, cell >> 2), block (lane & 1, (lane >> 1) & 1) of the cell
  const int bx = (cell & 3) * GSR_CELL + (lane & 1) * 2, by = (cell >> 2) * GSR_CELL + ((lane >> 1) & 1) * 2;

  if (producer) {
    // =============================== producer ===============================
#if GSR_CFG_WS_MIN_CTAS == 4
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " GSR_STR(GSR_CFG_WS_PROD_REGS) ";");
#endif
    const int total_prod = gridDim.x * GSR_WS_PAIRS;
    auto claim_size = [&](int progress) { return progress + 8 * total_prod < nunits ? 2 : 1; };
    int uA, uB, uC, qn, qe, pend, pend_n;
    {
      pend_n = claim_size(3 * total_prod);
      int base = 0;
      if (lane == 0) base = atomicAdd(p.sched, 3 + pend_n);
      base = __shfl_sync(full, base, 0);
      uA = base, uB = base + 1, uC = base + 2;
      qn = qe = base + 3;
      pend = base + 3;
    }
    auto take_unit = [&]() {
      if (qn == qe) {
        qn = __shfl_sync(full, pend, 0);
        qe = qn + pend_n;
        pend_n = claim_size(qn);
        asm volatile("{\n\t.reg .pred pl0;\n\tsetp.eq.s32 pl0, %2, 0;\n\t@pl0 atom.global.add.u32 %0, [%1], %3;\n\t}"
                     : "+r"(pend) : "l"(p.sched), "r"(lane), "r"(pend_n) : "memory");
      }
      return qn++;
    };
    auto count_of = [&](int u) { return u < nunits ? __ldg(p.reg_count + u) : 0; };
    auto chunks_of = [&](int n) { return n > CH ? (n + CH - 1) / CH : 1; };
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

    // Chunk generator: walks (unit, chunk) pairs two chunks AHEAD of the one being staged, so that a chunk's entries
    // -- requested when the generator passes it -- have two producer iterations to arrive (an iteration is shorter
    // than a global load now that the evaluation runs elsewhere).  Units: A is being walked, B and C are known one
    // and two units ahead (their bucket lengths are requested on arrival).
    int nA = min(count_of(uA), p.reg_cap), nB = min(count_of(uB), p.reg_cap), nC = count_of(uC);
    int nchA = chunks_of(nA), nchB = chunks_of(nB);
    int gci = 0;
    struct Chunk {
      int unit, flags;     // unit < 0: stop marker; flags: 1 first, 2 last
      uint32_t ea, eb;     // this lane's two entries
      bool va, vb;
    };
    auto next_chunk = [&]() {
      Chunk d;
      d.ea = d.eb = 0u;
      d.va = d.vb = false;
      if (uA >= nunits) {
        d.unit = -1;
        d.flags = 3;
        return d;
      }
      const bool last = gci + 1 >= nchA;
      d.unit = uA;
      d.flags = (gci == 0 ? 1 : 0) | (last ? 2 : 0);
      request_entries(uA, gci, nA, nchA);
      d.ea = e1a, d.eb = e1b, d.va = v1a, d.vb = v1b;
      if (!last) {
        ++gci;
      } else {
        uA = uB, nA = nB, nchA = nchB;
        uB = uC, nB = min(nC, p.reg_cap), nchB = chunks_of(nB);
        uC = take_unit();
        nC = count_of(uC);
        gci = 0;
      }
      return d;
    };
    Chunk d0 = next_chunk(), d1 = next_chunk(), d2 = next_chunk();
    int it = 0;
    for (;;) {  // one stage per iteration: chunk d0, or the stop marker
      const int s = it % NS;
      const uint32_t sb = st_s + s * STB;
      const bool stop = d0.unit < 0;
      gsr_mbar_wait(empty_s + s * 8, (((unsigned)it / NS) & 1u) ^ 1u);  // the consumer is done with this stage
      if (!stop) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t en = t ? d0.eb : d0.ea;
          const bool v = t ? d0.vb : d0.va;
          if (v) {
            const uint32_t gi = en & GSR_ENT_INDEX;
            const int k = lane + 32 * t;
            const char* src = reinterpret_cast<const char*>(p.rec_in + gi);
            gsr_cp_async16ca(sb + k * 16, src);
            gsr_cp_async16ca(sb + GSR_FR_HI + k * 16, src + 16);
            if (en >> 31) gsr_cp_async8(sb + BOX_OFF + k * 8, p.box_in + gi);
          }
        }
      }
      // the lane's share of "full": arrives by itself when the copies this lane issued so far have landed
#if GSR_CFG_WS_SYNC_ARRIVE
      // (checking aid: compute-sanitizer's racecheck does not follow cp.async.mbarrier.arrive -- it reports every
      // staged record as a write/read hazard; waiting for the copies and arriving explicitly is equivalent and clean)
      asm volatile("cp.async.wait_all;" ::: "memory");
      gsr_mbar_arrive(full_s + s * 8);
#else
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full_s + s * 8) : "memory");
#endif
      int trip = 0;
      unsigned slow_a = 0, slow_b = 0;
      if (!stop) {
        slow_a = __ballot_sync(full, d0.va && (d0.ea >> 31));
        slow_b = __ballot_sync(full, d0.vb && (d0.eb >> 31));
        const int mine = gsr_fr_build_lists(sb + LIST_OFF, sb, lane, cell, d0.va, d0.ea, d0.vb, d0.eb);
        trip = (__reduce_max_sync(full, mine) + 3) & ~3;
      }
      if (lane == 0) {
        int* h = sm.st[pair][s].hdr;
        h[0] = trip;
        h[1] = d0.unit;
        h[2] = d0.flags;
        h[3] = (int)slow_a;
        h[4] = (int)slow_b;
      }
      __syncwarp();                       // lane 0's header
      gsr_mbar_arrive(full_s + s * 8);    // this lane's lists (and the header) are written
      ++it;
      if (stop) break;
      d0 = d1;
      d1 = d2;
      d2 = next_chunk();
    }
    // the last producer warp to leave resets the work counters for the next launch
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(p.sched + 1, 1) == total_prod - 1) {
        p.sched[0] = 0;
        p.sched[1] = 0;
      }
    }
    return;
  }

  // =============================== consumer ===============================
#if GSR_CFG_WS_MIN_CTAS == 4
  asm volatile("setmaxnreg.inc.sync.aligned.u32 " GSR_STR(GSR_CFG_WS_CONS_REGS) ";");
#endif
  const uint32_t tile_s = gsr_smem_addr(&sm.tile[pair][0]);
  gsr_f2 nx2 = gsr_pk(0.f, 0.f), ny2 = nx2;
  gsr_f2 r0 = gsr_pk(0.f, 0.f), g0 = r0, b0 = r0, r1 = r0, g1 = r0, b1 = r0;
  for (int it = 0;; ++it) {
    const int s = it % NS;
    const uint32_t sb = st_s + s * STB;
    const GsrWsStage& stg = sm.st[pair][s];
    gsr_mbar_wait(full_s + s * 8, ((unsigned)it / NS) & 1u);
    const int trip = stg.hdr[0], unit = stg.hdr[1], flags = stg.hdr[2];
    const unsigned slow_a = (unsigned)stg.hdr[3], slow_b = (unsigned)stg.hdr[4];
    if (unit < 0) break;
    const uint32_t lb = sb + LIST_OFF + cell * GSR_FR_LIST;
    const int uy = unit / p.nrx, ux = unit - uy * p.nrx;
    const int wi0 = ux * GSR_RGW + bx, hi0 = uy * GSR_RGH + by;
    if (flags & 1) {  // first chunk of a unit: this lane's (negated) pixel coordinates -- the two tables are a few
                      // KB that every warp of the SM reads: L1 hits --, fresh accumulators
      nx2 = gsr_pk(-__ldg(p.px_tab + min(wi0, p.w - 1)), -__ldg(p.px_tab + min(wi0 + 1, p.w - 1)));
      ny2 = gsr_pk(-__ldg(p.py_tab + min(hi0, p.h - 1)), -__ldg(p.py_tab + min(hi0 + 1, p.h - 1)));
      r0 = g0 = b0 = r1 = g1 = b1 = gsr_pk(0.f, 0.f);
    }
    if ((slow_a | slow_b) == 0) {
      for (int t = 0; t < trip; t += 4) {
        uint32_t a4[4];
        gsr_fr_load4(lb, t, a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = a4[k];
          gsr_eval_quad<false>(a, a + GSR_FR_HI, nx2, ny2, true, true, true, true, r0, g0, b0, r1, g1, b1);
        }
      }
    } else {
      for (int t = 0; t < trip; t += 4) {
        uint32_t a4[4];
        gsr_fr_load4(lb, t, a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t a = a4[k];
          const uint32_t slot = (a - sb) >> 4;
          const bool binds = slot < 32 ? ((slow_a >> slot) & 1u) : (slot < 64 ? ((slow_b >> (slot - 32)) & 1u) : false);
          bool m00 = true, m01 = true, m10 = true, m11 = true;
          if (binds) {  // exact inclusion
            int bx0, bx1, by0, by1;
            bool bd;
            gsr_box_unpack(stg.box[slot], bx0, bx1, by0, by1, bd);
            const bool y0in = hi0 >= by0 && hi0 <= by1, y1in = hi0 + 1 >= by0 && hi0 + 1 <= by1;
            const bool x0in = wi0 >= bx0 && wi0 <= bx1, x1in = wi0 + 1 >= bx0 && wi0 + 1 <= bx1;
            m00 = y0in && x0in, m01 = y0in && x1in, m10 = y1in && x0in, m11 = y1in && x1in;
          }
          gsr_eval_quad<true>(a, a + GSR_FR_HI, nx2, ny2, m00, m01, m10, m11, r0, g0, b0, r1, g1, b1);
        }
      }
    }
    gsr_mbar_arrive(empty_s + s * 8);  // the stage may be refilled
    if (!(flags & 2)) continue;

    // ---- unit finished: write out (as gsr_forward_region_kernel)
    float v[2][2][3];
    gsr_upk(r0, v[0][0][0], v[0][1][0]);
    gsr_upk(g0, v[0][0][1], v[0][1][1]);
    gsr_upk(b0, v[0][0][2], v[0][1][2]);
    gsr_upk(r1, v[1][0][0], v[1][1][0]);
    gsr_upk(g1, v[1][0][1], v[1][1][1]);
    gsr_upk(b1, v[1][0][2], v[1][1][2]);
    gsr_fr_write_unit<WINDOW>(p, tile_s, lane, ux, uy, bx, by, v);
  }
}
