// gsr_head.cu -- head-tail fusion: the five per-Gaussian MLPs of the fea2gs head
// (utils/fea2gs.py:496-551, 611-633) on the 5th-generation tensor cores.  See gsr_head.cuh.
#include "../../include/gsraster.h"
#include "gsr_umma.cuh"
#include <atomic>
#include <type_traits>

#define GSH_CUDA(x)                       \
  do {                                    \
    cudaError_t e_ = (x);                 \
    if (e_ != cudaSuccess) return GSR_ERR_CUDA; \
  } while (0)

typedef CUresult (*gsh_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static gsh_encode_fn gsh_encoder() {
  static std::atomic<void*> cache{nullptr};
  void* fn = cache.load(std::memory_order_relaxed);
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    cache.store(fn, std::memory_order_relaxed);
  }
  return (gsh_encode_fn)fn;
}
// rows x cols bf16, row-major (cols contiguous) -> tiles of box_rows x 64 columns, 128-byte swizzle
static int gsh_map_bf16(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  gsh_encode_fn enc = gsh_encoder();
  if (!enc) return GSR_ERR_CUDA;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GSR_OK : GSR_ERR_CUDA;
}

// ---- stage 1: one 128 x N tile of C = A * B^T (A: M x K, B: N x K, bf16, K contiguous; C fp32) per CTA ----------
constexpr int GSH_T_THREADS = 192;  // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue
template <int N, int KB>
__global__ void __launch_bounds__(GSH_T_THREADS) gsh_gemm_test_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                     const __grid_constant__ CUtensorMap tm_b,
                                                                     float* __restrict__ c, int ldc) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (gsu::smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms: 1024-byte aligned
  constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;
  constexpr int TCOLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
  unsigned char* sa = smem;
  unsigned char* sb = smem + KB * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + KB * B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t full = gsu::smem_u32(bars), done = gsu::smem_u32(bars + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  if (warp == 0 && lane == 0) {
    gsu::tma_prefetch_desc(&tm_a);
    gsu::tma_prefetch_desc(&tm_b);
    gsu::mbar_init(full, 1);
    gsu::mbar_init(done, 1);
    gsu::fence_barrier_init();
  }
  if (warp == 1) gsu::tmem_alloc<TCOLS>(gsu::smem_u32(tmem_slot));
  gsu::fence_before_sync();
  __syncthreads();
  gsu::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    if (gsu::elect_one()) {
      gsu::mbar_expect_tx(full, KB * (A_BYTES + B_BYTES));
      for (int kb = 0; kb < KB; ++kb) {
        gsu::tma_load_2d(gsu::smem_u32(sa + kb * A_BYTES), &tm_a, full, kb * 64, m0);
        gsu::tma_load_2d(gsu::smem_u32(sb + kb * B_BYTES), &tm_b, full, kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    gsu::mbar_wait(full, 0);
    gsu::fence_after_sync();
    if (gsu::elect_one()) {
      constexpr uint32_t idesc = gsu::idesc_bf16_f32(128, N);
      for (int kb = 0; kb < KB; ++kb)
        for (int k = 0; k < 4; ++k)
          gsu::umma_bf16(tmem, gsu::smem_desc_k_sw128(gsu::smem_u32(sa + kb * A_BYTES) + k * 32),
                         gsu::smem_desc_k_sw128(gsu::smem_u32(sb + kb * B_BYTES) + k * 32), idesc, (kb | k) != 0);
      gsu::umma_commit(done);
    }
    __syncwarp();
  } else {
    gsu::mbar_wait(done, 0);
    gsu::fence_after_sync();
    const int q = warp & 3, row = q * 32 + lane;  // a warp reads the TMEM lanes of its own quadrant
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      gsu::tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      gsu::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) c[(size_t)(m0 + row) * ldc + c0 + j] = __uint_as_float(v[j]);
    }
  }
  gsu::fence_before_sync();
  __syncthreads();
  if (warp == 1) gsu::tmem_dealloc<TCOLS>(tmem);
}

// ---- the head tail: five MLPs C -> C -> 4C -> k, k = (2, 1, 1, 3, 2), one 128-row tile per CTA ------------------
// utils/fea2gs.py:496-541 (mlp_block_sigma / rho / alpha / rgb / mean: Linear, ReLU, Linear, ReLU, Linear) and
// :611-633 (concatenation, query_mean / grid size + reference points).  Padded sizes: CP = 192 input / first hidden
// channels (C = 180 or 192), HP = 768 second hidden channels (4C = 720 or 768); padding weights are zero.
//
//   warp 0    TMA producer: the tile's activations X (128 x 192 bf16, three swizzled K-blocks), then the stream of
//             weight tiles through a ring of GSH_STAGES stages: per head W1 (three 192 x 64 tiles) and W2 (six
//             chunks of 128 output channels x three 128 x 64 tiles);
//   warp 1    MMA issuer (one elected lane): layer 1  D1[128 x 192] = X W1^T  (12 tcgen05.mma, N = 192) into TMEM
//             columns 0..191; layer 2 per chunk  D2[128 x 128] = H1 W2c^T  (12 tcgen05.mma, N = 128) into one of two
//             TMEM buffers (columns 256.., 384..), so that chunk c+1 is multiplied while chunk c is read out -- and
//             the next head's layer 1 while the last chunks of this head are;
//   warps 2.. epilogue, GSH_PARTS threads per row (each a share of the columns): D1 -> registers (tcgen05.ld) -> + b1, ReLU, bf16 -> H1 in shared memory
//             in the operand's own swizzled K-major layout (layer 2 reads it as its A operand: the hidden layer never
//             leaves the SM); D2 chunk -> registers -> + b2, ReLU -> layer 3 on the CUDA cores (k <= 3 dot products
//             of 768 per row, accumulated across the chunks; the 4C-wide hidden layer never exists in memory at all)
//             -> + b3 -> raw parameter columns; the mean columns get the grid normalisation and the reference point
//             of the row's grid position (:624-631, torch.linspace's own fp32 formula).
constexpr int GSH_CP = 192, GSH_HP = 768, GSH_KB = GSH_CP / 64, GSH_CHUNK = 128, GSH_NCHUNK = GSH_HP / GSH_CHUNK;
constexpr int GSH_HEADS = 5, GSH_STAGES = 4;
constexpr int GSH_L1_AFTER = 3;  // the next head's layer 1 is multiplied after this chunk of the current head's layer 2:
                                 // its accumulator is ready by the time the epilogue warps finish the current head
#ifndef GSH_CFG_CLUSTER
#define GSH_CFG_CLUSTER 1
#endif
// CTAs per cluster.  With more than one, every weight tile is fetched from L2 ONCE per cluster: each CTA loads
// 1 / GSH_CLUSTER of its rows and multicasts them into every CTA's ring (1.8 MB of weights stream through an SM per
// 128-row tile).  Measured at 2.1M rows: 1 CTA 3.64 ms, 2 CTAs 3.76 ms, 4 CTAs 4.97 ms -- the L2 keeps up with
// single CTAs and the clusters' lock-step costs more than the traffic they save, so 1 is the default.
constexpr int GSH_CLUSTER = GSH_CFG_CLUSTER;
static_assert(GSH_CLUSTER == 1 || GSH_CLUSTER == 2 || GSH_CLUSTER == 4, "weight tiles are cut into 1, 2 or 4 row slices of whole swizzle atoms");
constexpr int GSH_XT = 128 * 128;            // one K-block of a 128-row operand tile, bytes
constexpr int GSH_W1T = GSH_CP * 128;        // W1 tile: 192 output channels x 64 k
constexpr int GSH_W2T = GSH_CHUNK * 128;     // W2 tile: 128 output channels x 64 k
constexpr int GSH_STAGE_BYTES = GSH_W1T;
#ifndef GSH_CFG_EPI_PARTS
#define GSH_CFG_EPI_PARTS 4
#endif
constexpr int GSH_PARTS = GSH_CFG_EPI_PARTS;  // epilogue warps per TMEM lane quadrant: each takes 1 / PARTS of the columns of a row
constexpr int GSH_EPI_WARPS = 4 * GSH_PARTS;
constexpr int GSH_EPI_THREADS = 32 * GSH_EPI_WARPS;
constexpr int GSH_TAB_BYTES = 2 * GSH_HP * 16;   // two heads' tables {b2, w3 row 0, w3 row 1, w3 row 2} per hidden channel (current, next)
constexpr int GSH_B1_BYTES = GSH_HEADS * GSH_CP * 4;
constexpr int GSH_EXCH_BYTES = (GSH_PARTS - 1) * 128 * 4 * 4;  // partial dot products of the other column parts
constexpr int GSH_SMEM = 2 * GSH_KB * GSH_XT + GSH_STAGES * GSH_STAGE_BYTES + GSH_TAB_BYTES + GSH_B1_BYTES + GSH_EXCH_BYTES + 256;
constexpr int GSH_THREADS = 64 + GSH_EPI_THREADS;
static_assert(GSH_SMEM + 1024 <= 232448, "227 KB of shared memory per CTA");
constexpr uint32_t GSH_D2_COL0 = 256, GSH_D2_COL1 = 384;

#ifndef GSH_CFG_TRACE
#define GSH_CFG_TRACE 0
#endif
// development aid: nanosecond timestamps of CTA 0's hand-overs (tools/trace_head.py)
#if GSH_CFG_TRACE
#define GSH_TRACE(slot) do { if (blockIdx.x == 0 && p.trace) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[slot] = t_; } } while (0)
#else
#define GSH_TRACE(slot) do { } while (0)
#endif
struct GshArgs {
  const float* b1;   // [5][CP]
  const float* b2;   // [5][HP]
  const float* w3;   // [9][HP]: rows in output order sigma_x, sigma_y, rho, alpha, r, g, b, mean_x, mean_y
  const float* b3;   // [9]
  float* raw;        // [m][9]
  int m;             // rows
  int gh, gw;        // the Gaussian grid of one sample: row index = (sample, iy, ix)
  unsigned long long* trace;  // GSH_CFG_TRACE builds only
};

// torch.linspace(start, end, steps) in fp32 (ATen's symmetric formula): element i
__device__ __forceinline__ float gsh_linspace(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  const float step = (end - start) / (float)(steps - 1);
  return i < steps / 2 ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

__global__ void __cluster_dims__(GSH_CLUSTER, 1, 1) __launch_bounds__(GSH_THREADS, 1)
gsh_head_tail_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2, const GshArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (gsu::smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sx = smem;                          // X: 3 K-blocks
  unsigned char* sh = sx + GSH_KB * GSH_XT;          // H1: 3 K-blocks
  unsigned char* sw = sh + GSH_KB * GSH_XT;          // weight ring
  float4* tab = reinterpret_cast<float4*>(sw + GSH_STAGES * GSH_STAGE_BYTES);
  float* sb1 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tab) + GSH_TAB_BYTES);
  float4* exch = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(sb1) + GSH_B1_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(exch) + GSH_EXCH_BYTES);
  // barriers: 0 x_full | 1..S w_full | S+1..2S w_empty | then d1_full, h1_ready, h1_free, d2_full[2], d2_empty[2]
  const uint32_t bar0 = gsu::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int X_FULL = 0, W_FULL = 1, W_EMPTY = 1 + GSH_STAGES, D1_FULL = 1 + 2 * GSH_STAGES, H1_READY = D1_FULL + 1,
                H1_FREE = D1_FULL + 2, D2_FULL = D1_FULL + 3, D2_EMPTY = D1_FULL + 5, NBARS = D1_FULL + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBARS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;

  if (warp == 0 && lane == 0) {
    gsu::tma_prefetch_desc(&tm_x);
    gsu::tma_prefetch_desc(&tm_w1);
    gsu::tma_prefetch_desc(&tm_w2);
    gsu::mbar_init(BAR(X_FULL), 1);
    for (int s = 0; s < GSH_STAGES; ++s) {
      gsu::mbar_init(BAR(W_FULL + s), 1);
      gsu::mbar_init(BAR(W_EMPTY + s), GSH_CLUSTER);  // every CTA of the cluster has consumed the stage
    }
    gsu::mbar_init(BAR(D1_FULL), 1);
    gsu::mbar_init(BAR(H1_READY), GSH_EPI_THREADS);
    gsu::mbar_init(BAR(H1_FREE), 1);
    for (int b = 0; b < 2; ++b) {
      gsu::mbar_init(BAR(D2_FULL + b), 1);
      gsu::mbar_init(BAR(D2_EMPTY + b), GSH_EPI_THREADS);
    }
    gsu::fence_barrier_init();
  }
  if (warp == 1) gsu::tmem_alloc<512>(gsu::smem_u32(tmem_slot));
  gsu::fence_before_sync();
  if (GSH_CLUSTER > 1) gsu::cluster_sync();  // the peers' barriers are initialised before anything is multicast to them
  else __syncthreads();
  gsu::fence_after_sync();
  const uint32_t crank = GSH_CLUSTER > 1 ? gsu::cluster_ctarank() : 0u;
  constexpr uint16_t CMASK = (uint16_t)((1u << GSH_CLUSTER) - 1u);
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (gsu::elect_one()) {
      gsu::mbar_expect_tx(BAR(X_FULL), GSH_KB * GSH_XT);
      for (int kb = 0; kb < GSH_KB; ++kb) gsu::tma_load_2d(gsu::smem_u32(sx + kb * GSH_XT), &tm_x, BAR(X_FULL), kb * 64, m0);
      int t = 0;
      auto stage_of = [&]() {  // next ring stage, free again
        const int s = t % GSH_STAGES;
        gsu::mbar_wait(BAR(W_EMPTY + s), ((uint32_t)(t / GSH_STAGES) & 1u) ^ 1u);
        ++t;
        return s;
      };
      auto load_w1 = [&](int h) {  // three K-blocks of W1_h
        for (int kb = 0; kb < GSH_KB; ++kb) {
          const int s = stage_of();
          const uint32_t dst = gsu::smem_u32(sw + s * GSH_STAGE_BYTES);
          gsu::mbar_expect_tx(BAR(W_FULL + s), GSH_W1T);  // the whole tile: this CTA's slice and the peers'
          if (GSH_CLUSTER > 1)
            gsu::tma_load_2d_mc(dst + crank * (GSH_W1T / GSH_CLUSTER), &tm_w1, BAR(W_FULL + s), kb * 64,
                                h * GSH_CP + (int)crank * (GSH_CP / GSH_CLUSTER), CMASK);
          else
            gsu::tma_load_2d(dst, &tm_w1, BAR(W_FULL + s), kb * 64, h * GSH_CP);
        }
      };
      auto load_w2 = [&](int h, int c) {  // three K-blocks of chunk c of W2_h
        for (int kb = 0; kb < GSH_KB; ++kb) {
          const int s = stage_of();
          const uint32_t dst = gsu::smem_u32(sw + s * GSH_STAGE_BYTES);
          gsu::mbar_expect_tx(BAR(W_FULL + s), GSH_W2T);
          if (GSH_CLUSTER > 1)
            gsu::tma_load_2d_mc(dst + crank * (GSH_W2T / GSH_CLUSTER), &tm_w2, BAR(W_FULL + s), kb * 64,
                                h * GSH_HP + c * GSH_CHUNK + (int)crank * (GSH_CHUNK / GSH_CLUSTER), CMASK);
          else
            gsu::tma_load_2d(dst, &tm_w2, BAR(W_FULL + s), kb * 64, h * GSH_HP + c * GSH_CHUNK);
        }
      };
      load_w1(0);  // (the same order as the MMA issuer's: see GSH_L1_AFTER)
      for (int h = 0; h < GSH_HEADS; ++h)
        for (int c = 0; c < GSH_NCHUNK; ++c) {
          load_w2(h, c);
          if (c == GSH_L1_AFTER && h + 1 < GSH_HEADS) load_w1(h + 1);
        }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (gsu::elect_one()) {
      constexpr uint32_t idesc1 = gsu::idesc_bf16_f32(128, GSH_CP), idesc2 = gsu::idesc_bf16_f32(128, GSH_CHUNK);
      gsu::mbar_wait(BAR(X_FULL), 0);
      gsu::fence_after_sync();
      int t = 0;
      auto stage_of = [&]() {  // next ring stage, filled
        const int s = t % GSH_STAGES;
        gsu::mbar_wait(BAR(W_FULL + s), (uint32_t)(t / GSH_STAGES) & 1u);
        gsu::fence_after_sync();
        ++t;
        return s;
      };
      auto release = [&](int s) {
        if (GSH_CLUSTER > 1) gsu::umma_commit_mc(BAR(W_EMPTY + s), CMASK); else gsu::umma_commit(BAR(W_EMPTY + s));
      };
      auto mma_l1 = [&](int h) {  // D1 = X W1_h^T
        GSH_TRACE(h * 64 + 0);
        for (int kb = 0; kb < GSH_KB; ++kb) {
          const int s = stage_of();
          const uint32_t a = gsu::smem_u32(sx + kb * GSH_XT), b = gsu::smem_u32(sw + s * GSH_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            gsu::umma_bf16(tmem, gsu::smem_desc_k_sw128(a + k * 32), gsu::smem_desc_k_sw128(b + k * 32), idesc1, (kb | k) != 0);
          release(s);
        }
        gsu::umma_commit(BAR(D1_FULL));
        GSH_TRACE(h * 64 + 1);
      };
      mma_l1(0);
      for (int h = 0; h < GSH_HEADS; ++h) {
        gsu::mbar_wait(BAR(H1_READY), (uint32_t)h & 1u);  // the epilogue has written this head's hidden layer (and read D1)
        gsu::fence_after_sync();
        GSH_TRACE(h * 64 + 2);
        for (int c = 0; c < GSH_NCHUNK; ++c) {            // layer 2
          const int u = h * GSH_NCHUNK + c, buf = u & 1;
          gsu::mbar_wait(BAR(D2_EMPTY + buf), ((uint32_t)(u >> 1) & 1u) ^ 1u);
          gsu::fence_after_sync();
          GSH_TRACE(h * 64 + 8 + 4 * c);
          const uint32_t d = tmem + (buf ? GSH_D2_COL1 : GSH_D2_COL0);
          for (int kb = 0; kb < GSH_KB; ++kb) {
            const int s = stage_of();
            const uint32_t a = gsu::smem_u32(sh + kb * GSH_XT), b = gsu::smem_u32(sw + s * GSH_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              gsu::umma_bf16(d, gsu::smem_desc_k_sw128(a + k * 32), gsu::smem_desc_k_sw128(b + k * 32), idesc2, (kb | k) != 0);
            release(s);
          }
          gsu::umma_commit(BAR(D2_FULL + buf));
          GSH_TRACE(h * 64 + 9 + 4 * c);
          if (c == GSH_L1_AFTER && h + 1 < GSH_HEADS) mma_l1(h + 1);  // D1 is free: its readers are behind H1_READY
        }
        gsu::umma_commit(BAR(H1_FREE));  // every multiplication that reads H1 has completed
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue: GSH_PARTS threads per row ===============================
    // warp w reads the TMEM lanes of quadrant w & 3; the warps of a quadrant split the columns of every piece of work
    const int ew = warp - 2, q = warp & 3, part = ew >> 2, row = q * 32 + lane, et = threadIdx.x - 64;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const bool live = m0 + row < p.m;
    const uint32_t hrow = gsu::smem_u32(sh) + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
    auto epi_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(GSH_EPI_THREADS) : "memory"); };
    for (int i = et; i < GSH_HEADS * GSH_CP; i += GSH_EPI_THREADS) sb1[i] = __ldg(p.b1 + i);
    auto kout_of = [](int h) { return h == 3 ? 3 : (h == 0 || h == 4) ? 2 : 1; };
    auto koff_of = [](int h) { return h == 0 ? 0 : h == 1 ? 2 : h == 2 ? 3 : h == 3 ? 4 : 7; };
    // table of head h, two float4 per PAIR of hidden channels (n, n+1): {b2[n], b2[n+1], w3_0[n], w3_0[n+1]} and
    // {w3_1[n], w3_1[n+1], w3_2[n], w3_2[n+1]}: the chunk loop works on channel pairs with packed FP32x2 arithmetic
    auto tab_entry = [&](int h, int i) {  // i: float4 index, 0 .. GSH_HP - 1
      const int kout = kout_of(h), n = (i >> 1) * 2;
      const float* w3 = p.w3 + (size_t)koff_of(h) * GSH_HP + n;
      if ((i & 1) == 0) return make_float4(__ldg(p.b2 + h * GSH_HP + n), __ldg(p.b2 + h * GSH_HP + n + 1), __ldg(w3), __ldg(w3 + 1));
      return make_float4(kout > 1 ? __ldg(w3 + GSH_HP) : 0.f, kout > 1 ? __ldg(w3 + GSH_HP + 1) : 0.f,
                         kout > 2 ? __ldg(w3 + 2 * GSH_HP) : 0.f, kout > 2 ? __ldg(w3 + 2 * GSH_HP + 1) : 0.f);
    };
    // a finished head's outputs: this thread's sums + the other column parts' (exchange buffer xb)
    auto write_out = [&](int h, float a0, float a1, float a2, const float4* xb) {
      float tot[3] = {a0, a1, a2};
#pragma unroll
      for (int pp = 0; pp < GSH_PARTS - 1; ++pp) {
        const float4 o4 = xb[pp * 128 + row];
        tot[0] += o4.x, tot[1] += o4.y, tot[2] += o4.z;
      }
      const int kout = kout_of(h), koff = koff_of(h);
      float* o = p.raw + (size_t)(m0 + row) * 9 + koff;
      for (int jj = 0; jj < kout; ++jj) {
        float val = tot[jj] + __ldg(p.b3 + koff + jj);
        if (h == 4) {  // mean: / grid size + reference point of the row's grid position (fea2gs.py:624-631)
          const int cell = (m0 + row) % (p.gh * p.gw), iy = cell / p.gw, ix = cell - iy * p.gw;
          const int nn = jj == 0 ? p.gw : p.gh, ii = jj == 0 ? ix : iy;
          // (python computes step / 2 and 1 - step / 2 in double before torch.linspace rounds them to fp32)
          const double hs = 0.5 / (double)nn;
          val = val / (float)nn + gsh_linspace((float)hs, (float)(1.0 - hs), nn, ii);
        }
        o[jj] = val;
      }
    };
    unsigned long long acc0, acc1, acc2;  // FP32x2: {sum over even channels, sum over odd channels}
    auto pk2 = [](float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; };
    auto upk2 = [](unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); };
    auto add2 = [](unsigned long long a, unsigned long long b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; };
    auto fma2 = [](unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; };
    // layer 2's chunks of one head: + b2, ReLU, layer 3 (KOUT dot products) on this warp's columns, two channels at a time
    auto chunks = [&](auto kout_c, int h) {
      constexpr int KOUT = decltype(kout_c)::value;
      const float4* tabh = tab + (h & 1) * GSH_HP;
#pragma unroll 1
      for (int c = 0; c < GSH_NCHUNK; ++c) {
        const int u = h * GSH_NCHUNK + c, buf = u & 1;
        gsu::mbar_wait(BAR(D2_FULL + buf), (uint32_t)(u >> 1) & 1u);
        gsu::fence_after_sync();
        if (et == 0) GSH_TRACE(h * 64 + 10 + 4 * c);
        const uint32_t d = tmem + lane_base + (buf ? GSH_D2_COL1 : GSH_D2_COL0);
#pragma unroll 1
        for (int c0 = part * (GSH_CHUNK / GSH_PARTS); c0 < (part + 1) * (GSH_CHUNK / GSH_PARTS); c0 += 32) {
          uint32_t v[32];
          gsu::tmem_ld_32x32(d + c0, v);
          gsu::tmem_ld_wait();
          const float4* tb = tabh + c * GSH_CHUNK + c0;  // (two float4 per channel pair = one per channel)
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float4 t0 = tb[j];  // broadcast LDS.128
            float xa, xb;
            upk2(add2(pk2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), pk2(t0.x, t0.y)), xa, xb);
            const unsigned long long x2 = pk2(fmaxf(xa, 0.f), fmaxf(xb, 0.f));
            acc0 = fma2(x2, pk2(t0.z, t0.w), acc0);
            if (KOUT > 1) {
              const float4 t1 = tb[j + 1];
              acc1 = fma2(x2, pk2(t1.x, t1.y), acc1);
              if (KOUT > 2) acc2 = fma2(x2, pk2(t1.z, t1.w), acc2);
            }
          }
        }
        gsu::fence_before_sync();
        gsu::mbar_arrive(BAR(D2_EMPTY + buf));
        if (et == 0) GSH_TRACE(h * 64 + 11 + 4 * c);
      }
    };
    for (int i = et; i < GSH_HP; i += GSH_EPI_THREADS) tab[i] = tab_entry(0, i);
    epi_sync();
    float prev0 = 0.f, prev1 = 0.f, prev2 = 0.f;
#pragma unroll 1
    for (int h = 0; h < GSH_HEADS; ++h) {
      gsu::mbar_wait(BAR(D1_FULL), (uint32_t)h & 1u);
      gsu::fence_after_sync();
      if (h > 0) gsu::mbar_wait(BAR(H1_FREE), (uint32_t)(h - 1) & 1u);
      if (et == 0) GSH_TRACE(h * 64 + 3);
      const float* b1 = sb1 + h * GSH_CP;
#pragma unroll 1
      for (int c0 = part * 16; c0 < GSH_CP; c0 += 16 * GSH_PARTS) {  // 16-column pieces: twelve over the parts, evenly
        uint32_t v[16];
        gsu::tmem_ld_32x16(tmem + lane_base + c0, v);
        gsu::tmem_ld_wait();
        const int kb = c0 >> 6, chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int g = 0; g < 2; ++g) {  // eight bf16 = one 16-byte chunk of the row
          const float4 ba = *reinterpret_cast<const float4*>(b1 + c0 + g * 8), bb = *reinterpret_cast<const float4*>(b1 + c0 + g * 8 + 4);
          const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = g * 8 + e * 2;
            const float x0 = fmaxf(__uint_as_float(v[j]) + bias[e * 2], 0.f);
            const float x1 = fmaxf(__uint_as_float(v[j + 1]) + bias[e * 2 + 1], 0.f);
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[e]) : "f"(x1), "f"(x0));
          }
          const uint32_t dst = hrow + (uint32_t)kb * GSH_XT + (uint32_t)(((chunk0 + g) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
      }
      gsu::fence_proxy_async();
      gsu::fence_before_sync();
      gsu::mbar_arrive(BAR(H1_READY));
      if (et == 0) GSH_TRACE(h * 64 + 4);
      // -- off the critical path from here (the tensor cores are busy with this head's layer 2) --
      // the previous head's outputs (its exchange buffer is complete: the barrier that ended that head)
      // (one buffer is enough: a writer reaches the end of this head's chunks only after every thread has arrived at
      // the barrier of chunk 3, i.e. long after this read)
      if (h > 0 && part == 0 && live) write_out(h - 1, prev0, prev1, prev2, exch);
      // the next head's table: requested now, stored after this head's chunks
      float4 nt[2];
      const int ni0 = et, ni1 = et + GSH_EPI_THREADS;
      if (h + 1 < GSH_HEADS) {
        nt[0] = tab_entry(h + 1, ni0);
        if (ni1 < GSH_HP) nt[1] = tab_entry(h + 1, ni1);
      }
      acc0 = acc1 = acc2 = 0ull;
      const int kout = kout_of(h);
      if (kout == 1) chunks(std::integral_constant<int, 1>{}, h);
      else if (kout == 2) chunks(std::integral_constant<int, 2>{}, h);
      else chunks(std::integral_constant<int, 3>{}, h);
      if (h + 1 < GSH_HEADS) {  // (the other table buffer was last read a head ago, behind the previous barrier)
        float4* tn = tab + ((h + 1) & 1) * GSH_HP;
        tn[ni0] = nt[0];
        if (ni1 < GSH_HP) tn[ni1] = nt[1];
      }
      {
        float e0, o0, e1, o1, e2, o2;
        upk2(acc0, e0, o0), upk2(acc1, e1, o1), upk2(acc2, e2, o2);
        prev0 = e0 + o0, prev1 = e1 + o1, prev2 = e2 + o2;
      }
      if (part > 0) exch[(part - 1) * 128 + row] = make_float4(prev0, prev1, prev2, 0.f);
      epi_sync();  // exchange buffer and next table complete; every reader of this head's table is done
    }
    if (part == 0 && live) write_out(GSH_HEADS - 1, prev0, prev1, prev2, exch);
  }
  gsu::fence_before_sync();
  if (GSH_CLUSTER > 1) gsu::cluster_sync();  // no CTA leaves while a peer may still multicast into it or signal its barriers
  else __syncthreads();
  if (warp == 1) gsu::tmem_dealloc<512>(tmem);
}

static unsigned long long* gsh_trace_buffer = nullptr;  // GSH_CFG_TRACE builds: set by gsr_head_tail_set_trace
#if GSH_CFG_TRACE
extern "C" void gsr_head_tail_set_trace(unsigned long long* device_buffer_320) { gsh_trace_buffer = device_buffer_320; }
#endif
extern "C" int gsr_head_tail_forward(const void* x_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16,
                                     const float* b2, const float* w3, const float* b3, float* raw, int m, int grid_h,
                                     int grid_w, void* stream) {
  if (!x_bf16 || !w1_bf16 || !b1 || !w2_bf16 || !b2 || !w3 || !b3 || !raw) return GSR_ERR_NULL_POINTER;
  if (m < 0 || grid_h < 1 || grid_w < 1) return GSR_ERR_BAD_SHAPE;
  if (m == 0) return GSR_OK;
  CUtensorMap tx, t1, t2;
  int rc = gsh_map_bf16(&tx, x_bf16, (uint64_t)m, GSH_CP, 128);
  if (rc) return rc;
  rc = gsh_map_bf16(&t1, w1_bf16, (uint64_t)GSH_HEADS * GSH_CP, GSH_CP, GSH_CP / GSH_CLUSTER);
  if (rc) return rc;
  rc = gsh_map_bf16(&t2, w2_bf16, (uint64_t)GSH_HEADS * GSH_HP, GSH_CP, GSH_CHUNK / GSH_CLUSTER);
  if (rc) return rc;
  GshArgs a;
  a.b1 = b1, a.b2 = b2, a.w3 = w3, a.b3 = b3, a.raw = raw, a.m = m, a.gh = grid_h, a.gw = grid_w;
  a.trace = gsh_trace_buffer;
  static std::atomic<int> optin[64];
  int dev = 0;
  GSH_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !optin[dev].load(std::memory_order_relaxed)) {
    GSH_CUDA(cudaFuncSetAttribute(gsh_head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GSH_SMEM + 1024));
    if (dev >= 0 && dev < 64) optin[dev].store(1, std::memory_order_relaxed);
  }
  const int tiles = (m + 127) / 128, grid = (tiles + GSH_CLUSTER - 1) / GSH_CLUSTER * GSH_CLUSTER;  // whole clusters
  gsh_head_tail_kernel<<<grid, GSH_THREADS, GSH_SMEM + 1024, (cudaStream_t)stream>>>(tx, t1, t2, a);
  GSH_CUDA(cudaGetLastError());
  return GSR_OK;
}

extern "C" int gsr_test_umma_gemm(const void* a_bf16, const void* b_bf16, float* c, int m, int n, int k, void* stream) {
  if (!a_bf16 || !b_bf16 || !c) return GSR_ERR_NULL_POINTER;
  if (m <= 0 || m % 128 || k != 192 || (n != 192 && n != 128)) return GSR_ERR_BAD_SHAPE;
  CUtensorMap ta, tb;
  int rc = gsh_map_bf16(&ta, a_bf16, (uint64_t)m, (uint64_t)k, 128);
  if (rc) return rc;
  rc = gsh_map_bf16(&tb, b_bf16, (uint64_t)n, (uint64_t)k, (uint32_t)n);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 192) {
    constexpr int SM = 3 * (128 * 128 + 192 * 128) + 64;
    GSH_CUDA(cudaFuncSetAttribute(gsh_gemm_test_kernel<192, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM + 1024));
    gsh_gemm_test_kernel<192, 3><<<m / 128, GSH_T_THREADS, SM + 1024, st>>>(ta, tb, c, n);
  } else {
    constexpr int SM = 3 * (128 * 128 + 128 * 128) + 64;
    GSH_CUDA(cudaFuncSetAttribute(gsh_gemm_test_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM + 1024));
    gsh_gemm_test_kernel<128, 3><<<m / 128, GSH_T_THREADS, SM + 1024, st>>>(ta, tb, c, n);
  }
  GSH_CUDA(cudaGetLastError());
  return GSR_OK;
}
