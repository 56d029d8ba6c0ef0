/*
 * gsraster_test.h -- CPU test hooks exported by libgsraster.so.  They run the library's shared
 * host/device culling code (gsasr_b200/csrc/gsr_common.cuh) on the HOST so that the CPU test
 * suite can check it against the oracle without a GPU.  Not part of the product API.
 */
#ifndef GSRASTER_TEST_H_
#define GSRASTER_TEST_H_
#ifdef __cplusplus
extern "C" {
#endif

/* out is (s,9) int32: live, x0, x1, y0, y1 (inclusive cull box), binds, large, home bin, extent */
void gsr_host_setup(const float* sigmas, const float* coords, const float* colors, int s, int h,
                    int w, float dmax, float ksigma, int* out);
/* Row-band view (gsr_forward_band): out is s x 6 = live, x0, x1, y0, y1 (band-local rows), binds. */
void gsr_host_setup_band(const float* sigmas, const float* coords, const float* colors, int s, int h, int w,
                         int row0, int rows, float dmax, float ksigma, int* out);
/* inclusive pixel range of the reference's dmax window on an n-pixel axis (gs.cu:39-50) */
void gsr_host_window_range(int n, float ctr, float dmax, int* lo, int* hi);
/* 16-bit mask of the 8x8 regions of the 32x32 tile at (tx0,ty0) that Gaussian i may touch */
unsigned gsr_host_region_mask(const float* sigmas, const float* coords, const float* colors, int i,
                              int h, int w, float dmax, float ksigma, int tx0, int ty0);
/* bucket entries of Gaussian i (forward fast path): out is n x 3 = region column (16 px), region row (8 px),
 * 8-bit mask of the region's 4x4-pixel cells (bit = cell row * 4 + cell column); returns n (may exceed cap) */
int gsr_host_entries(const float* sigmas, const float* coords, const float* colors, int i, int h, int w,
                     float dmax, float ksigma, int* out, int cap);
void gsr_host_geometry(int* tile_w, int* tile_h, int* bin, int* region, int* large_px);
/* unit test of the tcgen05 / TMEM / TMA building block of the head-tail kernel (needs a GPU): c (m x n fp32) =
 * a (m x 192 bf16) * b (n x 192 bf16)^T, m % 128 == 0, n = 128 or 192, one 128-row tile per CTA */
int gsr_test_umma_gemm(const void* a_bf16, const void* b_bf16, float* c, int m, int n, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif
