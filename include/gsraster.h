/*
 * gsraster.h -- C ABI of the B200-native 2-D Gaussian rasteriser (libgsraster.so).
 *
 * This is the drop-in boundary for GSASR's render path.  Every entry point below
 * replaces one piece of the reference's native interface (paths relative to the
 * GSASR tree):
 *
 *   gsr_forward           <-  _gs_render            utils/gs_cuda_dmax/gs.h:2-12, gs.cu:67-83
 *                             gs_render (pybind)    utils/gs_cuda_dmax/gswrapper.cpp:9-35
 *                             (and the window-less  utils/gs_cuda/gs.h, gs.cu:64-79 with dmax=+inf)
 *   gsr_backward          <-  _gs_render_backward   utils/gs_cuda_dmax/gs.h:14-27, gs.cu:167-186
 *                             gs_render_backward    utils/gs_cuda_dmax/gswrapper.cpp:37-73
 *   gsr_forward_batch /   <-  the per-sample Python loop that calls the two functions above B times
 *   gsr_backward_batch        TrainTestGSASR/basicsr/models/gsasr_model.py:191-233
 *   gsr_frontend_forward  <-  activations + unit mapping + render + HWC->CHW transpose
 *                             utils/gaussian_splatting.py:119-131,158-217
 *   gsr_frontend_backward <-  autograd of the same chain down to the raw (N,9) head output
 *   gsr_*_batch_uniform   <-  the same per-sample loop for a batch of ONE shape, in one launch each way
 *   gsr_forward_band /    <-  (new) rows [row0, row0+rows) of one image: a single large render split
 *   gsr_backward_band         over several GPUs
 *   gsr_forward_window    <-  (new) the tile buffer + paste pass of utils/split_and_joint_image.py:160-227:
 *                             the kernel writes a tile's pixels straight into the (possibly remote) canvas
 *   GSR_FLAG_U8           <-  the clamp / x255 / round / uint8 post-processing of inference_paper.py:136-138
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named *_host;
 *   - fp32, contiguous, layouts as in the reference: sigmas (s,3) = (sigma_x, sigma_y, rho),
 *     coords (s,2) = (x, y) on the align_corners=True [-1,1] grid, colors (s,3),
 *     image / grads (h,w,3) HWC;
 *   - the library never allocates or frees device memory, never synchronises the device and
 *     never throws: scratch comes from the caller (`workspace`, sized by gsr_workspace_bytes),
 *     work is enqueued on `stream` (a cudaStream_t / CUstream; NULL = legacy default stream,
 *     which is what the reference launches on), and the return value is a gsr_status;
 *   - re-entrant: no global mutable state (bar a write-once per-device cache of launch geometry);
 *     concurrent calls need distinct workspaces.
 *
 * Semantics that are identical to the reference:
 *   - pixel (wi,hi) of Gaussian g is evaluated iff  |px(wi)-x| <= dmax  and  |py(hi)-y| <= dmax
 *     with px(i) = (float)(2.0*i/(n-1) - 1.0) and the subtraction in fp32 (gs.cu:39-50,124-132):
 *     the inclusion set is bit-identical;
 *   - forward ACCUMULATES into `img` (gs.cu:58-60), backward ACCUMULATES into the three
 *     gradient arrays (gs.cu:152-159) -- callers pre-zero them (gswrapper.py:40-43);
 *   - no normalisation constant, no alpha compositing: img += sum_g color_g * exp(-q_g/2).
 *
 * Semantics that are new (documented deviations):
 *   - `ksigma`: contributions with Mahalanobis distance^2 > ksigma^2 are dropped (each is
 *     < exp(-ksigma^2/2) * |color|).  ksigma <= 0 selects GSR_DEFAULT_KSIGMA; ksigma = +inf
 *     selects the exact mode, which drops only terms that are exactly 0 in flushed fp32
 *     (ksigma is capped at GSR_EXACT_KSIGMA because exp2 of anything below -126 flushes to 0);
 *   - Gaussians with non-finite parameters, sigma == 0 or |rho| >= 1 are skipped (the reference
 *     renders NaN/Inf for them);
 *   - c must be 3 (the reference's forward hard-codes 3 channels, gs.cu:29-31,58-60);
 *     h, w must be in [2, 32767] (the reference divides by (n-1)).
 */
#ifndef GSRASTER_H_
#define GSRASTER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_VERSION 100

#define GSR_DEFAULT_KSIGMA 5.0f
#define GSR_EXACT_KSIGMA 13.25f

typedef enum gsr_status {
  GSR_OK = 0,
  GSR_ERR_NULL_POINTER = 1,
  GSR_ERR_BAD_SHAPE = 2,       /* s < 0, h/w outside [2,32767] */
  GSR_ERR_BAD_CHANNELS = 3,    /* c != 3 */
  GSR_ERR_WORKSPACE = 4,       /* workspace NULL, misaligned (256 B) or too small */
  GSR_ERR_BAD_ARGUMENT = 5,
  GSR_ERR_CUDA = 6             /* a CUDA runtime call failed; see gsr_last_cuda_error() */
} gsr_status;

/* flags */
#define GSR_FLAG_OVERWRITE 0x1u /* forward: img = sum instead of img += sum (skips the read) */
#define GSR_FLAG_CHW 0x2u       /* forward: img is (3,h,w); backward: grads is (3,h,w)       */
/* forward only, with GSR_FLAG_OVERWRITE, (h,w,3) layout: `img` is UINT8 -- the kernel writes
 * round(clamp(value, 0, 1) * 255) (round-half-even, as numpy), i.e. the post-processing of
 * inference_paper.py:136-138 fused into the write-out: no fp32 image, a quarter of the bytes. */
#define GSR_FLAG_U8 0x4u
#define GSR_FLAG_BGR 0x8u       /* with GSR_FLAG_U8: channel order b, g, r (cv2.imwrite's; :137) */
/* forward, with GSR_FLAG_OVERWRITE, (h,w,3) fp32 layout, w % 4 == 0, img 16-byte aligned: finished 16x8-pixel
 * regions leave the SM as 128-bit stores of whole region rows (192 contiguous bytes) staged through shared
 * memory, instead of 4-byte stores.  For an image in ANOTHER GPU's memory (peer mapping over NVLink): 16-byte
 * write packets.  Same pixels either way; ~1 % slower on a local image, hence opt-in.  Ignored where it does not
 * apply.  (The uint8 image of GSR_FLAG_U8 always leaves that way when w % 16 == 0.) */
#define GSR_FLAG_ROW_STORES 0x10u
/* bit-reproducible output.  Forward: the lists a region's pixels are summed from are filled with atomics, so the
 * fp32 summation order -- like that of the reference's atomicAdd (gs.cu:58-60) -- differs from run to run in the
 * last bits.  With this flag every list is sorted by Gaussian index first (HL: +30 % of the step): the image is
 * then a function of the inputs alone.  Holds while the region lists fit their buckets (no fallback to the
 * home-bin kernel) and no bucket exceeds 8192 entries.  Backward: the default kernel walks the same region buckets
 * and its partial sums meet in atomics (like the reference's, gs.cu:163-174); with this flag the Gaussian-centric
 * kernel runs instead (one warp owns a Gaussian, fixed sweep order, one writer per output: always reproducible,
 * 1.9x slower at the headline shape). */
#define GSR_FLAG_DETERMINISTIC 0x20u

int gsr_version(void);
const char* gsr_status_string(int status);
/* Last CUDA error code observed by the calling thread inside this library (0 = none). */
int gsr_last_cuda_error(void);

/* Scratch bytes needed by gsr_forward / gsr_backward for s Gaussians on an h x w image.
 * Depends on sizes only (never on data), so it can be called once and cached. 0 on bad sizes. */
size_t gsr_workspace_bytes(int s, int h, int w);

int gsr_forward(const float* sigmas, const float* coords, const float* colors, float* img,
                int s, int h, int w, int c, float dmax, float ksigma, uint32_t flags,
                void* workspace, size_t workspace_bytes, void* stream);

/* Gradients are ACCUMULATED into grads_* (the reference's atomicAdd contract, gs.cu:163-174).  flags:
 * GSR_FLAG_CHW (grads is (3,h,w)); GSR_FLAG_DETERMINISTIC (the Gaussian-centric kernel: bit-reproducible, slower).
 * Default: the backward over the region buckets -- its partial sums meet in fp32 REDs, so the last bits vary run
 * to run, as the reference's do. */
int gsr_backward(const float* sigmas, const float* coords, const float* colors,
                 const float* grads, float* grads_sigmas, float* grads_coords,
                 float* grads_colors, int s, int h, int w, int c, float dmax, float ksigma,
                 uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);

/* Row band of one image: rows [row0, row0 + rows) of the h x w image, for splitting a single large
 * image over several GPUs (the reference has no counterpart: its kernels always walk
 * the whole image, gs.cu:38-62).  img_band / grads_band hold `rows` rows ((rows,w,3), or (3,rows,w)
 * with GSR_FLAG_CHW); pixel coordinates, dmax windows and inclusion sets are those of the FULL image,
 * so the bands of a partition of [0,h) stacked together equal gsr_forward on the whole image (up to
 * summation order), and the gradients of the bands SUM to gsr_backward's (the gradients are linear
 * in the pixels).  rows >= 2.  Workspace: gsr_workspace_bytes(s, rows, w). */
int gsr_forward_band(const float* sigmas, const float* coords, const float* colors, float* img_band,
                     int s, int h, int w, int c, int row0, int rows, float dmax, float ksigma,
                     uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);
int gsr_backward_band(const float* sigmas, const float* coords, const float* colors,
                      const float* grads_band, float* grads_sigmas, float* grads_coords,
                      float* grads_colors, int s, int h, int w, int c, int row0, int rows, float dmax,
                      float ksigma, uint32_t flags, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Render into a WINDOW of a larger destination (tiled inference: utils/split_and_joint_image.py:160-227
 * renders every tile into its own buffer and pastes it into the canvas afterwards; here the raster
 * kernel writes the tile's pixels straight into the canvas -- which may live on another GPU: peer
 * stores over NVLink).  Pixel (y, x), channel ch of the h x w render is written to
 *     origin + y * row_stride + x * pix_stride + ch * chan_stride            (strides in floats)
 * if nclip == 0 or (x, y) lies inside one of the clip rectangles {x0, y0, x1, y1} (inclusive, render
 * coordinates); other pixels are neither read nor written, so `origin` may point outside the
 * destination as long as every clipped-in pixel is inside.  GSR_FLAG_OVERWRITE or accumulate;
 * GSR_FLAG_CHW is expressed through the strides instead.  `win` is HOST memory. */
#define GSR_MAX_CLIP 8
typedef struct gsr_window {
  long long row_stride, pix_stride, chan_stride;
  int nclip;
  int clip[GSR_MAX_CLIP][4];
} gsr_window;
int gsr_forward_window(const float* sigmas, const float* coords, const float* colors, float* origin,
                       const gsr_window* win, int s, int h, int w, int c, float dmax, float ksigma,
                       uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);

/* Split-phase form.  gsr_forward == gsr_prepare + gsr_forward_prepared, gsr_backward ==
 * gsr_prepare + gsr_backward_prepared.  gsr_prepare runs the O(N) set-up (cull boxes, counting
 * sort by home bin) and leaves it in `workspace`; as long as the workspace is untouched and
 * (s, h, w, dmax, ksigma) are the same, any number of raster passes can reuse it -- a training
 * step prepares once and runs forward and backward (GSCUDA.forward / .backward,
 * gswrapper.py:25-44, see the same Gaussians).  A whole-image gsr_forward call leaves the same state behind
 * as far as gsr_backward_prepared without GSR_FLAG_DETERMINISTIC needs it (region buckets; the home-bin arrays
 * if a bucket overflowed): gsr_forward + gsr_backward_prepared on the same workspace is how this repo's autograd
 * boundaries avoid the second set-up (gsasr_b200/gswrapper.py). */
int gsr_prepare(const float* sigmas, const float* coords, const float* colors, int s, int h, int w,
                float dmax, float ksigma, void* workspace, size_t workspace_bytes, void* stream);
int gsr_forward_prepared(float* img, int s, int h, int w, float ksigma, uint32_t flags,
                         void* workspace, size_t workspace_bytes, void* stream);
int gsr_backward_prepared(const float* sigmas, const float* grads, float* grads_sigmas,
                          float* grads_coords, float* grads_colors, int s, int h, int w,
                          uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);

/* One sample of a ragged batch (heterogeneous h, w, dmax -- the training loop's shape,
 * gsasr_model.py:191-233).  Pointers are device pointers; the descriptor array is HOST memory. */
typedef struct gsr_sample {
  const float* sigmas;
  const float* coords;
  const float* colors;
  float* img;                 /* forward: output;  backward: unused (may be NULL)    */
  const float* grads;         /* backward: dL/dimg; forward: unused (may be NULL)    */
  float* grads_sigmas;        /* backward outputs; forward: unused                   */
  float* grads_coords;
  float* grads_colors;
  int s, h, w;
  float dmax;
} gsr_sample;

size_t gsr_workspace_bytes_batch(const gsr_sample* samples_host, int n);
int gsr_forward_batch(const gsr_sample* samples_host, int n, float ksigma, uint32_t flags,
                      void* workspace, size_t workspace_bytes, void* stream);
int gsr_backward_batch(const gsr_sample* samples_host, int n, float ksigma, uint32_t flags,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Uniform batch: `batch` samples of the same shape, s_per Gaussians each, contiguous:
 * sigmas (batch*s_per,3), coords (batch*s_per,2), colors (batch*s_per,3), imgs / grads (batch,h,w,3).
 * The training loop's per-sample render (gsasr_model.py:191-233) for a batch whose samples share
 * (h, w, dmax) in ONE set-up and ONE raster launch: when h is a multiple of 8 the samples are stacked
 * into a (batch*h, w) image (cut to 32767 rows per launch); otherwise, and with GSR_FLAG_CHW, one
 * call per sample.  Results equal `batch` calls of gsr_forward / gsr_backward (up to summation order).
 * Workspace: gsr_workspace_bytes_batch_uniform(batch, s_per, h, w). */
size_t gsr_workspace_bytes_batch_uniform(int batch, int s_per, int h, int w);
int gsr_forward_batch_uniform(const float* sigmas, const float* coords, const float* colors, float* imgs,
                              int batch, int s_per, int h, int w, int c, float dmax, float ksigma,
                              uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);
int gsr_backward_batch_uniform(const float* sigmas, const float* coords, const float* colors,
                               const float* grads, float* grads_sigmas, float* grads_coords,
                               float* grads_colors, int batch, int s_per, int h, int w, int c, float dmax,
                               float ksigma, uint32_t flags, void* workspace, size_t workspace_bytes,
                               void* stream);

/* Padded (ragged) batch: `batch` samples of s_per Gaussians each, rendered at their OWN sizes
 * hw_host[2b] x hw_host[2b+1] (HOST array; 2 <= h_b <= hmax, 2 <= w_b <= wmax) and windows dmax_host[b]
 * (HOST array, or NULL: `dmax` for all) into the top-left corner of their slot of the padded buffer
 * imgs / grads (batch, hmax, wmax, 3) -- the training loop's per-sample render + F.pad to the largest
 * size (gsasr_model.py:191-233) in one set-up and one raster launch each way.  hmax % 8 == 0.  Pixels of a
 * slot outside its sample's image are written 0 (forward) / ignored (backward).  The dmax inclusion set of
 * every sample is exact; values differ from `batch` separate calls by fp32 round-off only (~1e-6: the
 * records are rescaled to the padded image's coordinate normalisation). */
size_t gsr_workspace_bytes_batch_padded(int batch, int s_per, int hmax, int wmax);
int gsr_forward_batch_padded(const float* sigmas, const float* coords, const float* colors, float* imgs,
                             int batch, int s_per, int hmax, int wmax, const int* hw_host,
                             const float* dmax_host, float dmax, float ksigma, uint32_t flags,
                             void* workspace, size_t workspace_bytes, void* stream);
int gsr_backward_batch_padded(const float* sigmas, const float* coords, const float* colors,
                              const float* grads, float* grads_sigmas, float* grads_coords,
                              float* grads_colors, int batch, int s_per, int hmax, int wmax,
                              const int* hw_host, const float* dmax_host, float dmax, float ksigma,
                              uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);

/* Fused front end: raw head output (s,9) = (sx, sy, rho, alpha, r, g, b, mu_x, mu_y) ->
 * activations (gaussian_splatting.py:174-180) -> unit/coordinate mapping (:121-123) ->
 * render -> (3,h,w) image, written (not accumulated).  step_size = default_step_size / scale.
 * The mapped (sigmas, coords, colors) are also written to `mapped` (s*8 floats, layout
 * [sigmas (s,3) | coords (s,2) | colors (s,3)]) because the backward needs them. */
int gsr_frontend_forward(const float* raw_params, float* mapped, float* img_chw, int s, int h,
                         int w, float step_size, float dmax, float ksigma, void* workspace,
                         size_t workspace_bytes, void* stream);
/* grads_chw (3,h,w) -> grad_raw (s,9), written (not accumulated). */
int gsr_frontend_backward(const float* raw_params, const float* mapped, const float* grads_chw,
                          float* grad_raw, int s, int h, int w, float step_size, float dmax,
                          float ksigma, void* workspace, size_t workspace_bytes, void* stream);

/* The fused front end for a uniform batch: raw (batch*s_per,9) -> imgs (batch,h,w,3) (HWC per sample,
 * i.e. a (batch,3,h,w) tensor in channels-last layout), and its backward from grads (batch,h,w,3).
 * mapped: batch*s_per*8 floats.  Backward workspace: gsr_workspace_bytes_batch_uniform + 32 bytes per
 * Gaussian rounded up to 256. */
int gsr_frontend_forward_batch_uniform(const float* raw_params, float* mapped, float* imgs, int batch,
                                       int s_per, int h, int w, float step_size, float dmax, float ksigma,
                                       void* workspace, size_t workspace_bytes, void* stream);
int gsr_frontend_backward_batch_uniform(const float* raw_params, const float* mapped, const float* grads,
                                        float* grad_raw, int batch, int s_per, int h, int w,
                                        float step_size, float dmax, float ksigma, void* workspace,
                                        size_t workspace_bytes, void* stream);

/* The fused front end for a padded batch (see gsr_forward_batch_padded): every sample its own size
 * hw_host[b] and step size step_host[b] = default_step_size / scale_b (HOST arrays).  mapped:
 * batch*s_per*8 floats.  Backward workspace: gsr_workspace_bytes_batch_padded + 32 bytes per Gaussian of
 * the batch, rounded up to 256. */
int gsr_frontend_forward_batch_padded(const float* raw_params, float* mapped, float* imgs, int batch,
                                      int s_per, int hmax, int wmax, const int* hw_host,
                                      const float* step_host, const float* dmax_host, float dmax,
                                      float ksigma, void* workspace, size_t workspace_bytes, void* stream);
int gsr_frontend_backward_batch_padded(const float* raw_params, const float* mapped, const float* grads,
                                       float* grad_raw, int batch, int s_per, int hmax, int wmax,
                                       const int* hw_host, const float* step_host, const float* dmax_host,
                                       float dmax, float ksigma, void* workspace, size_t workspace_bytes,
                                       void* stream);

/* gsr_frontend_forward into a window (see gsr_forward_window): raw head output -> destination pixels. */
int gsr_frontend_forward_window(const float* raw_params, float* mapped, float* origin, const gsr_window* win,
                                int s, int h, int w, float step_size, float dmax, float ksigma,
                                uint32_t flags, void* workspace, size_t workspace_bytes, void* stream);

/* The training loop's pixel loss, fused with its crop and its gradient (gsasr_model.py:212-234: F.pad to the
 * batch's largest size, crop output and ground truth back to (h_b, w_b), L1Loss(reduction='mean') per sample,
 * divided by the batch size):
 *     loss = weight * sum_b 1 / (3 h_b w_b) * sum_{c, y < h_b, x < w_b} |sr[b,c,y,x] - gt[b,c,y,x]|
 * (the caller passes weight = loss_weight / batch size).  sr and gt are addressed as element (b,c,y,x) at
 * b*n + c*c + y*h + x*w of their stride arrays {n, c, h, w} (in floats) -- any layout, e.g. the channels-last
 * (batch,hmax,wmax,3) image of gsr_forward_batch_padded and a (batch,3,H,W) ground truth.  Writes *loss (device,
 * one float; accumulate != 0: adds to it) and grad = dloss/dsr in sr's layout over the whole (batch,3,hmax,wmax)
 * block, zero outside the crops -- the `grads` array of gsr_backward_batch_padded.  hw_host: batch x (h_b, w_b),
 * HOST memory, 1 <= h_b <= hmax, 1 <= w_b <= wmax.  workspace: gsr_l1_crop_workspace_bytes() bytes, 256-byte
 * aligned.  The sum is reduced in a fixed order: bit-reproducible.  No counterpart in the reference's extension
 * (basicsr/losses L1Loss + slicing: four elementwise passes and a reduction per sample). */
size_t gsr_l1_crop_workspace_bytes(void);
int gsr_l1_crop_loss(const float* sr, const long long* sr_strides, const float* gt, const long long* gt_strides,
                     float* grad, float* loss, int batch, int hmax, int wmax, const int* hw_host, float weight,
                     int accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Head-tail fusion (utils/fea2gs.py:496-541 and :611-633; the same code in utils/fea2gsropeamp.py:571-719): the five
 * per-Gaussian MLPs  Linear(C,C) - ReLU - Linear(C,4C) - ReLU - Linear(4C,k),  k = 2 (sigma), 1 (rho), 1 (alpha),
 * 3 (rgb), 2 (mean), on the (m, C) feature rows after UPNet, the normalisation of the means by the grid size, the
 * reference points and the concatenation -- ONE kernel on the tcgen05 tensor cores (bf16 operands, fp32 accumulation
 * in tensor memory; the hidden layers never reach HBM).  Forward only (inference).
 *   x_bf16   (m, 192) bf16: the feature rows, channels zero-padded to 192 (C = 180 or 192); row = (sample, iy, ix)
 *   w1_bf16  (5*192, 192) bf16: first Linear of the five MLPs, [out, in] as nn.Linear stores it, zero-padded
 *   b1       (5, 192) fp32
 *   w2_bf16  (5*768, 192) bf16: second Linear, [out, in], zero-padded (4C = 720 or 768)
 *   b2       (5, 768) fp32
 *   w3       (9, 768) fp32: last Linear, rows in output order sigma_x, sigma_y, rho, alpha, r, g, b, mean_x, mean_y
 *   b3       (9) fp32
 *   raw      (m, 9) fp32, written: what torch.cat([...], dim=-1) returns at fea2gs.py:632 -- the input of
 *            generate_2D_gaussian_splatting_step / gsr_frontend_forward
 *   grid_h, grid_w: the Gaussian grid of one sample (m = samples * grid_h * grid_w).
 * Pointers 16-byte aligned.  Arithmetic: the reference under bf16 autocast with the last Linear kept in fp32. */
int gsr_head_tail_forward(const void* x_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16,
                          const float* b2, const float* w3, const float* b3, float* raw, int m, int grid_h,
                          int grid_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSRASTER_H_ */
