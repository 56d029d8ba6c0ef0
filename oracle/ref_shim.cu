// ref_shim.cu -- extern "C" doorway into the UNMODIFIED reference kernels.  TEST INFRASTRUCTURE.
//
// Compiled together with /root/reference/utils/gs_cuda_dmax/gs.cu (-DGSREF_DMAX) or
// /root/reference/utils/gs_cuda/gs.cu by oracle/Makefile into oracle/_ref/; the reference's own
// host launchers (_gs_render, _gs_render_backward; gs.h) are called as they are, on the legacy
// default stream, exactly like the reference's pybind wrapper does (gswrapper.cpp:29-34,64-72).
// Nothing from the reference is copied into this repository: gs.h is included from where it lies.
#include "gs.h"
#include <cuda_runtime.h>

extern "C" {
#ifdef GSREF_DMAX
__attribute__((visibility("default"))) int gsref_forward(const float* sigmas, const float* coords,
                                                         const float* colors, float* img, int s,
                                                         int h, int w, int c, float dmax) {
  _gs_render(sigmas, coords, colors, img, s, h, w, c, dmax);
  return (int)cudaGetLastError();
}
__attribute__((visibility("default"))) int gsref_backward(const float* sigmas, const float* coords,
                                                          const float* colors, const float* grads,
                                                          float* gs, float* gc, float* gk, int s,
                                                          int h, int w, int c, float dmax) {
  _gs_render_backward(sigmas, coords, colors, grads, gs, gc, gk, s, h, w, c, dmax);
  return (int)cudaGetLastError();
}
#else
__attribute__((visibility("default"))) int gsref_forward(const float* sigmas, const float* coords,
                                                         const float* colors, float* img, int s,
                                                         int h, int w, int c, float /*dmax*/) {
  _gs_render(sigmas, coords, colors, img, s, h, w, c);
  return (int)cudaGetLastError();
}
__attribute__((visibility("default"))) int gsref_backward(const float* sigmas, const float* coords,
                                                          const float* colors, const float* grads,
                                                          float* gs, float* gc, float* gk, int s,
                                                          int h, int w, int c, float /*dmax*/) {
  _gs_render_backward(sigmas, coords, colors, grads, gs, gc, gk, s, h, w, c);
  return (int)cudaGetLastError();
}
#endif
__attribute__((visibility("default"))) int gsref_sync(void) { return (int)cudaDeviceSynchronize(); }
}
