"""Drop-in for the reference's pybind module ``gscuda`` (built by setup_gscuda.py:6-21 from
utils/gs_cuda_dmax/{gswrapper.cpp,gs.cu}).

Same two functions, same positional signatures and in-place accumulate semantics:

    gs_render(sigmas, coords, colors, rendered_img, s, h, w, c, dmax)            gswrapper.cpp:9-35
    gs_render_backward(sigmas, coords, colors, grads, grads_sigmas, grads_coords,
                       grads_colors, s, h, w, c, dmax)                           gswrapper.cpp:37-73

The window-less forms of utils/gs_cuda/gswrapper.cpp:9-34,36-71 (no ``dmax`` argument) are
accepted too and mean ``dmax = +inf``.  Like the reference, tensors must be CUDA and contiguous
(``RuntimeError`` otherwise, same wording as TORCH_CHECK); unlike it, dtype, shapes and c == 3 are
validated as well, and the kernels run on the CURRENT torch stream instead of the legacy default
stream.  Calls are asynchronous.  There is no CPU fallback.
"""
from __future__ import annotations

import os

import torch

from . import _lib

__all__ = ["gs_render", "gs_render_backward", "gs_render_band", "gs_render_backward_band", "gs_render_batch",
           "gs_render_backward_batch", "gs_render_batch_padded", "gs_render_backward_batch_padded", "gs_render_window", "frontend_render_window", "gs_render_u8", "set_ksigma", "get_ksigma", "set_deterministic", "get_deterministic", "set_reuse_setup", "get_reuse_setup", "gs_render_backward_prepared"]

_ksigma = float(os.environ.get("GSR_KSIGMA", "0"))  # 0 -> library default (GSR_DEFAULT_KSIGMA)


def set_ksigma(k: float) -> None:
    """Truncation radius in sigmas: contributions with Mahalanobis distance > k are dropped
    (each < exp(-k^2/2)*|colour|).  0 = library default (5); float('inf') = exact mode."""
    global _ksigma
    _ksigma = float(k)


def get_ksigma() -> float:
    return _ksigma


_deterministic = os.environ.get("GSR_DETERMINISTIC", "0") not in ("", "0")


def set_deterministic(on: bool) -> None:
    """Bit-reproducible renders and gradients (GSR_FLAG_DETERMINISTIC).  Forward: every region list is sorted by
    Gaussian index before it is rasterised (~30 % slower).  Backward: the Gaussian-centric kernel (one warp owns a
    Gaussian, fixed sweep order, one writer per output; ~1.9x slower at the headline shape) instead of the default
    backward over the region buckets, whose partial sums meet in atomics.  The reference's atomicAdd accumulation
    (gs.cu:58-60, :163-174) is not reproducible run to run; neither are this library's defaults (in the last bits).
    Also settable with GSR_DETERMINISTIC=1 in the environment."""
    global _deterministic
    _deterministic = bool(on)


def get_deterministic() -> bool:
    return _deterministic


_reuse_setup = os.environ.get("GSR_REUSE_SETUP", "1") not in ("", "0")


def set_reuse_setup(on: bool) -> None:
    """The autograd boundaries (gswrapper.GSCUDA, gaussian_splatting.render_chw) keep the forward call's workspace --
    region buckets, records, coordinate tables -- alive until the backward and run gsr_backward_prepared on it
    instead of setting the same Gaussians up a second time (HL: 0.80 -> 0.66 ms per backward).  Costs the workspace's
    memory (gsr_workspace_bytes) between the two calls; GSR_REUSE_SETUP=0 or set_reuse_setup(False) turns it off."""
    global _reuse_setup
    _reuse_setup = bool(on)


def get_reuse_setup() -> bool:
    return _reuse_setup and not _deterministic  # (the deterministic backward needs the home-bin sort: full call)


def _fwd_flags(flags) -> int:
    return int(flags) | (_lib.GSR_FLAG_DETERMINISTIC if _deterministic else 0)


def _check_input(t, name: str) -> None:
    # wording follows CHECK_CUDA / CHECK_CONTIGUOUS of gswrapper.cpp:5-7
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")


def _check_shape(t, shape, name: str) -> None:
    if tuple(t.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")


def workspace(s: int, h: int, w: int, device) -> torch.Tensor:
    """Scratch for one call, from torch's caching allocator (stream-ordered reuse)."""
    n = _lib.load().gsr_workspace_bytes(int(s), int(h), int(w))
    if n == 0:
        raise RuntimeError(f"libgsraster: bad sizes s={s}, h={h}, w={w} (need s>=0, 2<=h,w<=32767)")
    return torch.empty(n, dtype=torch.uint8, device=device)


def _ptr(t):
    return t.data_ptr() if t.numel() else None


def gs_render(sigmas, coords, colors, rendered_img, s, h, w, c, dmax=float("inf"), *,
              ksigma=None, flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (rendered_img, "rendered_img")):
        _check_input(t, n)
    s, h, w, c = int(s), int(h), int(w), int(c)
    if c != 3:
        raise RuntimeError("libgsraster: c must be 3 (the reference forward hard-codes 3 channels)")
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(coords, (s, 2), "coords")
    _check_shape(colors, (s, 3), "colors")
    if rendered_img.numel() != h * w * c:
        raise RuntimeError("rendered_img must have h*w*c elements")
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace(s, h, w, sigmas.device)
        rc = L.gsr_forward(_ptr(sigmas), _ptr(coords), _ptr(colors), rendered_img.data_ptr(), s, h, w, c,
                           float(dmax), float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags),
                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


def gs_render_backward(sigmas, coords, colors, grads, grads_sigmas, grads_coords, grads_colors,
                       s, h, w, c, dmax=float("inf"), *, ksigma=None, flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (grads, "grads"),
                 (grads_sigmas, "grads_sigmas"), (grads_coords, "grads_coords"),
                 (grads_colors, "grads_colors")):
        _check_input(t, n)
    s, h, w, c = int(s), int(h), int(w), int(c)
    if c != 3:
        raise RuntimeError("libgsraster: c must be 3")
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(coords, (s, 2), "coords")
    _check_shape(colors, (s, 3), "colors")
    _check_shape(grads_sigmas, (s, 3), "grads_sigmas")
    _check_shape(grads_coords, (s, 2), "grads_coords")
    _check_shape(grads_colors, (s, 3), "grads_colors")
    if grads.numel() != h * w * c:
        raise RuntimeError("grads must have h*w*c elements")
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace(s, h, w, sigmas.device)
        rc = L.gsr_backward(_ptr(sigmas), _ptr(coords), _ptr(colors), grads.data_ptr(),
                            _ptr(grads_sigmas), _ptr(grads_coords), _ptr(grads_colors), s, h, w, c,
                            float(dmax), float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags),
                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


# ---- row bands of one image (no counterpart in the reference, whose kernels always walk the whole
# image, gs.cu:38-62): rows [row0, row0 + rows) of the h x w image, same conventions as gs_render.
def gs_render_backward_prepared(sigmas, grads, grads_sigmas, grads_coords, grads_colors, s, h, w, workspace_buf, *,
                                flags=0):
    """gsr_backward_prepared: the backward on a workspace a gs_render (or gsr_prepare) call with the same Gaussians,
    sizes, dmax and ksigma has left behind -- no second set-up.  Gradients are accumulated into grads_*."""
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (grads, "grads"), (grads_sigmas, "grads_sigmas"), (grads_coords, "grads_coords"),
                 (grads_colors, "grads_colors")):
        _check_input(t, n)
    s, h, w = int(s), int(h), int(w)
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(grads_sigmas, (s, 3), "grads_sigmas")
    _check_shape(grads_coords, (s, 2), "grads_coords")
    _check_shape(grads_colors, (s, 3), "grads_colors")
    if grads.numel() != h * w * 3:
        raise RuntimeError("grads must have h*w*3 elements")
    if workspace_buf is None or workspace_buf.device != sigmas.device:
        raise RuntimeError("gs_render_backward_prepared needs the forward call's workspace")
    with torch.cuda.device(sigmas.device):
        rc = L.gsr_backward_prepared(_ptr(sigmas), grads.data_ptr(), _ptr(grads_sigmas), _ptr(grads_coords),
                                     _ptr(grads_colors), s, h, w, _fwd_flags(flags), workspace_buf.data_ptr(),
                                     workspace_buf.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


def gs_render_band(sigmas, coords, colors, band_img, s, h, w, c, row0, rows, dmax=float("inf"), *,
                   ksigma=None, flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors")):
        _check_input(t, n)
    if int(flags) & _lib.GSR_FLAG_U8:  # the fused uint8 post-processing: band_img is the (rows,w,3) uint8 band
        if not (isinstance(band_img, torch.Tensor) and band_img.is_cuda and band_img.is_contiguous()
                and band_img.dtype == torch.uint8):
            raise RuntimeError("band_img must be a contiguous uint8 CUDA tensor with GSR_FLAG_U8")
    else:
        _check_input(band_img, "band_img")
    s, h, w, c, row0, rows = int(s), int(h), int(w), int(c), int(row0), int(rows)
    if c != 3:
        raise RuntimeError("libgsraster: c must be 3 (the reference forward hard-codes 3 channels)")
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(coords, (s, 2), "coords")
    _check_shape(colors, (s, 3), "colors")
    if band_img.numel() != rows * w * c:
        raise RuntimeError("band_img must have rows*w*c elements")
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace(s, rows, w, sigmas.device)
        rc = L.gsr_forward_band(_ptr(sigmas), _ptr(coords), _ptr(colors), band_img.data_ptr(), s, h, w, c,
                                row0, rows, float(dmax), float(_ksigma if ksigma is None else ksigma),
                                _fwd_flags(flags), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


def gs_render_backward_band(sigmas, coords, colors, band_grads, grads_sigmas, grads_coords, grads_colors,
                            s, h, w, c, row0, rows, dmax=float("inf"), *, ksigma=None, flags=0,
                            workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (band_grads, "band_grads"),
                 (grads_sigmas, "grads_sigmas"), (grads_coords, "grads_coords"),
                 (grads_colors, "grads_colors")):
        _check_input(t, n)
    s, h, w, c, row0, rows = int(s), int(h), int(w), int(c), int(row0), int(rows)
    if c != 3:
        raise RuntimeError("libgsraster: c must be 3")
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(coords, (s, 2), "coords")
    _check_shape(colors, (s, 3), "colors")
    _check_shape(grads_sigmas, (s, 3), "grads_sigmas")
    _check_shape(grads_coords, (s, 2), "grads_coords")
    _check_shape(grads_colors, (s, 3), "grads_colors")
    if band_grads.numel() != rows * w * c:
        raise RuntimeError("band_grads must have rows*w*c elements")
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace(s, rows, w, sigmas.device)
        rc = L.gsr_backward_band(_ptr(sigmas), _ptr(coords), _ptr(colors), band_grads.data_ptr(),
                                 _ptr(grads_sigmas), _ptr(grads_coords), _ptr(grads_colors), s, h, w, c,
                                 row0, rows, float(dmax), float(_ksigma if ksigma is None else ksigma),
                                 _fwd_flags(flags), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


# ---- uniform batches (no counterpart in the reference, which renders sample by sample,
# gsasr_model.py:191-233): `batch` samples of the same shape in one set-up and one raster launch.
# sigmas (B,N,3), coords (B,N,2), colors (B,N,3), imgs / grads (B,h,w,3), all contiguous.
def workspace_batch(batch: int, s_per: int, h: int, w: int, device) -> torch.Tensor:
    n = _lib.load().gsr_workspace_bytes_batch_uniform(int(batch), int(s_per), int(h), int(w))
    if n == 0:
        raise RuntimeError(f"libgsraster: bad sizes batch={batch}, s={s_per}, h={h}, w={w}")
    return torch.empty(n, dtype=torch.uint8, device=device)


def gs_render_batch(sigmas, coords, colors, rendered_imgs, dmax=float("inf"), *, ksigma=None, flags=0,
                    workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (rendered_imgs, "rendered_imgs")):
        _check_input(t, n)
    if sigmas.dim() != 3 or rendered_imgs.dim() != 4:
        raise RuntimeError("gs_render_batch: sigmas must be (B,N,3) and rendered_imgs (B,h,w,3)")
    b, s = int(sigmas.shape[0]), int(sigmas.shape[1])
    h, w = int(rendered_imgs.shape[1]), int(rendered_imgs.shape[2])
    _check_shape(sigmas, (b, s, 3), "sigmas")
    _check_shape(coords, (b, s, 2), "coords")
    _check_shape(colors, (b, s, 3), "colors")
    _check_shape(rendered_imgs, (b, h, w, 3), "rendered_imgs")
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace_batch(b, s, h, w, sigmas.device)
        rc = L.gsr_forward_batch_uniform(_ptr(sigmas), _ptr(coords), _ptr(colors), rendered_imgs.data_ptr(), b, s,
                                         h, w, 3, float(dmax), float(_ksigma if ksigma is None else ksigma),
                                         _fwd_flags(flags), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


def gs_render_backward_batch(sigmas, coords, colors, grads, grads_sigmas, grads_coords, grads_colors,
                             dmax=float("inf"), *, ksigma=None, flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (grads, "grads"),
                 (grads_sigmas, "grads_sigmas"), (grads_coords, "grads_coords"),
                 (grads_colors, "grads_colors")):
        _check_input(t, n)
    if sigmas.dim() != 3 or grads.dim() != 4:
        raise RuntimeError("gs_render_backward_batch: sigmas must be (B,N,3) and grads (B,h,w,3)")
    b, s = int(sigmas.shape[0]), int(sigmas.shape[1])
    h, w = int(grads.shape[1]), int(grads.shape[2])
    _check_shape(coords, (b, s, 2), "coords")
    _check_shape(colors, (b, s, 3), "colors")
    _check_shape(grads, (b, h, w, 3), "grads")
    _check_shape(grads_sigmas, (b, s, 3), "grads_sigmas")
    _check_shape(grads_coords, (b, s, 2), "grads_coords")
    _check_shape(grads_colors, (b, s, 3), "grads_colors")
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace_batch(b, s, h, w, sigmas.device)
        rc = L.gsr_backward_batch_uniform(_ptr(sigmas), _ptr(coords), _ptr(colors), grads.data_ptr(),
                                          _ptr(grads_sigmas), _ptr(grads_coords), _ptr(grads_colors), b, s, h, w,
                                          3, float(dmax), float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags),
                                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


# ---- render into a window of a larger destination (gsr_forward_window): the raster kernel writes the
# tile's pixels straight into `dst` -- any float32 CUDA tensor, possibly the memory of ANOTHER GPU
# (peer stores) -- at element offset `origin` for pixel (0,0) with the given strides (in elements);
# only pixels inside one of `clips` = [(x0, y0, x1, y1), ...] (inclusive, render coordinates) are written.
def _window(dst, origin, row_stride, pix_stride, chan_stride, clips, h, w):
    if not (isinstance(dst, torch.Tensor) and dst.is_cuda and dst.dtype == torch.float32):
        raise RuntimeError("dst must be a float32 CUDA tensor")
    if not dst.is_contiguous():
        raise RuntimeError("dst must be contiguous (the strides address its flat storage)")
    clips = [tuple(int(v) for v in c) for c in (clips or [])]
    if len(clips) > _lib.GSR_MAX_CLIP:
        raise RuntimeError(f"at most {_lib.GSR_MAX_CLIP} clip rectangles")
    win = _lib.GsrWindow()
    win.row_stride, win.pix_stride, win.chan_stride, win.nclip = int(row_stride), int(pix_stride), int(chan_stride), len(clips)
    lo = hi = None
    for k, (x0, y0, x1, y1) in enumerate(clips if clips else [(0, 0, w - 1, h - 1)]):
        if not (0 <= x0 <= x1 < w and 0 <= y0 <= y1 < h):
            raise RuntimeError(f"clip rectangle {(x0, y0, x1, y1)} is not inside the {h}x{w} render")
        if clips:
            for j, v in enumerate((x0, y0, x1, y1)):
                win.clip[k][j] = v
        for yy, xx in ((y0, x0), (y1, x1)):  # bounds of the touched elements (strides are non-negative)
            a = int(origin) + yy * win.row_stride + xx * win.pix_stride
            lo = a if lo is None else min(lo, a)
            hi = a + 2 * win.chan_stride if hi is None else max(hi, a + 2 * win.chan_stride)
    if min(win.row_stride, win.pix_stride, win.chan_stride) < 0 or lo < 0 or hi >= dst.numel():
        raise RuntimeError("window leaves the destination tensor")
    return win


def gs_render_window(sigmas, coords, colors, dst, origin, row_stride, pix_stride, chan_stride, clips,
                     s, h, w, dmax=float("inf"), *, ksigma=None, flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors")):
        _check_input(t, n)
    s, h, w = int(s), int(h), int(w)
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(coords, (s, 2), "coords")
    _check_shape(colors, (s, 3), "colors")
    win = _window(dst, origin, row_stride, pix_stride, chan_stride, clips, h, w)
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace(s, h, w, sigmas.device)
        rc = L.gsr_forward_window(_ptr(sigmas), _ptr(coords), _ptr(colors), dst.data_ptr() + 4 * int(origin),
                                  win, s, h, w, 3, float(dmax), float(_ksigma if ksigma is None else ksigma),
                                  _fwd_flags(flags), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


def frontend_render_window(raw, dst, origin, row_stride, pix_stride, chan_stride, clips, h, w, step_size,
                           dmax=float("inf"), *, ksigma=None, flags=0):
    """raw head output (s,9) -> fused activations / mapping -> render into the window (inference)."""
    L = _lib.load()
    _check_input(raw, "raw")
    s, h, w = int(raw.shape[0]), int(h), int(w)
    _check_shape(raw, (s, 9), "raw")
    win = _window(dst, origin, row_stride, pix_stride, chan_stride, clips, h, w)
    with torch.cuda.device(raw.device):
        mapped = torch.empty(max(s, 1) * 8, dtype=torch.float32, device=raw.device)
        ws = workspace(s, h, w, raw.device)
        rc = L.gsr_frontend_forward_window(_ptr(raw), mapped.data_ptr(), dst.data_ptr() + 4 * int(origin), win, s, h,
                                           w, float(step_size), float(dmax),
                                           float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags),
                                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


# ---- fused post-processing: the (h,w,3) UINT8 image of inference_paper.py:136-138 straight from the
# raster kernel -- round(clamp(value, 0, 1) * 255), optionally in cv2's b, g, r channel order -- without
# an fp32 image in between (a quarter of the bytes to store and to copy to the host).
def gs_render_u8(sigmas, coords, colors, out_u8, s, h, w, dmax=float("inf"), *, bgr=False, ksigma=None,
                 workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors")):
        _check_input(t, n)
    if not (isinstance(out_u8, torch.Tensor) and out_u8.is_cuda and out_u8.dtype == torch.uint8 and out_u8.is_contiguous()):
        raise RuntimeError("out_u8 must be a contiguous uint8 CUDA tensor")
    s, h, w = int(s), int(h), int(w)
    _check_shape(sigmas, (s, 3), "sigmas")
    _check_shape(coords, (s, 2), "coords")
    _check_shape(colors, (s, 3), "colors")
    _check_shape(out_u8, (h, w, 3), "out_u8")
    flags = _lib.GSR_FLAG_OVERWRITE | _lib.GSR_FLAG_U8 | (_lib.GSR_FLAG_BGR if bgr else 0)
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace(s, h, w, sigmas.device)
        rc = L.gsr_forward(_ptr(sigmas), _ptr(coords), _ptr(colors), out_u8.data_ptr(), s, h, w, 3, float(dmax),
                           float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags), ws.data_ptr(), ws.numel(),
                           torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


# ---- padded (ragged) batches: B samples of N Gaussians each, rendered at their OWN sizes into the
# top-left corner of their slot of a (B, hmax, wmax, 3) buffer (the training loop's per-sample render +
# F.pad, gsasr_model.py:191-233) in one set-up and one raster launch each way.  hmax % 8 == 0.
def _sizes(sizes, b, hmax, wmax):
    import ctypes

    hw = (ctypes.c_int * (2 * b))()
    if len(sizes) != b:
        raise RuntimeError(f"sizes must list (h, w) for each of the {b} samples")
    for i, (h, w) in enumerate(sizes):
        hw[2 * i], hw[2 * i + 1] = int(h), int(w)
    return hw


def _dmaxes(dmax, b):
    import ctypes

    if isinstance(dmax, (int, float)):
        return None, float(dmax)
    arr = (ctypes.c_float * b)(*[float(v) for v in dmax])
    return arr, 0.0


def workspace_batch_padded(batch: int, s_per: int, hmax: int, wmax: int, device) -> torch.Tensor:
    n = _lib.load().gsr_workspace_bytes_batch_padded(int(batch), int(s_per), int(hmax), int(wmax))
    if n == 0:
        raise RuntimeError(f"libgsraster: bad sizes batch={batch}, s={s_per}, hmax={hmax} (multiple of 8), wmax={wmax}")
    return torch.empty(n, dtype=torch.uint8, device=device)


def gs_render_batch_padded(sigmas, coords, colors, rendered_imgs, sizes, dmax=float("inf"), *, ksigma=None,
                           flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (rendered_imgs, "rendered_imgs")):
        _check_input(t, n)
    b, s = int(sigmas.shape[0]), int(sigmas.shape[1])
    hmax, wmax = int(rendered_imgs.shape[1]), int(rendered_imgs.shape[2])
    _check_shape(sigmas, (b, s, 3), "sigmas")
    _check_shape(coords, (b, s, 2), "coords")
    _check_shape(colors, (b, s, 3), "colors")
    _check_shape(rendered_imgs, (b, hmax, wmax, 3), "rendered_imgs")
    hw = _sizes(sizes, b, hmax, wmax)
    dm_arr, dm = _dmaxes(dmax, b)
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace_batch_padded(b, s, hmax, wmax, sigmas.device)
        rc = L.gsr_forward_batch_padded(_ptr(sigmas), _ptr(coords), _ptr(colors), rendered_imgs.data_ptr(), b, s, hmax,
                                        wmax, hw, dm_arr, dm, float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags),
                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)


def gs_render_backward_batch_padded(sigmas, coords, colors, grads, grads_sigmas, grads_coords, grads_colors, sizes,
                                    dmax=float("inf"), *, ksigma=None, flags=0, workspace_buf=None):
    L = _lib.load()
    for t, n in ((sigmas, "sigmas"), (coords, "coords"), (colors, "colors"), (grads, "grads"),
                 (grads_sigmas, "grads_sigmas"), (grads_coords, "grads_coords"), (grads_colors, "grads_colors")):
        _check_input(t, n)
    b, s = int(sigmas.shape[0]), int(sigmas.shape[1])
    hmax, wmax = int(grads.shape[1]), int(grads.shape[2])
    _check_shape(coords, (b, s, 2), "coords")
    _check_shape(colors, (b, s, 3), "colors")
    _check_shape(grads, (b, hmax, wmax, 3), "grads")
    _check_shape(grads_sigmas, (b, s, 3), "grads_sigmas")
    _check_shape(grads_coords, (b, s, 2), "grads_coords")
    _check_shape(grads_colors, (b, s, 3), "grads_colors")
    hw = _sizes(sizes, b, hmax, wmax)
    dm_arr, dm = _dmaxes(dmax, b)
    with torch.cuda.device(sigmas.device):
        ws = workspace_buf if workspace_buf is not None else workspace_batch_padded(b, s, hmax, wmax, sigmas.device)
        rc = L.gsr_backward_batch_padded(_ptr(sigmas), _ptr(coords), _ptr(colors), grads.data_ptr(), _ptr(grads_sigmas),
                                         _ptr(grads_coords), _ptr(grads_colors), b, s, hmax, wmax, hw, dm_arr, dm,
                                         float(_ksigma if ksigma is None else ksigma), _fwd_flags(flags), ws.data_ptr(),
                                         ws.numel(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
