"""Mirror of the render front end, utils/gaussian_splatting.py (GSASR): same function names,
same arguments, same (3, H, W) float32 result.

    generate_2D_gaussian_splatting_step(sr_size, gs_parameters, scale, scale_modify, ...)   :158-217
    generate_2D_gaussian_splatting_step_buffer(..., buffer_size=4000000)                    :219-265
    rendering_cuda_dmax / rendering_cuda / *_buffer                                          :86-155

The activations (:174-180) and the unit / coordinate mapping (:121-123) are executed with the
very same torch expressions as the reference, so the tensors handed to the rasteriser are
bit-identical to the reference's.  What changes is below that line: the B200 kernels write the
(3,H,W) image directly (no zero-fill, no permute().contiguous() pass), and the `_buffer`
variants differentiate correctly (the reference drops the gradient of all but the last chunk,
gswrapper.py:44).  ``fused=True`` additionally replaces the ~15 elementwise torch kernels by the
library's fused front end (gsr_frontend_forward / _backward; parity 1e-4, not bitwise).

``cuda_rendering=False`` selects rendering_python (:11-84), the reference's PyTorch resampling renderer: every
Gaussian is tabulated on a num_step x num_step grid, normalised by its largest sample, and resampled bilinearly
onto the image.  It is a DIFFERENT function of the parameters than the CUDA kernels (max-abs 0.65 apart on a
random field) and the path BASELINE config 1 times; mirrored here in plain torch ops (any device)
so that callers that pass cuda_rendering=False keep working.  It is not a fallback of the CUDA path: nothing
routes to it unless the caller asks for it by name.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from . import gscuda as _gs

__all__ = [
    "generate_2D_gaussian_splatting_step", "generate_2D_gaussian_splatting_step_buffer", "rendering_python",
    "generate_2D_gaussian_splatting_step_batch", "generate_2D_gaussian_splatting_step_batch_padded",
    "generate_2D_gaussian_splatting_step_u8", "render_into_canvas",
    "rendering_cuda", "rendering_cuda_buffer", "rendering_cuda_dmax", "rendering_cuda_dmax_buffer",
    "map_gaussians", "render_chw",
]

_CHW = _lib.GSR_FLAG_CHW
_OVER = _lib.GSR_FLAG_OVERWRITE


class _RenderCHW(Function):
    """(sigmas, coords, colors) -> (3,H,W); Gaussians optionally processed in chunks of buffer_size."""

    @staticmethod
    def forward(ctx, sigmas, coords, colors, h, w, dmax, buffer_size):
        sigmas, coords, colors = sigmas.contiguous(), coords.contiguous(), colors.contiguous()
        ctx.save_for_backward(sigmas, coords, colors)
        ctx.meta = (int(h), int(w), float(dmax), buffer_size)
        h, w = int(h), int(w)
        out = torch.empty(3, h, w, device=sigmas.device, dtype=torch.float32)
        n = sigmas.shape[0]
        step = n if not buffer_size else int(buffer_size)
        first = True
        # one chunk: its set-up (region buckets, records) serves the backward too (gscuda.set_reuse_setup)
        keep = _gs.get_reuse_setup() and 0 < n <= step and any(ctx.needs_input_grad[:3])
        ctx.ws = _gs.workspace(n, h, w, sigmas.device) if keep else None
        for a in range(0, max(n, 1), max(step, 1)):
            b = min(a + step, n)
            _gs.gs_render(sigmas[a:b], coords[a:b], colors[a:b], out, b - a, h, w, 3, dmax,
                          flags=_CHW | (_OVER if first else 0), workspace_buf=ctx.ws)
            first = False
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        sigmas, coords, colors = ctx.saved_tensors
        h, w, dmax, buffer_size = ctx.meta
        grad = grad.contiguous()
        gs, gc, gk = torch.zeros_like(sigmas), torch.zeros_like(coords), torch.zeros_like(colors)
        n = sigmas.shape[0]
        step = n if not buffer_size else int(buffer_size)
        ws, ctx.ws = getattr(ctx, "ws", None), None
        if ws is not None and _gs.get_reuse_setup():
            _gs.gs_render_backward_prepared(sigmas, grad, gs, gc, gk, n, h, w, ws, flags=_CHW)
            return gs, gc, gk, None, None, None, None
        for a in range(0, n, max(step, 1)):
            b = min(a + step, n)
            _gs.gs_render_backward(sigmas[a:b], coords[a:b], colors[a:b], grad, gs[a:b], gc[a:b],
                                   gk[a:b], b - a, h, w, 3, dmax, flags=_CHW)
        return gs, gc, gk, None, None, None, None


def render_chw(sigmas, coords, colors, h, w, dmax=float("inf"), buffer_size=None):
    """Differentiable render straight into a (3,h,w) image."""
    return _RenderCHW.apply(sigmas, coords, colors, h, w, dmax, buffer_size)


def map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size):
    """Unit / coordinate mapping of rendering_cuda_dmax (:121-123), same expressions.
    NOTE the x/y swap: the head's sigma_x is the vertical std.  ``coords`` is modified in place,
    as in the reference."""
    sigmas = torch.cat([sigma_y / step_size * 2 / (sr_size[1] - 1),
                        sigma_x / step_size * 2 / (sr_size[0] - 1), rho], dim=-1).contiguous()
    coords[:, 0] = (coords[:, 0] + 1 - 1 / sr_size[1]) * sr_size[1] / (sr_size[1] - 1) - 1.0
    coords[:, 1] = (coords[:, 1] + 1 - 1 / sr_size[0]) * sr_size[0] / (sr_size[0] - 1) - 1.0
    return sigmas, coords


def rendering_cuda_dmax(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size, step_size, device,
                        dmax=1):
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    return render_chw(sigmas, coords, colours_with_alpha, int(sr_size[0]), int(sr_size[1]), dmax)


def rendering_cuda(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size, step_size, device):
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    return render_chw(sigmas, coords, colours_with_alpha, int(sr_size[0]), int(sr_size[1]))


def rendering_cuda_dmax_buffer(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size, step_size,
                               device, dmax=1, buffer_size=1000000):
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    return render_chw(sigmas, coords, colours_with_alpha, int(sr_size[0]), int(sr_size[1]), dmax,
                      buffer_size)


def rendering_cuda_buffer(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size, step_size, device,
                          buffer_size=1000000):
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    return render_chw(sigmas, coords, colours_with_alpha, int(sr_size[0]), int(sr_size[1]),
                      float("inf"), buffer_size)


class _FusedFrontend(Function):
    """raw (N,9) -> (3,H,W) through gsr_frontend_forward / gsr_frontend_backward."""

    @staticmethod
    def forward(ctx, raw, h, w, step_size, dmax):
        L = _lib.load()
        raw = raw.contiguous().float()
        n = raw.shape[0]
        mapped = torch.empty(max(n, 1) * 8, device=raw.device, dtype=torch.float32)
        out = torch.empty(3, h, w, device=raw.device, dtype=torch.float32)
        with torch.cuda.device(raw.device):
            ws = _gs.workspace(n, h, w, raw.device)
            rc = L.gsr_frontend_forward(raw.data_ptr(), mapped.data_ptr(), out.data_ptr(), n, h, w,
                                        float(step_size), float(dmax), float(_gs.get_ksigma()),
                                        ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        ctx.save_for_backward(raw, mapped)
        ctx.meta = (h, w, float(step_size), float(dmax))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        L = _lib.load()
        raw, mapped = ctx.saved_tensors
        h, w, step_size, dmax = ctx.meta
        n = raw.shape[0]
        grad = grad.contiguous()
        g_raw = torch.zeros_like(raw)
        if n:
            with torch.cuda.device(raw.device):
                need = L.gsr_workspace_bytes(n, h, w) + (n * 32 + 255) // 256 * 256
                ws = torch.empty(need, dtype=torch.uint8, device=raw.device)
                rc = L.gsr_frontend_backward(raw.data_ptr(), mapped.data_ptr(), grad.data_ptr(),
                                             g_raw.data_ptr(), n, h, w, step_size, dmax,
                                             float(_gs.get_ksigma()), ws.data_ptr(), ws.numel(),
                                             torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
        return g_raw, None, None, None, None


def _step_size(scale, scale_modify, default_step_size, mode):
    """Step size from the scale factor (:164-171): the same checks for the unfused and the fused paths."""
    if mode == 'scale':
        final_scale = scale
    elif mode == 'scale_modify':
        assert scale_modify[0] == scale_modify[1], f"scale_modify is not the same-{scale_modify}"
        final_scale = scale_modify[0]
    else:
        raise ValueError(f"mode-{mode} must be scale or scale_modify")
    return default_step_size / final_scale


def _prepare(gs_parameters, scale, scale_modify, default_step_size, mode):
    step_size = _step_size(scale, scale_modify, default_step_size, mode)
    # prepare gaussian properties (:174-180)
    sigma_x = 0.99999 * torch.sigmoid(gs_parameters[:, 0:1]) + 1e-6
    sigma_y = 0.99999 * torch.sigmoid(gs_parameters[:, 1:2]) + 1e-6
    rho = 0.999999 * torch.tanh(gs_parameters[:, 2:3])
    alpha = torch.sigmoid(gs_parameters[:, 3:4])
    colours = torch.sigmoid(gs_parameters[:, 4:7])
    coords = (gs_parameters[:, 7:9] * 2 - 1)
    colours_with_alpha = colours * alpha
    return step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha


def _resolve_dmax(if_dmax, dmax_mode, dmax, sr_size):
    if not if_dmax:
        return float("inf")
    if dmax_mode == 'dynamic':  # (:203-204)
        return float((dmax + 2) / min(sr_size[0], sr_size[1]))
    if dmax_mode == 'fix':
        return float(dmax)
    raise ValueError(f"dmax_mode-{dmax_mode} must be fix or dynamic")


def rendering_python(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size, step_size, device,
                     max_buffer=2000):
    """The reference's PyTorch renderer (utils/gaussian_splatting.py:11-84), restated.

    Per Gaussian: K[i,j] = N(d_ij; 0, Sigma) on the num_step x num_step grid d_ij = ((i, j) - (n-1)/2) * step_size
    (num_step = int(20 / step_size); sigma_x pairs with the ROW offset, :37-45), divided by (max_ij K + 1e-4)
    (:61-65); the table is resampled bilinearly with zero padding onto the (H, W) image through an affine grid
    that maps output coordinate o (align_corners=False) to table coordinate (o - centre) * size / num_step
    (:69-80); colour-weighted tables are summed.  Differences to the reference's code: the 2x2 covariance is
    inverted in closed form, and one channel is resampled instead of three identical ones (a third of the
    (2000,3,H,W) temporaries); results agree to fp32 round-off."""
    import torch.nn.functional as F

    sr_h, sr_w = int(sr_size[0]), int(sr_size[1])
    step_size = float(step_size)
    sx, sy, r = sigma_x.reshape(-1).float(), sigma_y.reshape(-1).float(), rho.reshape(-1).float()
    det = sx * sx * sy * sy - (r * sx * sy) ** 2
    if bool((det < 0).any()):
        raise ValueError("Covariance matrix must be positive semi-definite")
    # inverse covariance: [[sy^2, -r sx sy], [-r sx sy, sx^2]] / det
    ixx, ixy, iyy = sy * sy / det, -r * sx * sy / det, sx * sx / det
    num_step = int(10 * 2 / step_size)
    ax = torch.arange(num_step, dtype=torch.float32, device=device) * step_size
    ax = ax - ax.mean()
    u, v = ax[:, None], ax[None, :]  # u: row offset (pairs with sigma_x), v: column offset
    norm = 1.0 / (2.0 * torch.pi * torch.sqrt(det))
    final_image = torch.zeros((3, sr_h, sr_w), device=device)
    n = sx.shape[0]
    for a in range(0, n, max_buffer):
        b = min(a + max_buffer, n)
        z = -0.5 * (ixx[a:b, None, None] * u * u + 2.0 * ixy[a:b, None, None] * u * v + iyy[a:b, None, None] * v * v)
        kernel = torch.exp(z) * norm[a:b, None, None]
        kmax = kernel.amax(dim=(-2, -1), keepdim=True)
        kernel = (kernel / (kmax + 1e-4)).unsqueeze(1)  # (b,1,n,n)
        theta = torch.zeros(b - a, 2, 3, dtype=torch.float32, device=device)
        theta[:, 0, 0] = sr_w / num_step
        theta[:, 1, 1] = sr_h / num_step
        theta[:, 0, 2] = -coords[a:b, 0] * sr_w / num_step
        theta[:, 1, 2] = -coords[a:b, 1] * sr_h / num_step
        grid = F.affine_grid(theta, size=(b - a, 1, sr_h, sr_w), align_corners=False)
        moved = F.grid_sample(kernel, grid, align_corners=False)  # (b,1,H,W)
        final_image += torch.einsum('bhw,bc->chw', moved[:, 0], colours_with_alpha[a:b].float())
    return final_image


def _python_renderer(gs_parameters, sr_size, scale, scale_modify, default_step_size, mode):
    step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha = _prepare(
        gs_parameters, scale, scale_modify, default_step_size, mode)
    return rendering_python(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size, step_size,
                            device=sigma_x.device)


def _sample(final_image, sample_coords):
    if sample_coords is not None:  # (:214-216)
        sample_RGB_values = [final_image[:, coord[0], coord[1]] for coord in sample_coords]
        final_image = torch.stack(sample_RGB_values, dim=1)
    return final_image


def generate_2D_gaussian_splatting_step(sr_size, gs_parameters, scale, scale_modify,
                                        sample_coords=None, default_step_size=1.2,
                                        cuda_rendering=True, mode='scale_modify',
                                        if_dmax=True, dmax_mode='fix', dmax=25, fused=False):
    if not cuda_rendering:  # the reference's PyTorch renderer, on request (:210-212)
        return _sample(_python_renderer(gs_parameters, sr_size, scale, scale_modify, default_step_size, mode),
                       sample_coords)
    if fused and not _gs.get_deterministic():  # (the fused entry points carry no flags: deterministic mode renders unfused)
        step_size = float(_step_size(scale, scale_modify, default_step_size, mode))
        final_image = _FusedFrontend.apply(gs_parameters, int(sr_size[0]), int(sr_size[1]), step_size,
                                           _resolve_dmax(if_dmax, dmax_mode, dmax, sr_size))
        return _sample(final_image, sample_coords)
    step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha = _prepare(
        gs_parameters, scale, scale_modify, default_step_size, mode)
    if if_dmax:
        final_image = rendering_cuda_dmax(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size,
                                          step_size, dmax=_resolve_dmax(True, dmax_mode, dmax, sr_size),
                                          device=sigma_x.device)
    else:
        final_image = rendering_cuda(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size,
                                     step_size, device=sigma_x.device)
    return _sample(final_image, sample_coords)


def generate_2D_gaussian_splatting_step_buffer(sr_size, gs_parameters, scale, scale_modify,
                                               sample_coords=None, default_step_size=1.2,
                                               cuda_rendering=True, mode='scale_modify',
                                               if_dmax=True, dmax_mode='fix', dmax=25,
                                               buffer_size=4000000):
    if not cuda_rendering:  # (:258-260) the reference ignores buffer_size on this path too
        return _sample(_python_renderer(gs_parameters, sr_size, scale, scale_modify, default_step_size, mode),
                       sample_coords)
    step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha = _prepare(
        gs_parameters, scale, scale_modify, default_step_size, mode)
    if if_dmax:
        final_image = rendering_cuda_dmax_buffer(sigma_x, sigma_y, rho, coords, colours_with_alpha,
                                                 sr_size, step_size,
                                                 dmax=_resolve_dmax(True, dmax_mode, dmax, sr_size),
                                                 device=sigma_x.device, buffer_size=buffer_size)
    else:
        final_image = rendering_cuda_buffer(sigma_x, sigma_y, rho, coords, colours_with_alpha, sr_size,
                                            step_size, device=sigma_x.device, buffer_size=buffer_size)
    return _sample(final_image, sample_coords)


class _FusedFrontendBatch(Function):
    """raw (B,N,9) -> (B,H,W,3) through gsr_frontend_forward_batch_uniform / _backward_batch_uniform."""

    @staticmethod
    def forward(ctx, raw, h, w, step_size, dmax):
        L = _lib.load()
        raw = raw.contiguous().float()
        b, n = raw.shape[:2]
        mapped = torch.empty(max(b * n, 1) * 8, device=raw.device, dtype=torch.float32)
        out = torch.empty(b, h, w, 3, device=raw.device, dtype=torch.float32)
        with torch.cuda.device(raw.device):
            ws = _gs.workspace_batch(b, n, h, w, raw.device)
            rc = L.gsr_frontend_forward_batch_uniform(raw.data_ptr(), mapped.data_ptr(), out.data_ptr(), b, n, h, w,
                                                      float(step_size), float(dmax), float(_gs.get_ksigma()),
                                                      ws.data_ptr(), ws.numel(),
                                                      torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        ctx.save_for_backward(raw, mapped)
        ctx.meta = (h, w, float(step_size), float(dmax))
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        L = _lib.load()
        raw, mapped = ctx.saved_tensors
        h, w, step_size, dmax = ctx.meta
        b, n = raw.shape[:2]
        grad = grad.contiguous()
        g_raw = torch.zeros_like(raw)
        if b * n:
            with torch.cuda.device(raw.device):
                need = L.gsr_workspace_bytes_batch_uniform(b, n, h, w) + (b * n * 32 + 255) // 256 * 256
                ws = torch.empty(need, dtype=torch.uint8, device=raw.device)
                rc = L.gsr_frontend_backward_batch_uniform(raw.data_ptr(), mapped.data_ptr(), grad.data_ptr(),
                                                           g_raw.data_ptr(), b, n, h, w, step_size, dmax,
                                                           float(_gs.get_ksigma()), ws.data_ptr(), ws.numel(),
                                                           torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
        return g_raw, None, None, None, None


def generate_2D_gaussian_splatting_step_batch(sr_size, gs_parameters, scale, scale_modify,
                                              default_step_size=1.2, mode='scale_modify', if_dmax=True,
                                              dmax_mode='fix', dmax=25, fused=False):
    """generate_2D_gaussian_splatting_step for a whole batch of one shape: gs_parameters (B,N,9) ->
    (B,3,H,W), the training loop's per-sample calls (gsasr_model.py:191-233) in one set-up and one
    raster launch each way (gsr_forward_batch_uniform).  Same activations and mapping expressions as the
    per-sample function, applied to the flattened batch, so the tensors handed to the rasteriser are
    bit-identical (``fused=True``: the library's fused front end instead, parity 1e-4).  The result is a
    view of the (B,H,W,3) render: a (B,3,H,W) tensor in torch.channels_last memory format (no transpose
    pass)."""
    from .gswrapper import gaussiansplatting_render_batch

    if gs_parameters.dim() != 3 or gs_parameters.shape[-1] != 9:
        raise RuntimeError("gs_parameters must be (B,N,9)")
    b, n = gs_parameters.shape[:2]
    dm = _resolve_dmax(if_dmax, dmax_mode, dmax, sr_size)
    if fused and not _gs.get_deterministic():  # (the fused entry points carry no flags: deterministic mode renders unfused)
        step_size = float(_step_size(scale, scale_modify, default_step_size, mode))
        out = _FusedFrontendBatch.apply(gs_parameters, int(sr_size[0]), int(sr_size[1]), step_size, dm)
        return out.permute(0, 3, 1, 2)
    step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha = _prepare(
        gs_parameters.reshape(b * n, 9), scale, scale_modify, default_step_size, mode)
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    out = gaussiansplatting_render_batch(sigmas.view(b, n, 3), coords.view(b, n, 2),
                                         colours_with_alpha.contiguous().view(b, n, 3),
                                         (int(sr_size[0]), int(sr_size[1])), dm)
    return out.permute(0, 3, 1, 2)


def render_into_canvas(canvas, y0, x0, regions, sr_size, gs_parameters, scale, scale_modify,
                       default_step_size=1.2, mode='scale_modify', if_dmax=True, dmax_mode='fix', dmax=25,
                       fused=False):
    """Inference-only: generate_2D_gaussian_splatting_step whose (3,h,w) result is written straight into
    `canvas` -- a contiguous (..., 3, H, W) float32 CUDA tensor (its last three dimensions are addressed),
    possibly the memory of another GPU -- with its pixel (0,0) at canvas (y0, x0); only the pixels inside
    `regions` = [(ya, yb, xa, xb), ...] (half-open, CANVAS coordinates) are written, the others are left
    as they are.  Same activations / mapping expressions as the per-tile function; no tile buffer, no
    paste pass."""
    h, w = int(sr_size[0]), int(sr_size[1])
    H, W = int(canvas.shape[-2]), int(canvas.shape[-1])
    clips = []
    for ya, yb, xa, xb in regions:
        ya, yb, xa, xb = max(ya - y0, 0), min(yb - y0, h), max(xa - x0, 0), min(xb - x0, w)
        if ya < yb and xa < xb:
            clips.append((xa, ya, xb - 1, yb - 1))
    if not clips:
        return
    dm = _resolve_dmax(if_dmax, dmax_mode, dmax, sr_size)
    if fused and not _gs.get_deterministic():  # (the fused entry points carry no flags: deterministic mode renders unfused)  # the library's fused front end (parity 1e-4) instead of ~15 elementwise torch kernels
        step_size = float(_step_size(scale, scale_modify, default_step_size, mode))
        _gs.frontend_render_window(gs_parameters.contiguous().float(), canvas, y0 * W + x0, W, 1, H * W, clips,
                                   h, w, step_size, dm, flags=_OVER)
        return
    step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha = _prepare(
        gs_parameters, scale, scale_modify, default_step_size, mode)
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    _gs.gs_render_window(sigmas, coords.contiguous(), colours_with_alpha.contiguous(), canvas,
                         y0 * W + x0, W, 1, H * W, clips, sigmas.shape[0], h, w, dm, flags=_OVER)


def generate_2D_gaussian_splatting_step_u8(sr_size, gs_parameters, scale, scale_modify, default_step_size=1.2,
                                           mode='scale_modify', if_dmax=True, dmax_mode='fix', dmax=25, bgr=True):
    """Inference-only: generate_2D_gaussian_splatting_step followed by the post-processing of
    inference_paper.py:136-138 (clamp_(0,1), [2,1,0] channel swap, HWC, *255, round, uint8) in one
    pass: the raster kernel writes the (H,W,3) uint8 image (b,g,r order by default, as cv2.imwrite
    wants it) directly.  Equals the reference's chain except where a value falls within fp32 round-off
    of a rounding boundary (differences of one level on isolated pixels)."""
    step_size, sigma_x, sigma_y, rho, coords, colours_with_alpha = _prepare(
        gs_parameters, scale, scale_modify, default_step_size, mode)
    sigmas, coords = map_gaussians(sigma_x, sigma_y, rho, coords, sr_size, step_size)
    h, w = int(sr_size[0]), int(sr_size[1])
    out = torch.empty(h, w, 3, dtype=torch.uint8, device=sigmas.device)
    _gs.gs_render_u8(sigmas, coords.contiguous(), colours_with_alpha.contiguous(), out, sigmas.shape[0], h, w,
                     _resolve_dmax(if_dmax, dmax_mode, dmax, sr_size), bgr=bgr)
    return out


class _FusedFrontendBatchPadded(Function):
    """raw (B,N,9) -> (B,hmax,wmax,3) through gsr_frontend_forward_batch_padded / _backward_batch_padded."""

    @staticmethod
    def forward(ctx, raw, sizes, steps, dmaxes, hmax, wmax):
        import ctypes

        L = _lib.load()
        raw = raw.contiguous().float()
        b, n = raw.shape[:2]
        hw = (ctypes.c_int * (2 * b))(*[int(v) for s_ in sizes for v in s_])
        st = (ctypes.c_float * b)(*[float(v) for v in steps])
        dm = (ctypes.c_float * b)(*[float(v) for v in dmaxes])
        mapped = torch.empty(max(b * n, 1) * 8, device=raw.device, dtype=torch.float32)
        out = torch.empty(b, hmax, wmax, 3, device=raw.device, dtype=torch.float32)
        with torch.cuda.device(raw.device):
            ws = _gs.workspace_batch_padded(b, n, hmax, wmax, raw.device)
            rc = L.gsr_frontend_forward_batch_padded(raw.data_ptr(), mapped.data_ptr(), out.data_ptr(), b, n, hmax, wmax,
                                                     hw, st, dm, 0.0, float(_gs.get_ksigma()), ws.data_ptr(), ws.numel(),
                                                     torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        ctx.save_for_backward(raw, mapped)
        ctx.meta = (hw, st, dm, hmax, wmax)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        L = _lib.load()
        raw, mapped = ctx.saved_tensors
        hw, st, dm, hmax, wmax = ctx.meta
        b, n = raw.shape[:2]
        grad = grad.contiguous()
        g_raw = torch.zeros_like(raw)
        if b * n:
            with torch.cuda.device(raw.device):
                need = L.gsr_workspace_bytes_batch_padded(b, n, hmax, wmax) + (b * n * 32 + 255) // 256 * 256
                ws = torch.empty(need, dtype=torch.uint8, device=raw.device)
                rc = L.gsr_frontend_backward_batch_padded(raw.data_ptr(), mapped.data_ptr(), grad.data_ptr(),
                                                          g_raw.data_ptr(), b, n, hmax, wmax, hw, st, dm, 0.0,
                                                          float(_gs.get_ksigma()), ws.data_ptr(), ws.numel(),
                                                          torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
        return g_raw, None, None, None, None, None


def generate_2D_gaussian_splatting_step_batch_padded(sr_sizes, gs_parameters, scales, default_step_size=1.2,
                                                     if_dmax=True, dmax_mode='fix', dmax=25, hmax=None, wmax=None,
                                                     fused=False):
    """The training loop of gsasr_model.py:191-233 in one call: gs_parameters (B,N,9), sample b rendered at
    sr_sizes[b] = (H_b, W_b) with step size default_step_size / scales[b] and padded with zeros to the
    largest size (the loop's F.pad) -> (B,3,hmax,wmax) in channels-last layout.  Same activations and mapping
    expressions as the per-sample function with the per-sample constants broadcast over the batch
    (``fused=True``: the library's fused front end instead); one set-up and one raster launch each way
    (gsr_forward_batch_padded)."""
    from .gswrapper import gaussiansplatting_render_batch_padded

    if gs_parameters.dim() != 3 or gs_parameters.shape[-1] != 9:
        raise RuntimeError("gs_parameters must be (B,N,9)")
    b = gs_parameters.shape[0]
    dev = gs_parameters.device
    sizes = [(int(s[0]), int(s[1])) for s in sr_sizes]
    if not if_dmax:
        dm = float("inf")
    elif dmax_mode == 'dynamic':
        dm = [float((dmax + 2) / min(h, w)) for h, w in sizes]
    elif dmax_mode == 'fix':
        dm = float(dmax)
    else:
        raise ValueError(f"dmax_mode-{dmax_mode} must be fix or dynamic")
    if fused and not _gs.get_deterministic():  # (the fused entry points carry no flags: deterministic mode renders unfused)
        hm = (max(h for h, _ in sizes) + 7) // 8 * 8 if hmax is None else int(hmax)
        wm = max(w for _, w in sizes) if wmax is None else int(wmax)
        out = _FusedFrontendBatchPadded.apply(gs_parameters, sizes, [default_step_size / float(sc) for sc in scales],
                                              dm if isinstance(dm, list) else [dm] * b, hm, wm)
        return out.permute(0, 3, 1, 2)
    hs = torch.tensor([h for h, _ in sizes], device=dev).view(b, 1, 1)
    ws_ = torch.tensor([w for _, w in sizes], device=dev).view(b, 1, 1)
    step = torch.tensor([default_step_size / float(sc) for sc in scales], dtype=torch.float32, device=dev).view(b, 1, 1)
    # prepare gaussian properties (:174-180)
    sigma_x = 0.99999 * torch.sigmoid(gs_parameters[..., 0:1]) + 1e-6
    sigma_y = 0.99999 * torch.sigmoid(gs_parameters[..., 1:2]) + 1e-6
    rho = 0.999999 * torch.tanh(gs_parameters[..., 2:3])
    alpha = torch.sigmoid(gs_parameters[..., 3:4])
    colours = torch.sigmoid(gs_parameters[..., 4:7])
    coords = (gs_parameters[..., 7:9] * 2 - 1)
    colours_with_alpha = colours * alpha
    # unit / coordinate mapping (:121-123) with every sample's own size and step
    sigmas = torch.cat([sigma_y / step * 2 / (ws_ - 1), sigma_x / step * 2 / (hs - 1), rho], dim=-1).contiguous()
    cx = (coords[..., 0:1] + 1 - 1 / ws_) * ws_ / (ws_ - 1) - 1.0
    cy = (coords[..., 1:2] + 1 - 1 / hs) * hs / (hs - 1) - 1.0
    coords = torch.cat([cx, cy], dim=-1).contiguous()
    out = gaussiansplatting_render_batch_padded(sigmas, coords, colours_with_alpha.contiguous(), sizes, dm, hmax, wmax)
    return out.permute(0, 3, 1, 2)
