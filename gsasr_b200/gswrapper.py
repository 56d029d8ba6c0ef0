"""Mirror of utils/gs_cuda_dmax/gswrapper.py:22-53 (and of utils/gs_cuda/gswrapper.py:19-48):
the autograd boundary of the render path, same names, same argument meaning.

    GSCUDA.apply(sigmas, coords, colors, rendered_img, dmax) -> rendered_img (same tensor)
    gaussiansplatting_render(sigmas, coords, colors, image_size, dmax=100) -> (h, w, c)

As in the reference no gradient flows to ``rendered_img`` and backward is once-differentiable.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import gscuda as GSWrapper


class GSCUDA(Function):
    @staticmethod
    def forward(ctx, sigmas, coords, colors, rendered_img, dmax=float("inf")):
        ctx.save_for_backward(sigmas, coords, colors)
        ctx.dmax = dmax
        h, w, c = rendered_img.shape
        s = sigmas.shape[0]
        # the set-up of this call (region buckets, records) serves the backward too: see gscuda.set_reuse_setup
        keep = GSWrapper.get_reuse_setup() and c == 3 and s > 0 and any(ctx.needs_input_grad[:3])
        ws = GSWrapper.workspace(s, h, w, sigmas.device) if keep else None
        GSWrapper.gs_render(sigmas, coords, colors, rendered_img, s, h, w, c, dmax, workspace_buf=ws)
        ctx.ws = ws
        return rendered_img

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        sigmas, coords, colors = ctx.saved_tensors
        dmax = ctx.dmax
        h, w, c = grad_output.shape
        s = sigmas.shape[0]
        grads_sigmas = torch.zeros_like(sigmas)
        grads_coords = torch.zeros_like(coords)
        grads_colors = torch.zeros_like(colors)
        ws, ctx.ws = getattr(ctx, "ws", None), None
        if ws is not None and GSWrapper.get_reuse_setup():
            GSWrapper.gs_render_backward_prepared(sigmas, grad_output.contiguous(), grads_sigmas, grads_coords,
                                                  grads_colors, s, h, w, ws)
        else:
            GSWrapper.gs_render_backward(sigmas, coords, colors, grad_output.contiguous(), grads_sigmas,
                                         grads_coords, grads_colors, s, h, w, c, dmax)
        return (grads_sigmas, grads_coords, grads_colors, None, None)


def gaussiansplatting_render(sigmas, coords, colors, image_size, dmax=100):
    sigmas = sigmas.contiguous()  # (gs num, 3)
    coords = coords.contiguous()  # (gs num, 2)
    colors = colors.contiguous()  # (gs num, c)
    h, w = image_size[:2]
    c = colors.shape[-1]
    rendered_img = torch.zeros(h, w, c, device=colors.device, dtype=torch.float32)
    return GSCUDA.apply(sigmas, coords, colors, rendered_img, dmax)


# ---- uniform batches (no counterpart in the reference, whose training loop renders sample by sample,
# gsasr_model.py:191-233): the same autograd boundary for B samples of one shape, one launch each way.
class GSCUDABatch(Function):
    @staticmethod
    def forward(ctx, sigmas, coords, colors, rendered_imgs, dmax=float("inf")):
        ctx.save_for_backward(sigmas, coords, colors)
        ctx.dmax = dmax
        GSWrapper.gs_render_batch(sigmas, coords, colors, rendered_imgs, dmax)
        return rendered_imgs

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        sigmas, coords, colors = ctx.saved_tensors
        grads_sigmas = torch.zeros_like(sigmas)
        grads_coords = torch.zeros_like(coords)
        grads_colors = torch.zeros_like(colors)
        GSWrapper.gs_render_backward_batch(sigmas, coords, colors, grad_output.contiguous(), grads_sigmas,
                                           grads_coords, grads_colors, ctx.dmax)
        return (grads_sigmas, grads_coords, grads_colors, None, None)


def gaussiansplatting_render_batch(sigmas, coords, colors, image_size, dmax=100):
    """(B,N,3), (B,N,2), (B,N,3) -> (B,h,w,3): gaussiansplatting_render for every sample of a batch."""
    sigmas, coords, colors = sigmas.contiguous(), coords.contiguous(), colors.contiguous()
    h, w = image_size[:2]
    rendered = torch.zeros(sigmas.shape[0], h, w, 3, device=colors.device, dtype=torch.float32)
    return GSCUDABatch.apply(sigmas, coords, colors, rendered, dmax)


class GSCUDABatchPadded(Function):
    """Samples of different sizes in one launch: (B,N,·) parameters -> (B,hmax,wmax,3), sample b rendered at
    sizes[b] = (h_b, w_b) into the top-left corner of its slot, zeros elsewhere (render + F.pad)."""

    @staticmethod
    def forward(ctx, sigmas, coords, colors, rendered_imgs, sizes, dmax=float("inf")):
        ctx.save_for_backward(sigmas, coords, colors)
        ctx.meta = (tuple((int(h), int(w)) for h, w in sizes), dmax)
        GSWrapper.gs_render_batch_padded(sigmas, coords, colors, rendered_imgs, ctx.meta[0], dmax)
        return rendered_imgs

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        sigmas, coords, colors = ctx.saved_tensors
        sizes, dmax = ctx.meta
        grads_sigmas = torch.zeros_like(sigmas)
        grads_coords = torch.zeros_like(coords)
        grads_colors = torch.zeros_like(colors)
        GSWrapper.gs_render_backward_batch_padded(sigmas, coords, colors, grad_output.contiguous(), grads_sigmas,
                                                  grads_coords, grads_colors, sizes, dmax)
        return (grads_sigmas, grads_coords, grads_colors, None, None, None)


def gaussiansplatting_render_batch_padded(sigmas, coords, colors, sizes, dmax=100, hmax=None, wmax=None):
    """(B,N,3), (B,N,2), (B,N,3), sizes [(h_b, w_b)] -> (B,hmax,wmax,3); hmax defaults to the largest h_b
    rounded up to a multiple of 8, wmax to the largest w_b."""
    sigmas, coords, colors = sigmas.contiguous(), coords.contiguous(), colors.contiguous()
    hmax = (max(int(h) for h, _ in sizes) + 7) // 8 * 8 if hmax is None else int(hmax)
    wmax = max(int(w) for _, w in sizes) if wmax is None else int(wmax)
    rendered = torch.zeros(sigmas.shape[0], hmax, wmax, 3, device=colors.device, dtype=torch.float32)
    return GSCUDABatchPadded.apply(sigmas, coords, colors, rendered, sizes, dmax)
