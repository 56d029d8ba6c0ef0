"""The training loop's pixel loss as one kernel: crop + L1 + gradient (gsasr_model.py:212-234).

The reference renders every sample at its own size, pads it to the batch's largest size, crops output and
ground truth back to the sample's size and adds ``L1Loss(reduction='mean')`` per sample, divided by the batch
size.  ``l1_crop_loss_padded`` computes exactly that sum from the padded batch render
(``generate_2D_gaussian_splatting_step_batch_padded``) and the padded ground-truth batch in ONE pass that also
writes dloss/dsr -- zero in the padding -- so autograd's backward is a multiplication by the incoming scalar.
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib


def _strides4(t):
    n, c, h, w = t.stride()
    return (ctypes.c_longlong * 4)(n, c, h, w)


class _L1CropPadded(Function):
    @staticmethod
    def forward(ctx, sr, gt, sizes, weight):
        L = _lib.load()
        b, c, hmax, wmax = sr.shape
        if c != 3 or gt.dim() != 4 or gt.shape[0] != b or gt.shape[1] != 3:
            raise RuntimeError("sr must be (B,3,hmax,wmax) and gt (B,3,H,W)")
        if sr.dtype != torch.float32 or gt.dtype != torch.float32 or not sr.is_cuda or gt.device != sr.device:
            raise RuntimeError("sr and gt must be float32 CUDA tensors on one device")
        for hb, wb in sizes:
            if not (1 <= hb <= min(hmax, gt.shape[2]) and 1 <= wb <= min(wmax, gt.shape[3])):
                raise RuntimeError(f"sample size {(hb, wb)} outside sr {(hmax, wmax)} / gt {tuple(gt.shape[2:])}")
        grad = torch.empty_strided(sr.shape, sr.stride(), dtype=torch.float32, device=sr.device)
        loss = torch.empty((), dtype=torch.float32, device=sr.device)
        hw = (ctypes.c_int * (2 * b))(*[int(v) for s in sizes for v in s])
        with torch.cuda.device(sr.device):
            ws = torch.empty(L.gsr_l1_crop_workspace_bytes(), dtype=torch.uint8, device=sr.device)
            rc = L.gsr_l1_crop_loss(sr.data_ptr(), _strides4(sr), gt.data_ptr(), _strides4(gt), grad.data_ptr(),
                                    loss.data_ptr(), b, hmax, wmax, hw, float(weight) / max(b, 1), 0, ws.data_ptr(),
                                    ws.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def l1_crop_loss_padded(sr, gt, sizes, loss_weight: float = 1.0):
    """loss_weight / B * sum_b L1Loss(mean)(sr[b, :, :h_b, :w_b], gt[b, :, :h_b, :w_b]) -- the `l_pix` of
    gsasr_model.py:226-234 for a padded batch.  sr: (B,3,hmax,wmax) float32 CUDA, any strides without overlap
    (e.g. the channels-last result of the padded batch render); gt: (B,3,H,W) with H >= h_b, W >= w_b; sizes:
    B pairs (h_b, w_b).  Differentiable w.r.t. sr; the gradient is zero in the padding."""
    sizes = [(int(s[0]), int(s[1])) for s in sizes]
    if len(sizes) != sr.shape[0]:
        raise RuntimeError("one (h, w) per sample")
    return _L1CropPadded.apply(sr, gt, sizes, float(loss_weight))
