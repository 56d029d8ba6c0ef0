"""The decoder half of inference_paper.py:117-138 with every fused piece of this library in line:

    encoder_output -> [the reference head's own body] -> fused tail (tcgen05: five MLPs, reference points)
                   -> front end folded into the set-up kernel -> raster -> fused clamp / x255 / round / uint8

i.e. ``decoder(encoder_output, scale)`` + ``generate_2D_gaussian_splatting_step`` + the numpy post-processing as three
kernels' worth of launches after the head's attention blocks.  The encoder and the head's body (embeddings, window
cross-attention, Gaussian self-attention, UPNet) stay the reference's modules, unmodified.
"""
from __future__ import annotations

import torch

from . import gaussian_splatting as _gsp
from . import head_tail as _ht


@torch.no_grad()
def render_from_features(head, encoder_output, scale: float, sr_size, *, dmax: float = 0.1, default_step_size: float = 1.2,
                         uint8: bool = True, bgr: bool = True, packed=None):
    """One image per sample of ``encoder_output`` (b, c, h, w): the reference's

        batch_gs_parameters = decoder(encoder_output, scale_vector)          # inference_paper.py:120
        generate_2D_gaussian_splatting_step(gs_parameters, sr_size, scale, ..., dmax_mode='fix', dmax=dmax)   # :122-132
        clamp_(0, 1) -> [2, 1, 0] -> HWC -> * 255 -> round -> uint8          # :136-138

    ``uint8=True`` returns a list of (H, W, 3) uint8 images (b, g, r order when ``bgr``: what cv2.imwrite takes);
    ``uint8=False`` the (3, H, W) float images of generate_2D_gaussian_splatting_step.  ``head`` is a reference Fea2GS /
    Fea2GS_ROPE_AMP instance."""
    b = encoder_output.shape[0]
    scale_vec = torch.full((b,), float(scale), dtype=torch.float32, device=encoder_output.device)
    raw = _ht.forward_fused_tail(head, encoder_output, scale_vec, packed)      # (b, N, 9)
    size = torch.as_tensor(sr_size)
    sm = torch.tensor([float(scale), float(scale)])
    outs = []
    for i in range(b):
        if uint8:
            outs.append(_gsp.generate_2D_gaussian_splatting_step_u8(size, raw[i], scale, sm, default_step_size=default_step_size,
                                                                    dmax=dmax, bgr=bgr))
        else:
            outs.append(_gsp.generate_2D_gaussian_splatting_step(size, raw[i], scale, sm, default_step_size=default_step_size,
                                                                 dmax=dmax, fused=True))
    return outs
