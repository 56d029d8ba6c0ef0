"""Seeded synthetic Gaussian fields with the shapes of BASELINE.json's configs.

A field is the raw head tensor p (N,9) the fea2gs head would emit (utils/fea2gs.py:553-563,
623-633): columns 0-6 ~ N(0,1) ("model-like": sigma = 0.99999*sigmoid(N(0,1))+1e-6, i.e. mean
sigma_px = scale/2.4), columns 7,8 = cell centres of a gh x gw grid in row-major order plus
N(0,(cell/2)^2) jitter.  `compact=True` draws columns 0,1 ~ N(-1.5, 0.5^2) (sigma_px ~ 0.2*scale).
Generated on the CPU with a torch.Generator so that the reference and the new kernels see the
very same tensors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class Config:
    name: str
    lr_h: int
    lr_w: int
    scale: float
    per_lr_side: int  # Gaussians per LR pixel side (2 -> 4 per LR pixel, 4 -> 16 per LR pixel)
    dmax: float

    @property
    def grid(self):
        return self.lr_h * self.per_lr_side, self.lr_w * self.per_lr_side

    @property
    def n(self):
        gh, gw = self.grid
        return gh * gw

    @property
    def hr(self):
        return math.floor(self.lr_h * self.scale), math.floor(self.lr_w * self.scale)


CONFIGS = {
    # BASELINE.json configs[0..2] and the headline shape
    "C1": Config("C1", 64, 64, 2.0, 2, 0.1),       # 16,384 Gaussians -> 128x128
    "C2": Config("C2", 256, 256, 4.0, 2, 0.1),     # 262,144 -> 1024x1024
    "C2d": Config("C2d", 256, 256, 4.0, 4, 0.1),   # 1,048,576 -> 1024x1024 (real 16/LR-px density)
    "C3": Config("C3", 512, 512, 8.0, 2, 0.1),     # 1,048,576 -> 4096x4096
    "HL": Config("HL", 512, 1024, 4.0, 2, 0.1),    # 2,097,152 -> 2048x4096  (headline metric)
    "T480": Config("T480", 480, 480, 4.0, 2, 0.1), # one split_and_joint_image tile of config 4
}


def raw_field(gh: int, gw: int, seed: int = 0, compact: bool = False) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    n = gh * gw
    p = torch.randn(n, 9, generator=g, dtype=torch.float32)
    if compact:
        p[:, 0:2] = -1.5 + 0.5 * p[:, 0:2]
    jj, ii = torch.meshgrid(torch.arange(gw, dtype=torch.float32), torch.arange(gh, dtype=torch.float32),
                            indexing="xy")
    cx = (jj.reshape(-1) + 0.5) / gw
    cy = (ii.reshape(-1) + 0.5) / gh
    p[:, 7] = cx + p[:, 7] * (0.5 / gw)
    p[:, 8] = cy + p[:, 8] * (0.5 / gh)
    return p


def map_field(p: torch.Tensor, h: int, w: int, scale: float, default_step_size: float = 1.2):
    """raw (N,9) -> (sigmas (N,3), coords (N,2), colors (N,3)) with the reference's own expressions
    (utils/gaussian_splatting.py:174-180, 121-123; sr_size / scale_modify as CPU tensors, as
    inference_paper.py:113-131 passes them)."""
    sr_size = torch.tensor([h, w])
    scale_modify = torch.tensor([scale, scale])  # inference_paper.py:125
    step_size = default_step_size / scale_modify[0]
    sigma_x = 0.99999 * torch.sigmoid(p[:, 0:1]) + 1e-6
    sigma_y = 0.99999 * torch.sigmoid(p[:, 1:2]) + 1e-6
    rho = 0.999999 * torch.tanh(p[:, 2:3])
    alpha = torch.sigmoid(p[:, 3:4])
    colours = torch.sigmoid(p[:, 4:7])
    coords = (p[:, 7:9] * 2 - 1)
    colors = (colours * alpha).contiguous()
    sigmas = torch.cat([sigma_y / step_size * 2 / (sr_size[1] - 1),
                        sigma_x / step_size * 2 / (sr_size[0] - 1), rho], dim=-1).contiguous()
    coords[:, 0] = (coords[:, 0] + 1 - 1 / sr_size[1]) * sr_size[1] / (sr_size[1] - 1) - 1.0
    coords[:, 1] = (coords[:, 1] + 1 - 1 / sr_size[0]) * sr_size[0] / (sr_size[0] - 1) - 1.0
    return sigmas, coords.contiguous(), colors


def make(cfg, seed: int = 0, compact: bool = False):
    """Returns (raw, sigmas, coords, colors, h, w) as CPU float32 tensors for a named config."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    gh, gw = cfg.grid
    h, w = cfg.hr
    p = raw_field(gh, gw, seed, compact)
    sigmas, coords, colors = map_field(p, h, w, cfg.scale)
    return p, sigmas, coords, colors, h, w
