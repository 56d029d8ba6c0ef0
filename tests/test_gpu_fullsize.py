"""Parity at BASELINE.json's FULL shapes -- the sizes at which the kernels take code paths the small tests
never reach (warp-ballot region build and gmem backward sweeps at x8, fp32 coordinate rounding at w = 4096,
the truncation margin of the real 16-Gaussians-per-LR-pixel density).  `-m gpu`.

Checkers: the CPU oracle on crops / Gaussian subsets (exact: every Gaussian whose dmax window reaches the crop
is summed), and the reference's own CUDA kernels (oracle/_ref) on whole images where they finish in seconds.
Tolerances (BASELINE.json north_star): forward <= 1e-4 abs; inclusion counts exact; backward <= 1e-3 of the
largest gradient of each tensor.  All renders use the library-default k-sigma truncation unless stated.
"""
import numpy as np
import pytest
import torch

from gsasr_b200 import fields, gscuda
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FWD_TOL = 1e-4
BWD_RTOL = 1e-3
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")


def _fwd(s, c, k, h, w, dmax, ksigma=None):
    img = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(s, c, k, img, s.shape[0], h, w, 3, dmax, ksigma=ksigma)
    torch.cuda.synchronize()
    return img


def _bwd(s, c, k, g, dmax, ksigma=None):
    gs, gc, gk = torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)
    gscuda.gs_render_backward(s, c, k, g, gs, gc, gk, s.shape[0], g.shape[0], g.shape[1], 3, dmax, ksigma=ksigma)
    torch.cuda.synchronize()
    return gs, gc, gk


def _corner_crops(h, w, ch, cw):
    """The four image corners plus one interior window."""
    return [(0, 0), (0, w - cw), (h - ch, 0), (h - ch, w - cw), ((h // 2) // 8 * 8 + 3, (w // 3) // 8 * 8 + 5)]


def _check_crops(img, s, c, k, h, w, dmax, ch, cw):
    sn, cn, kn = s.cpu().numpy(), c.cpu().numpy(), k.cpu().numpy()
    worst = 0.0
    rr = oracle.ranges(cn, h, w, dmax)
    for y0, x0 in _corner_crops(h, w, ch, cw):
        ref = oracle.forward_crop(sn, cn, kn, h, w, dmax, y0, x0, ch, cw, rr=rr)
        got = img[y0:y0 + ch, x0:x0 + cw].cpu().double().numpy()
        err = float(np.abs(got - ref).max())
        worst = max(worst, err)
        assert err <= FWD_TOL, f"crop ({y0},{x0}) {ch}x{cw}: max-abs {err:.3e}"
    return worst


def _subset_indices(n, gw, count=96):
    """Gaussians spread over the field, including the first / last grid rows and columns."""
    rng = np.random.default_rng(12)
    gh = n // gw
    idx = list(rng.integers(0, n, count - 16))
    idx += [0, gw - 1, n - gw, n - 1, gw // 2, n - gw // 2, (gh // 2) * gw, (gh // 2) * gw + gw - 1]
    idx += list((gh // 2) * gw + gw // 2 + np.arange(8))
    return np.unique(np.asarray(idx, dtype=np.int64))


def _check_bwd_subset(s, c, k, h, w, dmax, gw):
    """Gradients of a subset of Gaussians against the oracle (their gradients depend on dL/dimg only)."""
    g = torch.rand(h, w, 3, generator=torch.Generator().manual_seed(3))
    got = _bwd(s, c, k, g.to(DEV), dmax)
    idx = _subset_indices(s.shape[0], gw)
    want = oracle.backward(s[idx].cpu().numpy(), c[idx].cpu().numpy(), k[idx].cpu().numpy(), g.numpy(), dmax)
    for a, b, name in zip(got, want, ("sigmas", "coords", "colors")):
        a = a[idx].cpu().double().numpy()
        scale = max(np.abs(b).max(), 1e-12)
        assert np.abs(a - b).max() <= BWD_RTOL * scale, f"grad {name}: {np.abs(a - b).max():.3e} vs scale {scale:.3e}"
    for t in got:
        assert bool(torch.isfinite(t).all())


# ---------------------------------------------------------------- config 3: 512x512 LR -> x8, 4096x4096
@pytest.mark.parametrize("dmax", [0.1, 0.05])
def test_c3_full_shape_forward_crops_and_backward_subset(dmax):
    _, s, c, k, h, w = fields.make("C3", 0)
    assert (h, w, s.shape[0]) == (4096, 4096, 1048576)
    s, c, k = s.to(DEV), c.to(DEV), k.to(DEV)
    img = _fwd(s, c, k, h, w, dmax)
    _check_crops(img, s, c, k, h, w, dmax, 48, 64)
    _check_bwd_subset(s, c, k, h, w, dmax, fields.CONFIGS["C3"].grid[1])


# ---------------------------------------------------------------- headline shape: corners
def test_headline_corner_crops_against_oracle():
    _, s, c, k, h, w = fields.make("HL", 0)
    s, c, k = s.to(DEV), c.to(DEV), k.to(DEV)
    img = _fwd(s, c, k, h, w, 0.1)
    _check_crops(img, s, c, k, h, w, 0.1, 64, 96)
    _check_bwd_subset(s, c, k, h, w, 0.1, fields.CONFIGS["HL"].grid[1])


# ---------------------------------------------------------------- C2d: the real 16 Gaussians / LR pixel density
@needs_ref
def test_c2d_full_forward_backward_against_reference_kernel():
    """1,048,576 Gaussians on 1024x1024 (~170 terms per pixel): whole image and every gradient against the
    reference's own kernels, default k-sigma (the configuration with the smallest truncation margin)."""
    _, s, c, k, h, w = fields.make("C2d", 0)
    s, c, k = s.to(DEV), c.to(DEV), k.to(DEV)
    R = oracle.RefKernels(True)
    ref = R.forward(s, c, k, torch.zeros(h, w, 3, device=DEV), 0.1)
    out = _fwd(s, c, k, h, w, 0.1)
    assert float((out - ref).abs().max()) <= FWD_TOL
    exact = _fwd(s, c, k, h, w, 0.1, ksigma=float("inf"))
    assert float((exact - ref).abs().max()) <= FWD_TOL
    g = torch.rand(h, w, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    rs, rc, rk = R.backward(s, c, k, g, torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k), 0.1)
    for a, b, name in zip(_bwd(s, c, k, g, 0.1), (rs, rc, rk), ("sigmas", "coords", "colors")):
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= BWD_RTOL * scale, f"grad {name}"


def test_c2d_crops_against_oracle():
    _, s, c, k, h, w = fields.make("C2d", 1)
    s, c, k = s.to(DEV), c.to(DEV), k.to(DEV)
    _check_crops(_fwd(s, c, k, h, w, 0.1), s, c, k, h, w, 0.1, 64, 64)


# ---------------------------------------------------------------- densest field per HR pixel: scale ~ 1.5, 16 / LR px
@pytest.mark.parametrize("scale,dmax", [(1.5, 0.1), (1.0, 0.5), (2.0, 0.05)])
def test_low_scale_dense_field_default_ksigma(scale, dmax):
    """The training regime (scale ~ U[1,4], 16 Gaussians per LR pixel, continuous_bicubic_downsample_dataset.py:53-63):
    at scale 1.5 every HR pixel sums ~7 Gaussians per pixel area, sigma_px ~ 0.6 -- the most terms per pixel
    relative to their width.  Whole image against the oracle and the reference kernel at the DEFAULT k."""
    lr = 96
    p = fields.raw_field(lr * 4, lr * 4, seed=int(scale * 10))
    h = w = int(lr * scale)
    s, c, k = fields.map_field(p, h, w, scale)
    ref = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, dmax)
    sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
    out = _fwd(sd, cd, kd, h, w, dmax)
    assert np.abs(out.cpu().double().numpy() - ref).max() <= FWD_TOL
    if oracle.have_ref():
        rk = oracle.RefKernels(True).forward(sd, cd, kd, torch.zeros(h, w, 3, device=DEV), dmax)
        assert float((out - rk).abs().max()) <= FWD_TOL
    g = torch.rand(h, w, 3, generator=torch.Generator().manual_seed(2))
    want = oracle.backward(s.numpy(), c.numpy(), k.numpy(), g.numpy(), dmax)
    for a, b, name in zip(_bwd(sd, cd, kd, g.to(DEV), dmax), want, ("sigmas", "coords", "colors")):
        scale_ = max(np.abs(b).max(), 1e-12)
        assert np.abs(a.cpu().double().numpy() - b).max() <= BWD_RTOL * scale_, f"grad {name}"


# ---------------------------------------------------------------- inclusion set at the largest widths
@needs_ref
@pytest.mark.parametrize("h,w,dmax", [(48, 4096, 0.1), (4096, 40, 0.05), (96, 4096, 0.0137)])
def test_inclusion_counts_bit_identical_at_4096(h, w, dmax):
    """colour = 1, sigma huge: per-pixel COUNTS of contributing Gaussians equal the reference kernel's and the
    oracle's exactly at w = 4096 / h = 4096, where the fp32 rounding of the pixel coordinates (gs.cu:39,46)
    decides pixels on the window edge."""
    rng = np.random.default_rng(h + w)
    n = 2500
    sig = torch.tensor(np.stack([np.full(n, 1e4), np.full(n, 1e4), np.zeros(n)], 1), dtype=torch.float32, device=DEV)
    px = (2.0 * np.arange(w) / (w - 1) - 1.0).astype(np.float32)
    py = (2.0 * np.arange(h) / (h - 1) - 1.0).astype(np.float32)
    xy = rng.uniform(-1.02, 1.02, (n, 2)).astype(np.float32)
    xy[:400, 0] = px[rng.integers(0, w, 400)]
    xy[:400, 1] = py[rng.integers(0, h, 400)]
    xy[400:800, 0] = px[rng.integers(0, w, 400)] + np.float32(dmax)     # window edge exactly on a pixel
    xy[800:1200, 1] = py[rng.integers(0, h, 400)] - np.float32(dmax)
    xyd = torch.tensor(xy, device=DEV)
    col = torch.ones(n, 3, device=DEV)
    ref = oracle.RefKernels(True).forward(sig, xyd, col, torch.zeros(h, w, 3, device=DEV), dmax)
    out = _fwd(sig, xyd, col, h, w, dmax, ksigma=float("inf"))
    _, cnt = oracle.forward(sig.cpu().numpy(), xy, col.cpu().numpy(), h, w, dmax, with_count=True)
    ours = torch.round(out[..., 0]).cpu().numpy().astype(np.int64)
    theirs = torch.round(ref[..., 0]).cpu().numpy().astype(np.int64)
    assert np.array_equal(theirs, cnt), "oracle inclusion set differs from the reference kernel"
    assert np.array_equal(ours, cnt), "inclusion set differs from the reference kernel"


def test_more_gaussians_than_a_bucket_entry_can_index():
    """Maximum sizes: bucket entries hold 23-bit Gaussian indices, so a call with more than 2^23 Gaussians takes the
    home-bin path both ways (gsr_run_tiles raises the overflow flag itself; forward: the cooperative fallback kernel,
    backward: the guarded Gaussian-centric kernel on its strided grid).  Checked through linearity: the render of all
    Gaussians equals the two halves (each below the limit: region path) accumulated into one image, and the
    gradients of the whole call are those of the halves side by side."""
    n, h, w, dmax = (1 << 23) + 4097, 512, 512, 0.1
    g = torch.Generator(DEV).manual_seed(7)
    u = lambda lo, hi, *shape: torch.empty(*shape, device=DEV).uniform_(lo, hi, generator=g)
    s = torch.cat([u(0.004, 0.02, n, 2), u(-0.6, 0.6, n, 1)], 1).contiguous()
    c, k = u(-1.0, 1.0, n, 2), u(0.0, 1e-2, n, 3)
    whole = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(s, c, k, whole, n, h, w, 3, dmax)
    halves = torch.zeros(h, w, 3, device=DEV)
    m = n // 2
    for a, b in ((0, m), (m, n)):   # accumulate contract (gs.cu:58-60)
        gscuda.gs_render(s[a:b].contiguous(), c[a:b].contiguous(), k[a:b].contiguous(), halves, b - a, h, w, 3, dmax)
    torch.cuda.synchronize()
    assert float(halves.max()) > 1.0   # (hundreds of Gaussians per pixel)
    assert float((whole - halves).abs().max()) <= FWD_TOL * float(halves.abs().max())
    grd = u(0.0, 1.0, h, w, 3)
    got = _bwd(s, c, k, grd, dmax)
    for i, (a, b) in enumerate(((0, m), (m, n))):
        part = _bwd(s[a:b].contiguous(), c[a:b].contiguous(), k[a:b].contiguous(), grd, dmax)
        for x, y in zip(got, part):
            assert float((x[a:b] - y).abs().max()) <= 2e-4 * float(y.abs().max())
