"""Uniform batches (gsr_forward_batch_uniform / gsr_backward_batch_uniform): B samples of one shape rendered
as one stacked image must equal B single-sample calls.  Needs a GPU: `-m gpu`."""
import numpy as np
import pytest
import torch

from gsasr_b200 import _lib, fields, gscuda
from gsasr_b200.gswrapper import gaussiansplatting_render, gaussiansplatting_render_batch
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(b, gh, gw, h, w, scale):
    ss, cc, kk = [], [], []
    for i in range(b):
        s, c, k = fields.map_field(fields.raw_field(gh, gw, seed=10 + i), h, w, scale)
        ss.append(s); cc.append(c); kk.append(k)
    return tuple(torch.stack(t).to(DEV).contiguous() for t in (ss, cc, kk))


@pytest.mark.parametrize("b,h,w,scale,dmax", [
    (4, 64, 48, 2.0, 0.1),     # stacked: 4 x 64 rows
    (3, 40, 72, 4.0, 0.05),    # stacked, windows bind, samples differ
    (5, 36, 40, 2.0, 0.3),     # h % 8 != 0: one call per sample
    (1, 32, 32, 2.0, 0.1),
])
def test_batch_equals_single_sample_calls(b, h, w, scale, dmax):
    s, c, k = _batch(b, 24, 28, h, w, scale)
    n = s.shape[1]
    imgs = torch.zeros(b, h, w, 3, device=DEV)
    gscuda.gs_render_batch(s, c, k, imgs, dmax)
    g = torch.rand(b, h, w, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    gs, gc, gk = torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)
    gscuda.gs_render_backward_batch(s, c, k, g, gs, gc, gk, dmax)
    for i in range(b):
        one = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(s[i], c[i], k[i], one, n, h, w, 3, dmax)
        assert float((imgs[i] - one).abs().max()) <= 2e-6
        ref = oracle.forward(s[i].cpu().numpy(), c[i].cpu().numpy(), k[i].cpu().numpy(), h, w, dmax)
        assert np.abs(imgs[i].cpu().double().numpy() - ref).max() <= 1e-4
        ws = [torch.zeros_like(t[i]) for t in (s, c, k)]
        gscuda.gs_render_backward(s[i], c[i], k[i], g[i].contiguous(), *ws, n, h, w, 3, dmax)
        for a, w_ in zip((gs[i], gc[i], gk[i]), ws):
            assert float((a - w_).abs().max()) <= 1e-5 * float(w_.abs().max())


def test_batch_stack_is_cut_at_the_maximum_image_height():
    """h = 4096: at most 7 samples fit a 32767-row stack, so 9 samples take two launches (5 + 4)."""
    L = _lib.load()
    b, h, w, n = 9, 4096, 16, 50
    rng = np.random.default_rng(0)
    s = torch.tensor(np.stack([rng.uniform(0.05, 0.2, (b, n)), rng.uniform(2e-4, 2e-3, (b, n)),
                               np.zeros((b, n))], 2), dtype=torch.float32, device=DEV)
    c = torch.tensor(rng.uniform(-1, 1, (b, n, 2)), dtype=torch.float32, device=DEV)
    k = torch.tensor(rng.uniform(0, 1, (b, n, 3)), dtype=torch.float32, device=DEV)
    imgs = torch.zeros(b, h, w, 3, device=DEV)
    gscuda.gs_render_batch(s, c, k, imgs, 0.5)
    for i in (0, 6, 7, 8):
        one = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(s[i], c[i], k[i], one, n, h, w, 3, 0.5)
        assert float((imgs[i] - one).abs().max()) <= 2e-6
    assert L.gsr_workspace_bytes_batch_uniform(9, n, 4096, 16) >= L.gsr_workspace_bytes(5 * n, 5 * 4096, 16)


def test_batch_autograd_matches_per_sample_autograd():
    b, h, w = 3, 48, 56
    s, c, k = _batch(b, 20, 24, h, w, 2.0)
    leaves = [t.clone().requires_grad_(True) for t in (s, c, k)]
    out = gaussiansplatting_render_batch(*leaves, (h, w), 0.2)
    wgt = torch.rand(b, h, w, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    (out * wgt).sum().backward()
    for i in range(b):
        li = [t[i].clone().requires_grad_(True) for t in (s, c, k)]
        oi = gaussiansplatting_render(*li, (h, w), 0.2)
        (oi * wgt[i]).sum().backward()
        assert float((out[i] - oi).abs().max()) <= 2e-6
        for a, bb in zip(leaves, li):
            assert float((a.grad[i] - bb.grad).abs().max()) <= 1e-5 * float(bb.grad.abs().max())


@pytest.mark.parametrize("fused", [False, True])
def test_batched_front_end_matches_per_sample_front_end(fused):
    from gsasr_b200 import gaussian_splatting as gsp

    b, h, w, scale = 3, 64, 64, 4.0
    raw = torch.stack([fields.raw_field(16, 16, seed=20 + i) for i in range(b)]).to(DEV)
    wgt = torch.rand(b, 3, h, w, device=DEV, generator=torch.Generator(DEV).manual_seed(5))
    pb = raw.clone().requires_grad_(True)
    out = gsp.generate_2D_gaussian_splatting_step_batch(torch.tensor([h, w]), pb, scale, torch.tensor([scale] * 2),
                                                        dmax=0.1, fused=fused)
    assert tuple(out.shape) == (b, 3, h, w)
    (out * wgt).sum().backward()
    for i in range(b):
        pi = raw[i].clone().requires_grad_(True)
        oi = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), pi, scale, torch.tensor([scale] * 2),
                                                     dmax=0.1, fused=fused)
        (oi * wgt[i]).sum().backward()
        assert float((out[i].detach() - oi.detach()).abs().max()) <= 2e-6
        assert float((pb.grad[i] - pi.grad).abs().max()) <= 1e-5 * float(pi.grad.abs().max())


# ---- padded (ragged) batches: every sample at its own size, one launch ------------------------------
def _ragged(sizes, gh, gw, scales):
    ss, cc, kk = [], [], []
    for i, ((h, w), sc) in enumerate(zip(sizes, scales)):
        s, c, k = fields.map_field(fields.raw_field(gh, gw, seed=40 + i), h, w, sc)
        ss.append(s); cc.append(c); kk.append(k)
    return tuple(torch.stack(t).to(DEV).contiguous() for t in (ss, cc, kk))


@pytest.mark.parametrize("sizes,scales,dmax", [
    ([(64, 48), (40, 72), (61, 33), (16, 80)], [2.0, 2.5, 2.0, 3.0], 0.1),
    ([(50, 50), (48, 64), (33, 37)], [2.0, 2.0, 1.5], [0.05, 0.3, 0.02]),   # per-sample dmax, windows bind
    ([(24, 24)], [1.0], 0.2),
])
def test_padded_batch_equals_per_sample_render_plus_padding(sizes, scales, dmax):
    s, c, k = _ragged(sizes, 24, 28, scales)
    b, n = s.shape[:2]
    hmax = (max(h for h, _ in sizes) + 7) // 8 * 8
    wmax = max(w for _, w in sizes) + 3
    imgs = torch.full((b, hmax, wmax, 3), 9.0, device=DEV)
    gscuda.gs_render_batch_padded(s, c, k, imgs, sizes, dmax, flags=1)
    g = torch.rand(b, hmax, wmax, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(2))
    gs, gc, gk = torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)
    gscuda.gs_render_backward_batch_padded(s, c, k, g, gs, gc, gk, sizes, dmax)
    for i, (h, w) in enumerate(sizes):
        dm = dmax if isinstance(dmax, float) else dmax[i]
        one = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(s[i], c[i], k[i], one, n, h, w, 3, dm)
        assert float((imgs[i, :h, :w] - one).abs().max()) <= 2e-5      # fp32 round-off of the rescaled records
        ref = oracle.forward(s[i].cpu().numpy(), c[i].cpu().numpy(), k[i].cpu().numpy(), h, w, dm)
        assert np.abs(imgs[i, :h, :w].cpu().double().numpy() - ref).max() <= 1e-4
        pad = imgs[i].clone()
        pad[:h, :w] = 0
        assert float(pad.abs().max()) == 0.0                             # the padding is written 0, exactly
        ws = [torch.zeros_like(t[i]) for t in (s, c, k)]
        gscuda.gs_render_backward(s[i], c[i], k[i], g[i, :h, :w].contiguous(), *ws, n, h, w, 3, dm)
        for a, w_ in zip((gs[i], gc[i], gk[i]), ws):
            assert float((a - w_).abs().max()) <= 2e-4 * float(w_.abs().max())


def test_padded_batch_inclusion_counts_are_exact():
    """Colours 1, huge sigma: the rounded images are per-pixel inclusion counts of each sample's OWN dmax
    window on its OWN pixel grid -- compared exactly with the oracle."""
    sizes, n, dmax = [(45, 61), (64, 40), (23, 77)], 200, 0.13
    rng = np.random.default_rng(8)
    s = torch.tensor(np.broadcast_to(np.array([1e3, 1e3, 0.0]), (3, n, 3)).copy(), dtype=torch.float32, device=DEV)
    c = torch.tensor(rng.uniform(-1.05, 1.05, (3, n, 2)), dtype=torch.float32, device=DEV)
    k = torch.ones(3, n, 3, device=DEV)
    imgs = torch.zeros(3, 64, 80, 3, device=DEV)
    gscuda.gs_render_batch_padded(s, c, k, imgs, sizes, dmax, ksigma=float("inf"))
    for i, (h, w) in enumerate(sizes):
        cnt = np.zeros((h, w))
        for x0, x1, y0, y1 in oracle.ranges(c[i].cpu().numpy(), h, w, dmax):
            if x1 >= x0 and y1 >= y0:
                cnt[y0:y1 + 1, x0:x1 + 1] += 1
        assert np.array_equal(np.rint(imgs[i, :h, :w, 0].cpu().numpy()), cnt)


def test_padded_batch_autograd():
    from gsasr_b200.gswrapper import gaussiansplatting_render_batch_padded

    sizes, scales = [(40, 56), (48, 32)], [2.0, 2.0]
    s, c, k = _ragged(sizes, 20, 24, scales)
    leaves = [t.clone().requires_grad_(True) for t in (s, c, k)]
    out = gaussiansplatting_render_batch_padded(*leaves, sizes, 0.2)
    assert tuple(out.shape) == (2, 48, 56, 3)
    wgt = torch.rand_like(out)
    (out * wgt).sum().backward()
    for i, (h, w) in enumerate(sizes):
        li = [t[i].clone().requires_grad_(True) for t in (s, c, k)]
        oi = gaussiansplatting_render(*li, (h, w), 0.2)
        (oi * wgt[i, :h, :w]).sum().backward()
        for a, bb in zip(leaves, li):
            assert float((a.grad[i] - bb.grad).abs().max()) <= 2e-4 * float(bb.grad.abs().max())


@pytest.mark.parametrize("fused", [False, True])
def test_padded_front_end_matches_the_per_sample_training_loop(fused):
    """generate_2D_gaussian_splatting_step_batch_padded against the loop of gsasr_model.py:191-233:
    per-sample render at its own scale / size, F.pad to the largest, values and gradients."""
    import torch.nn.functional as F

    from gsasr_b200 import gaussian_splatting as gsp

    scales = [2.0, 3.0, 2.5]
    sizes = [(32, 32), (48, 48), (40, 40)]                 # 16x16 LR at each scale
    raw = torch.stack([fields.raw_field(32, 32, seed=60 + i) for i in range(3)]).to(DEV)
    pb = raw.clone().requires_grad_(True)
    out = gsp.generate_2D_gaussian_splatting_step_batch_padded([torch.tensor(s) for s in sizes], pb, scales, dmax=0.1,
                                                               fused=fused)
    assert tuple(out.shape) == (3, 3, 48, 48)
    wgt = torch.rand(3, 3, 48, 48, device=DEV, generator=torch.Generator(DEV).manual_seed(4))
    (out * wgt).sum().backward()
    for i, ((h, w), sc) in enumerate(zip(sizes, scales)):
        pi = raw[i].clone().requires_grad_(True)
        oi = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), pi, sc, torch.tensor([sc, sc]), dmax=0.1,
                                                     fused=fused)
        oi = F.pad(oi, (0, 48 - w, 0, 48 - h), 'constant', 0)
        (oi * wgt[i]).sum().backward()
        assert float((out[i].detach() - oi.detach()).abs().max()) <= 2e-5
        assert float((pb.grad[i] - pi.grad).abs().max()) <= 2e-4 * float(pi.grad.abs().max())
