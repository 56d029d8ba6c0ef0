"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there), and
the front-end mirror (utils/gaussian_splatting.py) on the GPU.  `-m gpu`."""
import numpy as np
import pytest
import torch

from conftest import golden
from gsasr_b200 import _lib, fields, gscuda
from gsasr_b200 import gaussian_splatting as gsp
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fwd(s, c, k, h, w, dmax, **kw):
    img = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(s, c, k, img, s.shape[0], h, w, 3, dmax, **kw)
    return img


@pytest.fixture(scope="module")
def headline():
    _, s, c, k, h, w = fields.make("HL", 0)
    return s.to(DEV), c.to(DEV), k.to(DEV), h, w


def test_headline_linearity_in_colours(headline):
    """render(a*k1 + k2) == a*render(k1) + render(k2): the image is a plain sum (gs.cu:58-60)."""
    s, c, k, h, w = headline
    k2 = torch.rand_like(k)
    a = _fwd(s, c, 0.5 * k + k2, h, w, 0.1)
    b = 0.5 * _fwd(s, c, k, h, w, 0.1) + _fwd(s, c, k2, h, w, 0.1)
    assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))


def test_headline_chunk_additivity_and_permutation(headline):
    """Any split / order of the Gaussians gives the same image (order-independent sum)."""
    s, c, k, h, w = headline
    whole = _fwd(s, c, k, h, w, 0.1)
    perm = torch.randperm(s.shape[0], device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    shuffled = _fwd(s[perm].contiguous(), c[perm].contiguous(), k[perm].contiguous(), h, w, 0.1)
    assert float((whole - shuffled).abs().max()) <= 2e-5 * max(1.0, float(whole.abs().max()))
    img = torch.zeros(h, w, 3, device=DEV)
    n = s.shape[0]
    for a in range(0, n, 700_001):
        gscuda.gs_render(s[a:a + 700_001], c[a:a + 700_001], k[a:a + 700_001], img, min(700_001, n - a), h, w, 3, 0.1)
    assert float((whole - img).abs().max()) <= 2e-5 * max(1.0, float(whole.abs().max()))


def test_headline_crop_against_oracle(headline):
    """A 96x128 window of the 2048x4096 headline image against the oracle, fed only the Gaussians
    that can reach the window (dmax window semantics make that subset exact)."""
    s, c, k, h, w = headline
    whole = _fwd(s, c, k, h, w, 0.1)
    y0, x0, ch, cw = 1000, 2100, 96, 128
    rr = oracle.ranges(c.cpu().numpy(), h, w, 0.1)
    sel = (rr[:, 0] < x0 + cw) & (rr[:, 1] >= x0) & (rr[:, 2] < y0 + ch) & (rr[:, 3] >= y0)
    idx = np.nonzero(sel)[0]
    # restrict further to Gaussians within 64 px of the window: the rest contribute < 1e-30
    cx = (c[:, 0].cpu().numpy() + 1) * (w - 1) / 2
    cy = (c[:, 1].cpu().numpy() + 1) * (h - 1) / 2
    near = (cx > x0 - 64) & (cx < x0 + cw + 64) & (cy > y0 - 64) & (cy < y0 + ch + 64)
    idx = np.nonzero(sel & near)[0]
    ref = oracle.forward(s[idx].cpu().numpy(), c[idx].cpu().numpy(), k[idx].cpu().numpy(), h, w, 0.1)
    got = whole[y0:y0 + ch, x0:x0 + cw].cpu().double().numpy()
    assert np.abs(got - ref[y0:y0 + ch, x0:x0 + cw]).max() <= 1e-4


def test_headline_backward_linearity_and_colour_grad(headline):
    s, c, k, h, w = headline
    g1 = torch.rand(h, w, 3, device=DEV)
    g2 = torch.rand(h, w, 3, device=DEV)

    def bwd(g):
        gs, gc, gk = torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)
        gscuda.gs_render_backward(s, c, k, g, gs, gc, gk, s.shape[0], h, w, 3, 0.1)
        return gs, gc, gk

    a = bwd(g1 + 2 * g2)
    b1, b2 = bwd(g1), bwd(g2)
    for x, y1, y2 in zip(a, b1, b2):
        y = y1 + 2 * y2
        assert float((x - y).abs().max()) <= 1e-4 * float(y.abs().max())
    # <g, render(k)> == <d/dk, k>  (the image is linear in the colours)
    img = _fwd(s, c, k, h, w, 0.1)
    lhs = float((g1.double() * img.double()).sum())
    rhs = float((b1[2].double() * k.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * abs(lhs)


def test_exact_mode_equals_default_mode_within_budget(headline):
    s, c, k, h, w = headline
    a = _fwd(s, c, k, h, w, 0.1)
    b = _fwd(s, c, k, h, w, 0.1, ksigma=float("inf"))
    assert float((a - b).abs().max()) <= 2e-5


# ---------------------------------------------------------------- front end mirror
@pytest.mark.parametrize("name", ["frontend_x4_fix.npz", "frontend_x2p5_dynamic.npz"])
def test_frontend_mirror_matches_reference_tensors_and_oracle(name):
    g = golden(name)
    h, w, scale = int(g["h"]), int(g["w"]), float(g["scale"])
    raw = torch.tensor(g["raw"], device=DEV)
    out = gsp.generate_2D_gaussian_splatting_step(
        sr_size=torch.tensor([h, w]), gs_parameters=raw, scale=scale,
        scale_modify=torch.tensor([scale, scale]), if_dmax=True, dmax_mode=str(g["dmax_mode"]),
        dmax=float(g["dmax_in"]))
    assert out.shape == (3, h, w) and out.dtype == torch.float32
    ref = oracle.forward(g["sigmas"], g["coords"], g["colors"], h, w, float(g["dmax"]))
    assert np.abs(out.permute(1, 2, 0).cpu().double().numpy() - ref).max() <= 1e-4
    fused = gsp.generate_2D_gaussian_splatting_step(
        sr_size=torch.tensor([h, w]), gs_parameters=raw, scale=scale,
        scale_modify=torch.tensor([scale, scale]), if_dmax=True, dmax_mode=str(g["dmax_mode"]),
        dmax=float(g["dmax_in"]), fused=True)
    assert float((fused - out).abs().max()) <= 1e-4


def test_frontend_backward_matches_torch_autograd_of_the_mapping():
    """d(loss)/d(raw) through the mirror (torch ops + CUDA backward) and through the fused path."""
    p = fields.raw_field(24, 24, seed=3).to(DEV)
    h, w, scale = 50, 46, 2.0
    wgt = torch.rand(3, h, w, device=DEV)

    def run(fused):
        raw = p.clone().requires_grad_(True)
        out = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), raw, scale,
                                                      torch.tensor([scale, scale]), dmax=0.2, fused=fused)
        (out * wgt).sum().backward()
        return out.detach(), raw.grad.detach()

    o1, g1 = run(False)
    o2, g2 = run(True)
    assert float((o1 - o2).abs().max()) <= 1e-4
    assert float((g1 - g2).abs().max()) <= 1e-3 * float(g1.abs().max())
    # oracle check of the mirror path: analytic grads of the raster, chained by torch autograd
    s, c, k = fields.map_field(p.cpu(), h, w, scale)
    gs, gc, gk = oracle.backward(s.numpy(), c.numpy(), k.numpy(), wgt.permute(1, 2, 0).cpu().numpy(), 0.2)
    raw = p.cpu().clone().requires_grad_(True)
    s2, c2, k2 = fields.map_field(raw, h, w, scale)
    (s2 * torch.tensor(gs, dtype=torch.float32)).sum().add((c2 * torch.tensor(gc, dtype=torch.float32)).sum()).add(
        (k2 * torch.tensor(gk, dtype=torch.float32)).sum()).backward()
    assert float((g1.cpu() - raw.grad).abs().max()) <= 1e-3 * float(raw.grad.abs().max())


def test_buffer_variant_equals_unbuffered_and_differentiates():
    p = fields.raw_field(40, 40, seed=4).to(DEV)
    h, w, scale = 80, 80, 2.0
    args = dict(sr_size=torch.tensor([h, w]), scale=scale, scale_modify=torch.tensor([scale, scale]), dmax=0.1)
    raw1 = p.clone().requires_grad_(True)
    a = gsp.generate_2D_gaussian_splatting_step(gs_parameters=raw1, **args)
    raw2 = p.clone().requires_grad_(True)
    b = gsp.generate_2D_gaussian_splatting_step_buffer(gs_parameters=raw2, buffer_size=500, **args)
    assert float((a - b).abs().max()) <= 2e-5
    a.sum().backward()
    b.sum().backward()
    assert float((raw1.grad - raw2.grad).abs().max()) <= 1e-4 * float(raw1.grad.abs().max())


@pytest.mark.parametrize("bgr", [False, True])
def test_fused_uint8_output_equals_the_reference_post_processing(bgr):
    """GSR_FLAG_U8: clamp_(0,1) -> [2,1,0] -> HWC -> *255 -> round -> uint8 (inference_paper.py:136-138) inside
    the raster kernel, against the same chain applied with numpy to our fp32 render."""
    from gsasr_b200 import gaussian_splatting as gsp

    _, s, c, k, h, w = fields.make("C1", 3)
    k = k * 3.0 - 0.2                                   # values on both sides of the clamp
    sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
    img = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(sd, cd, kd, img, s.shape[0], h, w, 3, 0.1)
    ref = img.cpu().clamp_(0, 1).numpy()
    if bgr:
        ref = ref[:, :, [2, 1, 0]]
    ref = (ref * 255.0).round().astype(np.uint8)
    out = torch.full((h, w, 3), 77, dtype=torch.uint8, device=DEV)
    gscuda.gs_render_u8(sd, cd, kd, out, s.shape[0], h, w, 0.1, bgr=bgr)
    got = out.cpu().numpy()
    diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3   # only fp32 round-off at a rounding boundary
    assert got.min() == 0 and got.max() == 255
    raw = fields.raw_field(32, 32, seed=1).to(DEV)
    u8 = gsp.generate_2D_gaussian_splatting_step_u8(torch.tensor([64, 64]), raw, 2.0, torch.tensor([2.0, 2.0]), dmax=0.1, bgr=bgr)
    f32 = gsp.generate_2D_gaussian_splatting_step(torch.tensor([64, 64]), raw, 2.0, torch.tensor([2.0, 2.0]), dmax=0.1)
    want = f32.cpu().clamp_(0, 1).numpy()
    want = np.transpose(want[[2, 1, 0]] if bgr else want, (1, 2, 0))
    want = (want * 255.0).round().astype(np.uint8)
    assert np.abs(u8.cpu().numpy().astype(np.int16) - want.astype(np.int16)).max() <= 1
    with pytest.raises(RuntimeError):                    # uint8 output cannot be accumulated into
        L = _lib.load()
        ws = gscuda.workspace(s.shape[0], h, w, DEV)
        _lib.check(L.gsr_forward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), out.data_ptr(), s.shape[0], h, w, 3,
                                 0.1, 0.0, _lib.GSR_FLAG_U8, ws.data_ptr(), ws.numel(), 0))


def test_deterministic_mode_is_bit_reproducible():
    """GSR_FLAG_DETERMINISTIC: region lists sorted by Gaussian index -> the fp32 summation order, and with it every
    bit of the image, is a function of the inputs alone (the reference's atomicAdd accumulation, gs.cu:58-60, is
    not reproducible).  Checked on a dense field (C2d-like density: hundreds of entries per list), on an image with
    partial edge regions, and against the default mode to summation-order tolerance."""
    for name, seed in (("C2", 0), ("C1", 3)):
        _, s, c, k, h, w = fields.make(name, seed)
        sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
        imgs = []
        for rep in range(3):
            img = torch.zeros(h, w, 3, device=DEV)
            gscuda.gs_render(sd, cd, kd, img, s.shape[0], h, w, 3, 0.1, flags=0x20)
            imgs.append(img)
        torch.cuda.synchronize()
        assert torch.equal(imgs[0], imgs[1]) and torch.equal(imgs[0], imgs[2])
        ref = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(sd, cd, kd, ref, s.shape[0], h, w, 3, 0.1)
        assert float((ref - imgs[0]).abs().max()) <= 1e-5
    # odd sizes (partial regions), huge lists (every Gaussian covers the image: the long-bucket sort)
    rng = np.random.default_rng(5)
    n, h, w = 3000, 37, 53
    sg = torch.tensor(np.stack([rng.uniform(0.5, 2, n), rng.uniform(0.5, 2, n), rng.uniform(-0.5, 0.5, n)], 1), dtype=torch.float32, device=DEV)
    xy = torch.tensor(rng.uniform(-1, 1, (n, 2)), dtype=torch.float32, device=DEV)
    col = torch.tensor(rng.uniform(0, 1e-3, (n, 3)), dtype=torch.float32, device=DEV)
    outs = []
    for rep in range(2):
        img = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(sg, xy, col, img, n, h, w, 3, flags=0x20)
        outs.append(img)
    assert torch.equal(outs[0], outs[1])
    try:
        gscuda.set_deterministic(True)
        a = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(sg, xy, col, a, n, h, w, 3)
        assert torch.equal(a, outs[0])
        # ... and gradients: the flag selects the Gaussian-centric backward (no atomics)
        _, s, c, k, h, w = fields.make("C2", 1)
        sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
        g = torch.rand(h, w, 3, device=DEV)
        runs = []
        for rep in range(3):
            out = [torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)]
            gscuda.gs_render_backward(sd, cd, kd, g, *out, s.shape[0], h, w, 3, 0.1)
            runs.append(out)
        for x, y, z in zip(*runs):
            assert torch.equal(x, y) and torch.equal(x, z)
    finally:
        gscuda.set_deterministic(False)
    fast = [torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)]
    gscuda.gs_render_backward(sd, cd, kd, g, *fast, s.shape[0], h, w, 3, 0.1)  # default: over the region buckets
    for x, f in zip(runs[0], fast):
        assert float((f - x).abs().max()) <= 2e-4 * float(x.abs().max())  # fp32 summation order (sigma gradients cancel)


def test_calls_are_cuda_graph_capturable():
    """no allocation, no synchronisation, no host-side data dependence inside the library, so a
    forward (incl. the cooperatively launched fallback kernel and the deterministic sort) and a backward call can be
    captured in a CUDA graph and replayed on new parameter values."""
    _, s, c, k, h, w = fields.make("C1", 2)
    sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
    n = s.shape[0]
    ws = gscuda.workspace(n, h, w, torch.device(DEV))
    img = torch.zeros(h, w, 3, device=DEV)
    g = torch.rand(h, w, 3, device=DEV)
    gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up outside the capture (lazy module loading, attribute opt-ins)
        gscuda.gs_render(sd, cd, kd, img, n, h, w, 3, 0.1, flags=0x1 | 0x20, workspace_buf=ws)
        gscuda.gs_render_backward(sd, cd, kd, g, gs, gc, gk, n, h, w, 3, 0.1, flags=0x20, workspace_buf=ws)
        gscuda.gs_render_backward(sd, cd, kd, g, gs, gc, gk, n, h, w, 3, 0.1, workspace_buf=ws)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    gs.zero_(), gc.zero_(), gk.zero_()
    fast = [torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)]  # the default (region) backward
    with torch.cuda.graph(graph):
        gscuda.gs_render(sd, cd, kd, img, n, h, w, 3, 0.1, flags=0x1 | 0x20, workspace_buf=ws)
        gscuda.gs_render_backward(sd, cd, kd, g, gs, gc, gk, n, h, w, 3, 0.1, flags=0x20, workspace_buf=ws)
        gscuda.gs_render_backward(sd, cd, kd, g, *fast, n, h, w, 3, 0.1, workspace_buf=ws)
    kd.mul_(0.5)  # new values in the captured buffers
    gs.zero_(), gc.zero_(), gk.zero_()
    for t_ in fast:
        t_.zero_()
    graph.replay()
    torch.cuda.synchronize()
    want = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(sd, cd, kd, want, n, h, w, 3, 0.1, flags=0x1 | 0x20)
    ws2 = [torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)]
    gscuda.gs_render_backward(sd, cd, kd, g, *ws2, n, h, w, 3, 0.1, flags=0x20)
    assert torch.equal(img, want)              # deterministic mode: bit-identical, replayed or not
    for a, b, f in zip((gs, gc, gk), ws2, fast):
        assert torch.equal(a, b)               # (Gaussian-centric backward: one writer per output, fixed order)
        assert float((f - b).abs().max()) <= 2e-4 * float(b.abs().max()) + 1e-12  # region backward: summation order


def test_autograd_backward_reuses_the_forward_setup():
    """gswrapper.GSCUDA / gaussian_splatting.render_chw keep the forward's workspace and run gsr_backward_prepared on it
    (gscuda.set_reuse_setup): same gradients as the full backward call, a second backward through a retained graph
    (workspace already released) included; clustered field: the buckets overflow and both passes take the fallback."""
    from gsasr_b200 import gaussian_splatting as gsp
    from gsasr_b200.gswrapper import gaussiansplatting_render
    rng = np.random.default_rng(11)
    cases = [fields.make("C1", 4)[1:], None]
    n, h, w = 4000, 128, 128   # every Gaussian near the centre: bucket overflow -> home-bin fallback both ways
    cases[1] = (torch.tensor(np.stack([np.full(n, 0.02), np.full(n, 0.02), np.zeros(n)], 1), dtype=torch.float32),
                torch.tensor(rng.normal(0, 0.01, (n, 2)), dtype=torch.float32), torch.rand(n, 3), h, w)
    for s, c, k, h, w in cases:
        wgt = torch.rand(h, w, 3, device=DEV)
        grads = {}
        for reuse in (True, False):
            gscuda.set_reuse_setup(reuse)
            try:
                leaves = [t.to(DEV).requires_grad_(True) for t in (s, c, k)]
                img = gaussiansplatting_render(*leaves, (h, w), 0.5)
                (img * wgt).sum().backward(retain_graph=True)
                first = [t.grad.clone() for t in leaves]
                (img * wgt).sum().backward()   # accumulates: twice the gradient
                for t, f in zip(leaves, first):
                    assert float((t.grad - 2 * f).abs().max()) <= 2e-4 * float(f.abs().max()) + 1e-12
                grads[reuse] = first
                lv2 = [t.detach().clone().requires_grad_(True) for t in leaves]
                (gsp.render_chw(*lv2, h, w, 0.5) * wgt.permute(2, 0, 1)).sum().backward()
                for t, f in zip(lv2, first):
                    assert float((t.grad - f).abs().max()) <= 2e-4 * float(f.abs().max()) + 1e-12
            finally:
                gscuda.set_reuse_setup(True)
        for a, b in zip(grads[True], grads[False]):
            assert float((a - b).abs().max()) <= 2e-4 * float(b.abs().max()) + 1e-12



@pytest.mark.parametrize("flags", [0, 0x20])
def test_non_finite_gradient_pixel_stays_inside_its_dmax_windows(flags):
    """One inf in dL/dimg (an AMP overflow step): the reference sums a Gaussian's gradient over its dmax window only
    (gs.cu:112-131), so Gaussians whose window does not contain that pixel keep finite gradients.  Both backward
    kernels evaluate pixels beyond the window (whole cells / whole patches) and must not turn 0 * inf into NaN."""
    rng = np.random.default_rng(21)
    n, h, w, dmax = 600, 64, 64, 0.12
    s = torch.tensor(np.stack([rng.uniform(0.02, 0.2, n), rng.uniform(0.02, 0.2, n), rng.uniform(-0.7, 0.7, n)], 1),
                     dtype=torch.float32, device=DEV)
    c = torch.tensor(rng.uniform(-1, 1, (n, 2)), dtype=torch.float32, device=DEV)
    k = torch.rand(n, 3, device=DEV)
    g = torch.rand(h, w, 3, device=DEV)
    py, px = 21, 38
    g[py, px, 1] = float("inf")
    out = [torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)]
    gscuda.gs_render_backward(s, c, k, g, *out, n, h, w, 3, dmax, flags=flags)
    torch.cuda.synchronize()
    xs, ys = 2.0 * px / (w - 1) - 1.0, 2.0 * py / (h - 1) - 1.0   # pixel centre in the reference's coordinates
    far = ((c[:, 0] - xs).abs() > dmax + 1e-3) | ((c[:, 1] - ys).abs() > dmax + 1e-3)
    assert int(far.sum()) > n // 2 and int((~far).sum()) > 3
    for t in out:
        assert bool(torch.isfinite(t[far]).all())
