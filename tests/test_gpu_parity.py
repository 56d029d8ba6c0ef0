"""Parity of the sm_100a kernels (through the C ABI) against the oracle, the committed golden
vectors and the reference's own CUDA kernels (oracle/_ref).  Needs a GPU: `-m gpu`.

Tolerances (BASELINE.json north_star): forward <= 1e-4 abs fp32 against the reference kernel and
against the fp64 oracle; the dmax inclusion set is compared EXACTLY; backward <= 1e-3 relative to
the largest gradient of each tensor.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from gsasr_b200 import _lib, fields, gscuda
from gsasr_b200.gswrapper import GSCUDA, gaussiansplatting_render
from oracle import oracle

pytestmark = pytest.mark.gpu
FWD_TOL = 1e-4
BWD_RTOL = 1e-3
DEV = "cuda:0"


def _render(s, c, k, h, w, dmax, ksigma=None, init=None, flags=0):
    s, c, k = (torch.as_tensor(a, dtype=torch.float32, device=DEV).contiguous() for a in (s, c, k))
    if flags & _lib.GSR_FLAG_CHW:
        img = torch.zeros(3, h, w, device=DEV) if init is None else init
    else:
        img = torch.zeros(h, w, 3, device=DEV) if init is None else init
    gscuda.gs_render(s, c, k, img, s.shape[0], h, w, 3, dmax, ksigma=ksigma, flags=flags)
    torch.cuda.synchronize()
    return img


def _backward(s, c, k, g, dmax, ksigma=None, flags=0):
    s, c, k, g = (torch.as_tensor(a, dtype=torch.float32, device=DEV).contiguous() for a in (s, c, k, g))
    gs, gc, gk = torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)
    h, w = (g.shape[1], g.shape[2]) if flags & _lib.GSR_FLAG_CHW else (g.shape[0], g.shape[1])
    gscuda.gs_render_backward(s, c, k, g, gs, gc, gk, s.shape[0], h, w, 3, dmax, ksigma=ksigma, flags=flags)
    torch.cuda.synchronize()
    return gs.cpu().double().numpy(), gc.cpu().double().numpy(), gk.cpu().double().numpy()


def _assert_grads(got, want, rtol=BWD_RTOL):
    for a, b, name in zip(got, want, ("sigmas", "coords", "colors")):
        scale = max(np.abs(b).max(), 1e-12)
        assert np.abs(a - b).max() <= rtol * scale, f"grad {name}: {np.abs(a - b).max():.3e} vs scale {scale:.3e}"


# ---------------------------------------------------------------- golden vectors
CHECK_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "tests/golden/check_*.npz")))


@pytest.mark.parametrize("name", CHECK_FILES)
@pytest.mark.parametrize("ksigma", [None, float("inf")])
def test_golden_check_py_forward_backward(name, ksigma):
    g = golden(name)
    h, w, dmax = int(g["h"]), int(g["w"]), float(g["dmax"])
    img = _render(g["sigmas"], g["coords"], g["colors"], h, w, dmax, ksigma).cpu().numpy()
    assert np.abs(img - g["img"]).max() <= FWD_TOL
    got = _backward(g["sigmas"], g["coords"], g["colors"], g["weight"], dmax, ksigma)
    _assert_grads(got, (g["g_sigmas"], g["g_coords"], g["g_colors"]))


# ---------------------------------------------------------------- oracle, seeded fields
@pytest.mark.parametrize("cfg,dmax,seed", [("C1", 0.1, 0), ("C1", 0.05, 1), ("C1", 0.1, 2), ("C1", 0.02, 3)])
def test_forward_matches_oracle_c1(cfg, dmax, seed):
    _, s, c, k, h, w = fields.make(cfg, seed)
    ref = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, dmax)
    for ks in (None, float("inf")):
        out = _render(s, c, k, h, w, dmax, ks).cpu().double().numpy()
        assert np.abs(out - ref).max() <= FWD_TOL


def test_forward_backward_match_oracle_c2():
    """BASELINE config 2: 256x256 LR -> x4, 262,144 Gaussians, fwd + bwd."""
    _, s, c, k, h, w = fields.make("C2", 0)
    ref = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, 0.1)
    out = _render(s, c, k, h, w, 0.1).cpu().double().numpy()
    assert np.abs(out - ref).max() <= FWD_TOL
    g = torch.rand(h, w, 3, generator=torch.Generator().manual_seed(5))
    want = oracle.backward(s.numpy(), c.numpy(), k.numpy(), g.numpy(), 0.1)
    _assert_grads(_backward(s, c, k, g, 0.1), want)


@pytest.mark.parametrize("h,w", [(33, 70), (97, 31), (2, 2), (2, 129), (300, 3)])
def test_ragged_sizes_and_partial_tiles(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    n = 400
    sig = np.stack([rng.uniform(0.01, 0.3, n), rng.uniform(0.01, 0.3, n), np.tanh(rng.normal(0, 1, n)) * 0.999], 1)
    xy = rng.uniform(-1.2, 1.2, (n, 2))
    col = rng.uniform(0, 1, (n, 3))
    for dmax in (0.15, float("inf")):
        ref = oracle.forward(sig, xy, col, h, w, dmax)
        out = _render(sig, xy, col, h, w, dmax, float("inf")).cpu().double().numpy()
        assert np.abs(out - ref).max() <= FWD_TOL * max(1.0, np.abs(ref).max())
        g = rng.uniform(-1, 1, (h, w, 3))
        _assert_grads(_backward(sig, xy, col, g, dmax, float("inf")),
                      oracle.backward(sig, xy, col, g.astype(np.float32), dmax))


def test_empty_and_degenerate_inputs():
    h, w = 40, 56
    z3 = torch.zeros(0, 3, device=DEV)
    img = _render(z3, torch.zeros(0, 2, device=DEV), z3, h, w, 0.1)
    assert float(img.abs().max()) == 0.0
    # invalid Gaussians are skipped; valid neighbours are unaffected
    sig = np.array([[0.1, 0.1, 0.0], [0.0, 0.1, 0.0], [0.1, 0.1, 1.0], [np.nan, 0.1, 0.0], [0.1, 0.1, 0.2]], np.float32)
    xy = np.array([[0, 0], [0, 0], [0, 0], [0, 0], [0.5, -0.5]], np.float32)
    col = np.ones((5, 3), np.float32)
    ref = oracle.forward(sig[[0, 4]], xy[[0, 4]], col[[0, 4]], h, w, 0.3)
    out = _render(sig, xy, col, h, w, 0.3).cpu().double().numpy()
    assert np.isfinite(out).all() and np.abs(out - ref).max() <= FWD_TOL
    gs, gc, gk = _backward(sig, xy, col, np.ones((h, w, 3), np.float32), 0.3)
    assert np.isfinite(gs).all() and np.all(gs[1:4] == 0) and np.all(gc[1:4] == 0) and np.all(gk[1:4] == 0)


def test_large_gaussians_take_the_large_list():
    """sigma = 5 (check.py's range) on a 300x420 image: every Gaussian covers everything."""
    rng = np.random.default_rng(11)
    n, h, w = 64, 300, 420
    sig = np.stack([rng.uniform(0.5, 5, n), rng.uniform(0.5, 5, n), rng.uniform(-0.9, 0.9, n)], 1)
    xy = rng.uniform(-1, 1, (n, 2))
    col = rng.uniform(0, 1, (n, 3)) / n
    for dmax in (0.5, float("inf")):
        ref = oracle.forward(sig, xy, col, h, w, dmax)
        out = _render(sig, xy, col, h, w, dmax).cpu().double().numpy()
        assert np.abs(out - ref).max() <= FWD_TOL
        g = rng.uniform(-1, 1, (h, w, 3)).astype(np.float32)
        _assert_grads(_backward(sig, xy, col, g, dmax), oracle.backward(sig, xy, col, g, dmax))


def test_accumulate_overwrite_and_chw_flags():
    _, s, c, k, h, w = fields.make("C1", 4)
    base = torch.full((h, w, 3), 0.25, device=DEV)
    plain = _render(s, c, k, h, w, 0.1)
    acc = _render(s, c, k, h, w, 0.1, init=base.clone())
    assert torch.allclose(acc, plain + 0.25, atol=1e-6)                       # accumulates (gs.cu:58-60)
    over = _render(s, c, k, h, w, 0.1, init=base.clone(), flags=_lib.GSR_FLAG_OVERWRITE)
    assert torch.allclose(over, plain, atol=1e-6)   # (summation order varies run to run, as in the reference)
    chw = _render(s, c, k, h, w, 0.1, flags=_lib.GSR_FLAG_CHW | _lib.GSR_FLAG_OVERWRITE,
                  init=torch.full((3, h, w), 7.0, device=DEV))
    assert torch.allclose(chw, plain.permute(2, 0, 1).contiguous(), atol=1e-6)
    g = torch.rand(h, w, 3, device=DEV)
    a = _backward(s, c, k, g, 0.1)
    b = _backward(s, c, k, g.permute(2, 0, 1).contiguous(), 0.1, flags=_lib.GSR_FLAG_CHW)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-5 * np.abs(x).max())


def test_chunked_calls_accumulate_like_the_buffer_path():
    """rendering_cuda_dmax_buffer (gaussian_splatting.py:146-151) re-applies into the same image."""
    _, s, c, k, h, w = fields.make("C1", 5)
    whole = _render(s, c, k, h, w, 0.1)
    img = torch.zeros(h, w, 3, device=DEV)
    for a in range(0, s.shape[0], 5000):
        img = _render(s[a:a + 5000], c[a:a + 5000], k[a:a + 5000], h, w, 0.1, init=img)
    assert torch.allclose(img, whole, atol=2e-5)


# ---------------------------------------------------------------- the reference's own kernels
needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")


@needs_ref
@pytest.mark.parametrize("cfg,dmax,seed", [("C1", 0.1, 0), ("C1", 0.05, 1), ("C2", 0.1, 0)])
def test_forward_backward_match_reference_kernel(cfg, dmax, seed):
    _, s, c, k, h, w = fields.make(cfg, seed)
    sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
    R = oracle.RefKernels(True)
    ref = R.forward(sd, cd, kd, torch.zeros(h, w, 3, device=DEV), dmax)
    out = _render(sd, cd, kd, h, w, dmax)
    assert float((out - ref).abs().max()) <= FWD_TOL
    g = torch.rand(h, w, 3, device=DEV)
    rs, rc, rk = R.backward(sd, cd, kd, g, torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd), dmax)
    got = _backward(sd, cd, kd, g, dmax)
    _assert_grads(got, tuple(t.cpu().double().numpy() for t in (rs, rc, rk)))


@needs_ref
def test_no_dmax_variant_matches_gs_cuda():
    """utils/gs_cuda (no window) == dmax=+inf."""
    rng = np.random.default_rng(2)
    n, h, w = 300, 49, 49
    sig = torch.tensor(np.stack([rng.uniform(0.02, 0.4, n), rng.uniform(0.02, 0.4, n), rng.uniform(-0.95, 0.95, n)], 1), dtype=torch.float32, device=DEV)
    xy = torch.tensor(rng.uniform(-1, 1, (n, 2)), dtype=torch.float32, device=DEV)
    col = torch.tensor(rng.uniform(0, 1, (n, 3)), dtype=torch.float32, device=DEV)
    R = oracle.RefKernels(False)
    ref = R.forward(sig, xy, col, torch.zeros(h, w, 3, device=DEV))
    img = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(sig, xy, col, img, n, h, w, 3, ksigma=float("inf"))   # 8-argument form
    torch.cuda.synchronize()
    assert float((img - ref).abs().max()) <= FWD_TOL * max(1.0, float(ref.abs().max()))
    g = torch.rand(h, w, 3, device=DEV)
    rs, rc, rk = R.backward(sig, xy, col, g, torch.zeros_like(sig), torch.zeros_like(xy), torch.zeros_like(col))
    _assert_grads(_backward(sig, xy, col, g, float("inf"), float("inf")),
                  tuple(t.cpu().double().numpy() for t in (rs, rc, rk)))


@needs_ref
@pytest.mark.parametrize("h,w,dmax", [(128, 128, 0.05), (200, 333, 0.1), (64, 512, 0.013)])
def test_inclusion_set_bit_identical_to_reference(h, w, dmax):
    """colour = 1, sigma huge => every included pixel receives ~1 per Gaussian: the per-pixel
    COUNT of contributing Gaussians must equal the reference's exactly (and the oracle's)."""
    rng = np.random.default_rng(int(dmax * 1e4) + h)
    n = 3000
    sig = torch.tensor(np.stack([np.full(n, 1e4), np.full(n, 1e4), np.zeros(n)], 1), dtype=torch.float32, device=DEV)
    # centres include exact pixel coordinates and window edges that land exactly on pixels
    px = (2.0 * np.arange(w) / (w - 1) - 1.0).astype(np.float32)
    py = (2.0 * np.arange(h) / (h - 1) - 1.0).astype(np.float32)
    xy = rng.uniform(-1.05, 1.05, (n, 2)).astype(np.float32)
    xy[:500, 0] = px[rng.integers(0, w, 500)]
    xy[:500, 1] = py[rng.integers(0, h, 500)]
    xy[500:800, 0] = px[rng.integers(0, w, 300)] + np.float32(dmax)
    xyd = torch.tensor(xy, device=DEV)
    col = torch.ones(n, 3, device=DEV)
    R = oracle.RefKernels(True)
    ref = R.forward(sig, xyd, col, torch.zeros(h, w, 3, device=DEV), dmax)
    out = _render(sig, xyd, col, h, w, dmax, float("inf"))
    _, cnt = oracle.forward(sig.cpu().numpy(), xy, col.cpu().numpy(), h, w, dmax, with_count=True)
    ours = torch.round(out[..., 0]).cpu().numpy().astype(np.int64)
    theirs = torch.round(ref[..., 0]).cpu().numpy().astype(np.int64)
    assert np.array_equal(theirs, cnt), "oracle inclusion set differs from the reference kernel"
    assert np.array_equal(ours, cnt), "inclusion set differs from the reference kernel"


# ---------------------------------------------------------------- autograd boundary
def test_gscuda_autograd_function_and_render_helper():
    g = golden("check_dmax_narrow.npz")
    h, w, dmax = int(g["h"]), int(g["w"]), float(g["dmax"])
    s = torch.tensor(g["sigmas"], device=DEV, requires_grad=True)
    c = torch.tensor(g["coords"], device=DEV, requires_grad=True)
    k = torch.tensor(g["colors"], device=DEV, requires_grad=True)
    gscuda.set_ksigma(float("inf"))
    try:
        img = gaussiansplatting_render(s, c, k, (h, w), dmax)
        assert img.shape == (h, w, 3)
        (torch.tensor(g["weight"], device=DEV) * img).sum().backward()
        buf = torch.zeros(h, w, 3, device=DEV)
        same = GSCUDA.apply(s.detach(), c.detach(), k.detach(), buf, dmax)
        assert same.data_ptr() == buf.data_ptr()          # returns the very tensor it was given
    finally:
        gscuda.set_ksigma(0.0)
    assert np.abs(img.detach().cpu().numpy() - g["img"]).max() <= FWD_TOL
    _assert_grads(tuple(t.grad.cpu().double().numpy() for t in (s, c, k)),
                  (g["g_sigmas"], g["g_coords"], g["g_colors"]))


def test_errors_like_the_reference_wrapper():
    s = torch.rand(8, 3, device=DEV)
    c = torch.rand(8, 2, device=DEV)
    k = torch.rand(8, 3, device=DEV)
    img = torch.zeros(16, 16, 3, device=DEV)
    with pytest.raises(RuntimeError, match="contiguous"):
        gscuda.gs_render(s.t().contiguous().t(), c, k, img, 8, 16, 16, 3, 0.5)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        gscuda.gs_render(s.cpu(), c, k, img, 8, 16, 16, 3, 0.5)
    with pytest.raises(RuntimeError):
        gscuda.gs_render(s, c, torch.rand(8, 4, device=DEV), torch.zeros(16, 16, 4, device=DEV), 8, 16, 16, 4, 0.5)
    with pytest.raises(RuntimeError):
        gscuda.gs_render(s.double(), c, k, img, 8, 16, 16, 3, 0.5)


def test_runs_on_the_current_stream():
    _, s, c, k, h, w = fields.make("C1", 6)
    sd, cd, kd = s.to(DEV), c.to(DEV), k.to(DEV)
    want = _render(sd, cd, kd, h, w, 0.1)
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        img = torch.zeros(h, w, 3, device=DEV)
        gscuda.gs_render(sd, cd, kd, img, sd.shape[0], h, w, 3, 0.1)
    st.synchronize()
    assert torch.allclose(img, want, atol=1e-5)


def test_clustered_field_overflows_buckets_and_falls_back():
    """All Gaussians in one corner: the fixed-capacity region buckets overflow, the device-side flag
    routes the forward to the home-bin kernel; the result must not change."""
    rng = np.random.default_rng(21)
    n, h, w = 20000, 256, 320
    sig = np.stack([rng.uniform(0.004, 0.02, n), rng.uniform(0.004, 0.02, n), np.tanh(rng.normal(0, 1, n)) * 0.99], 1)
    xy = np.stack([rng.uniform(-0.95, -0.75, n), rng.uniform(0.7, 0.95, n)], 1)
    col = rng.uniform(0, 1, (n, 3)) * 0.05
    ref = oracle.forward(sig, xy, col, h, w, 0.1)
    out = _render(sig, xy, col, h, w, 0.1).cpu().double().numpy()
    assert np.abs(out - ref).max() <= FWD_TOL * max(1.0, np.abs(ref).max())
    g = rng.uniform(-1, 1, (h, w, 3)).astype(np.float32)
    _assert_grads(_backward(sig, xy, col, g, 0.1), oracle.backward(sig, xy, col, g, 0.1))


def test_shuffled_input_takes_the_incoherent_path():
    """Random input order: the warp-cooperative bucket append sees a large tile union and every lane
    walks its own tiles; same image."""
    _, s, c, k, h, w = fields.make("C2", 3)
    perm = torch.randperm(s.shape[0], generator=torch.Generator().manual_seed(3))
    a = _render(s, c, k, h, w, 0.1)
    b = _render(s[perm].contiguous(), c[perm].contiguous(), k[perm].contiguous(), h, w, 0.1)
    assert float((a - b).abs().max()) <= 2e-5


@pytest.mark.parametrize("h,w", [(8, 32767), (32767, 6)])
def test_maximum_dimension(h, w):
    """The largest extent the C ABI accepts (32767: cull boxes are stored as 15-bit) against the oracle,
    and one more is refused."""
    rng = np.random.default_rng(h)
    n = 600
    sig = np.stack([rng.uniform(2e-4, 2e-3, n), rng.uniform(2e-4, 2e-3, n), np.tanh(rng.normal(0, 1, n)) * 0.99], 1)
    if h > w:
        sig = sig[:, [1, 0, 2]]
    sig[:, 0 if h > w else 1] *= 200.0  # a few pixels wide along the short axis too
    xy = rng.uniform(-1.0, 1.0, (n, 2))
    col = rng.uniform(0, 1, (n, 3))
    ref = oracle.forward(sig, xy, col, h, w, 0.01)
    out = _render(sig, xy, col, h, w, 0.01).cpu().double().numpy()
    assert np.abs(out - ref).max() <= FWD_TOL
    g = rng.uniform(-1, 1, (h, w, 3)).astype(np.float32)
    _assert_grads(_backward(sig, xy, col, g, 0.01), oracle.backward(sig, xy, col, g, 0.01))
    L = _lib.load()
    assert L.gsr_workspace_bytes(n, 32768, 8) == 0 and L.gsr_workspace_bytes(n, 8, 32768) == 0


@pytest.mark.parametrize("ksigma", [None, float("inf")])
def test_x8_head_field_takes_the_region_path(ksigma):
    """A fea2gs-shaped field at x8 (sigma up to 6.7 px, 5 sigma = 33 px, ~23 regions per Gaussian): the CTA's
    region rectangle exceeds the shared-memory budget of the set-up kernel (warp-ballot path) and the
    buckets must still hold everything (no fallback), with oracle parity."""
    p = fields.raw_field(48, 64, seed=4)
    h, w = 48 * 4, 64 * 4   # 24x32 LR at x8 with 2x2 Gaussians per LR pixel
    s, c, k = fields.map_field(p, h, w, 8.0)
    ref = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, 0.3)
    out = _render(s, c, k, h, w, 0.3, ksigma).cpu().double().numpy()
    assert np.abs(out - ref).max() <= FWD_TOL
    g = np.random.default_rng(0).uniform(-1, 1, (h, w, 3)).astype(np.float32)
    _assert_grads(_backward(s, c, k, g, 0.3, ksigma), oracle.backward(s.numpy(), c.numpy(), k.numpy(), g, 0.3))
