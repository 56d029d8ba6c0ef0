"""The library's culling code (gsr_common.cuh), run on the HOST through the test hooks, against
the oracle: exact dmax inclusion set, cull boxes, region masks, k-sigma truncation error."""
import ctypes

import numpy as np
import pytest

import emulate
from gsasr_b200 import _lib, fields
from oracle import oracle


def _window(n, ctr, dmax):
    L = _lib.load()
    lo, hi = ctypes.c_int(), ctypes.c_int()
    L.gsr_host_window_range(n, ctypes.c_float(ctr), ctypes.c_float(dmax), ctypes.byref(lo), ctypes.byref(hi))
    return lo.value, hi.value


@pytest.mark.parametrize("n", [2, 3, 10, 128, 1023, 4096])
def test_window_range_is_the_reference_inclusion_set(n):
    rng = np.random.default_rng(n)
    ctrs = np.concatenate([rng.uniform(-1.3, 1.3, 300), oracle_px(n)[rng.integers(0, n, 100)]]).astype(np.float32)
    for dmax in (0.0, 1e-4, 0.05, 0.1, 0.5, 2.5, np.float32(2.0 / (n - 1)), np.inf):
        want = oracle.ranges(np.stack([ctrs, ctrs], 1), n, n, dmax)
        for c, r in zip(ctrs, want):
            lo, hi = _window(n, float(c), float(dmax))
            if r[1] < r[0]:
                assert hi < lo
            else:
                assert (lo, hi) == (r[0], r[1]), (n, c, dmax)


def oracle_px(n):
    return emulate.pix_coords(n)


def test_window_range_special_dmax():
    assert _window(16, 0.0, float("nan")) == (0, 15)   # NaN never skips (gs.cu:41)
    lo, hi = _window(16, 0.0, -0.5)
    assert hi < lo                                       # negative dmax skips everything
    assert _window(16, 5.0, float("inf")) == (0, 15)


def _random_field(rng, s, sig_lo, sig_hi):
    sig = np.stack([rng.uniform(sig_lo, sig_hi, s), rng.uniform(sig_lo, sig_hi, s),
                    np.tanh(rng.normal(0, 1.2, s)) * 0.999999], 1).astype(np.float32)
    xy = rng.uniform(-1.15, 1.15, (s, 2)).astype(np.float32)
    col = rng.uniform(0, 1, (s, 3)).astype(np.float32)
    return sig, xy, col


@pytest.mark.parametrize("h,w,dmax", [(40, 72, 0.2), (97, 33, 0.08), (64, 64, np.inf), (50, 50, 0.5)])
def test_exact_mode_reproduces_the_inclusion_set(h, w, dmax):
    """ksigma=inf: every (Gaussian,pixel) pair of the reference with a non-flushed value is kept."""
    rng = np.random.default_rng(h * w)
    sig, xy, col = _random_field(rng, 150, 0.01, 0.4)
    ref = oracle.forward(sig, xy, col, h, w, dmax)
    img, _ = emulate.emulate_forward(sig, xy, col, h, w, dmax, float("inf"))
    assert np.abs(img - ref).max() < 1e-12
    st = emulate.host_setup(sig, xy, col, h, w, dmax, float("inf"))
    rr = oracle.ranges(xy, h, w, dmax)
    live = st[:, 0] == 1
    # cull box is inside the window, and equal to it wherever the window binds on all sides
    assert np.all(st[live, 1] >= rr[live, 0]) and np.all(st[live, 2] <= rr[live, 1])
    assert np.all(st[live, 3] >= rr[live, 2]) and np.all(st[live, 4] <= rr[live, 3])


@pytest.mark.parametrize("cfg,dmax", [("C1", 0.1), ("C1", 0.05)])
def test_default_ksigma_error_budget(cfg, dmax):
    """Default k-sigma truncation stays a decade below the 1e-4 parity tolerance."""
    _, s, c, k, h, w = fields.make(cfg)
    s, c, k = s.numpy(), c.numpy(), k.numpy()
    ref = oracle.forward(s, c, k, h, w, dmax)
    img, _ = emulate.emulate_forward(s, c, k, h, w, dmax, 0.0)
    assert np.abs(img - ref).max() < 1e-5


@pytest.mark.parametrize("h,w,dmax,ksigma", [(40, 72, 0.2, float("inf")), (97, 33, 0.08, float("inf")),
                                             (64, 64, np.inf, float("inf")), (50, 50, 0.5, float("inf"))])
def test_cell_masks_exact_mode_reproduces_the_inclusion_set(h, w, dmax, ksigma):
    """Region-bucket path, ksigma=inf: the cell masks drop nothing the reference sums (values below 2^-126 aside)."""
    rng = np.random.default_rng(h * w + 1)
    sig, xy, col = _random_field(rng, 150, 0.01, 0.4)
    ref = oracle.forward(sig, xy, col, h, w, dmax)
    img, _ = emulate.emulate_forward_cells(sig, xy, col, h, w, dmax, ksigma)
    assert np.abs(img - ref).max() < 1e-12


@pytest.mark.parametrize("cfg,dmax", [("C1", 0.1), ("C1", 0.05)])
def test_cell_masks_default_ksigma_error_budget_and_work(cfg, dmax):
    """Default k-sigma through the cell masks: truncation a decade below the 1e-4 tolerance, and the number of
    (Gaussian, pixel) pairs the masks leave is well below what whole 16x8 regions would evaluate."""
    _, s, c, k, h, w = fields.make(cfg)
    s, c, k = s.numpy(), c.numpy(), k.numpy()
    ref = oracle.forward(s, c, k, h, w, dmax)
    img, pairs = emulate.emulate_forward_cells(s, c, k, h, w, dmax, 0.0)
    assert np.abs(img - ref).max() < 1e-5
    L = _lib.load()
    out = np.zeros((4096, 3), dtype=np.int32)
    entries = sum(L.gsr_host_entries(s.ctypes.data, c.ctypes.data, k.ctypes.data, g, h, w, float(dmax), 0.0,
                                     out.ctypes.data, 4096) for g in range(0, s.shape[0], 8))
    assert pairs < 0.75 * 128 * entries * 8


def test_degenerate_gaussians_are_skipped_or_kept_consistently():
    h, w = 48, 80
    sig = np.array([[0.1, 0.1, 0.0], [0.0, 0.1, 0.0], [0.1, 0.1, 1.0], [np.nan, 0.1, 0.0],
                    [1e-9, 1e-9, 0.0], [5.0, 5.0, 0.3], [0.1, -0.2, 0.5]], np.float32)
    xy = np.array([[0, 0], [0, 0], [0, 0], [0, 0], [0.0, 0.0], [0.2, -0.3], [3.0, 0.1]], np.float32)
    col = np.ones((7, 3), np.float32)
    st = emulate.host_setup(sig, xy, col, h, w, 0.3, 0.0)
    assert st[:, 0].tolist() == [1, 0, 0, 0, 0, 1, 0]   # sigma=0, |rho|=1, NaN dropped; tiny off-grid
                                                        # and far-outside Gaussians have empty boxes
    keep = [0, 5]
    ref = oracle.forward(sig[keep], xy[keep], col[keep], h, w, 0.3)
    img, _ = emulate.emulate_forward(sig, xy, col, h, w, 0.3, 0.0)
    assert np.abs(img - ref).max() < 1e-5


def test_negative_sigma_matches_reference_formula():
    h, w = 32, 32
    sig = np.array([[0.2, -0.3, 0.6]], np.float32)
    xy = np.array([[0.1, -0.2]], np.float32)
    col = np.array([[1.0, 0.5, 0.25]], np.float32)
    ref = oracle.forward(sig, xy, col, h, w, 10.0)
    img, _ = emulate.emulate_forward(sig, xy, col, h, w, 10.0, float("inf"))
    assert np.abs(img - ref).max() < 1e-12


def test_large_class_and_bins():
    tile_w, tile_h, bin_, region, large = emulate.geometry()
    assert tile_w % region == 0 and tile_h % region == 0
    h, w = 600, 900
    sig = np.array([[0.5, 0.5, 0.0], [0.002, 0.002, 0.0]], np.float32)
    xy = np.array([[0.0, 0.0], [-1.0, 1.0]], np.float32)
    col = np.ones((2, 3), np.float32)
    st = emulate.host_setup(sig, xy, col, h, w, 10.0, 0.0)
    assert st[0, 6] == 1 and st[0, 8] > large     # wide Gaussian -> large list
    assert st[1, 6] == 0
    nbx = (w + bin_ - 1) // bin_
    assert st[1, 7] == ((h - 1) // bin_) * nbx + 0  # bottom-left corner bin


@pytest.mark.parametrize("h,w,dmax,ksigma", [(97, 40, 0.07, 0.0), (64, 64, np.inf, 0.0), (50, 50, 0.5, float("inf")),
                                             (33, 20, 0.02, float("inf"))])
def test_band_setup_is_the_whole_image_box_cut_to_the_band(h, w, dmax, ksigma):
    """gsr_forward_band's set-up: the cull box of a Gaussian in rows [row0, row0+rows) is its whole-image
    box intersected with the band (band-local rows); x is untouched; a box that misses the band is dead."""
    L = _lib.load()
    rng = np.random.default_rng(h + w)
    sig, xy, col = _random_field(rng, 200, 0.01, 0.3)
    full = emulate.host_setup(sig, xy, col, h, w, dmax, ksigma)
    for row0, rows in ((0, h), (0, 8), (8, 16), (h - 9, 9), (24, 2), (h - 2, 2)):
        out = np.zeros((200, 6), np.int32)
        L.gsr_host_setup_band(sig.ctypes.data, xy.ctypes.data, col.ctypes.data, 200, h, w, row0, rows,
                              float(dmax), float(ksigma), out.ctypes.data)
        for f, b in zip(full, out):
            y0, y1 = max(int(f[3]), row0), min(int(f[4]), row0 + rows - 1)
            if not f[0] or y0 > y1:
                assert b[0] == 0
            else:
                assert b.tolist()[:5] == [1, int(f[1]), int(f[2]), y0 - row0, y1 - row0]


def test_padded_batch_rescaling_algebra():
    """gsr_forward_batch_padded renders a sample of size (h, w) inside a larger (H, W) image by rescaling its
    records to the larger image's coordinate normalisation: x_c = (x+1)/ax - 1, (a, b, c) -> (a ax^2, b ax ay,
    c ay^2) with ax = (W-1)/(w-1), ay = (H-1)/(h-1).  Checked here in float64 on the pixel grids: the exponent
    of every pixel is unchanged, and so is the backward's chain rule after its corrections (centre gradients
    / (ax, ay), bare Sxy * ax ay)."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        h, w = int(rng.integers(5, 60)), int(rng.integers(5, 60))
        H, W = h + int(rng.integers(0, 30)), w + int(rng.integers(0, 30))
        ax, ay = (W - 1) / (w - 1), (H - 1) / (h - 1)
        x, y = rng.uniform(-1.2, 1.2, 2)
        a, b, c = -rng.uniform(1, 500), rng.uniform(-50, 50), -rng.uniform(1, 500)
        ii, jj = np.arange(w), np.arange(h)
        dx_own = (2.0 * ii / (w - 1) - 1.0) - x
        dy_own = (2.0 * jj / (h - 1) - 1.0) - y
        xc, yc = (x + 1.0) / ax - 1.0, (y + 1.0) / ay - 1.0
        dx_c = (2.0 * ii / (W - 1) - 1.0) - xc
        dy_c = (2.0 * jj / (H - 1) - 1.0) - yc
        assert np.allclose(dx_own, ax * dx_c, rtol=0, atol=1e-12) and np.allclose(dy_own, ay * dy_c, rtol=0, atol=1e-12)
        e_own = a * dx_own[None, :] ** 2 + b * dx_own[None, :] * dy_own[:, None] + c * dy_own[:, None] ** 2
        ac, bc, cc = a * ax * ax, b * ax * ay, c * ay * ay
        e_c = ac * dx_c[None, :] ** 2 + bc * dx_c[None, :] * dy_c[:, None] + cc * dy_c[:, None] ** 2
        assert np.allclose(e_own, e_c, rtol=1e-11, atol=1e-9)
        # moments with arbitrary weights u: what the backward kernel accumulates, in either unit
        u = rng.uniform(-1, 1, (h, w))
        S = lambda dx, dy: (np.sum(u * dx[None, :]), np.sum(u * dy[:, None]), np.sum(u * dx[None, :] ** 2),
                            np.sum(u * dx[None, :] * dy[:, None]), np.sum(u * dy[:, None] ** 2))
        sx, sy, sxx, sxy, syy = S(dx_own, dy_own)
        tx, ty, txx, txy, tyy = S(dx_c, dy_c)
        gx_own, gx_c = -(2 * a * sx + b * sy), -(2 * ac * tx + bc * ty)
        gy_own, gy_c = -(2 * c * sy + b * sx), -(2 * cc * ty + bc * tx)
        assert np.isclose(gx_own, gx_c / ax, rtol=1e-9, atol=1e-9) and np.isclose(gy_own, gy_c / ay, rtol=1e-9, atol=1e-9)
        assert np.isclose(b * sxy + 2 * a * sxx, bc * txy + 2 * ac * txx, rtol=1e-9, atol=1e-9)      # sigma_x term
        assert np.isclose(a * sxx + b * sxy + c * syy, ac * txx + bc * txy + cc * tyy, rtol=1e-9, atol=1e-9)  # Q
        assert np.isclose(sxy, txy * ax * ay, rtol=1e-9, atol=1e-9)


def test_band_backward_gradient_views_share_one_buffer():
    """sharding._grad_views: the three gradient arrays of the band backward are contiguous views of one flat buffer,
    (N,3) | (N,2) | (N,3), so one all-reduce covers them and nothing is packed or unpacked."""
    import torch
    from gsasr_b200 import sharding

    n = 37
    s, c, k = torch.zeros(n, 3), torch.zeros(n, 2), torch.zeros(n, 3)
    flat, gs, gc, gk = sharding._grad_views(s, c, k)
    assert flat.shape == (8 * n,) and gs.shape == (n, 3) and gc.shape == (n, 2) and gk.shape == (n, 3)
    assert gs.is_contiguous() and gc.is_contiguous() and gk.is_contiguous()
    gs += 1.0
    gc += 2.0
    gk += 3.0
    assert torch.equal(flat, torch.cat([torch.full((3 * n,), 1.0), torch.full((2 * n,), 2.0), torch.full((3 * n,), 3.0)]))
    assert gs.data_ptr() == flat.data_ptr() and gk.data_ptr() == flat.data_ptr() + 4 * 5 * n


@pytest.mark.parametrize("h,w,dmax", [(40, 72, 0.2), (97, 33, 0.08), (64, 64, np.inf)])
def test_region_backward_exact_mode_sums_the_reference_window(h, w, dmax):
    """The backward over the region buckets, ksigma=inf: the (Gaussian, cell) pairs the entries name -- clipped to the
    cull box where the dmax window binds -- are exactly the pixels the reference's backward sums (gs.cu:112-131)."""
    rng = np.random.default_rng(h * w + 7)
    sig, xy, col = _random_field(rng, 120, 0.01, 0.4)
    g = rng.uniform(0, 1, (h, w, 3)).astype(np.float32)
    want = oracle.backward(sig, xy, col, g, dmax)
    got = emulate.emulate_backward_cells(sig, xy, col, g, h, w, dmax, float("inf"))
    for a, b in zip(got, want):
        assert np.abs(a - b).max() <= 1e-9 * max(np.abs(b).max(), 1.0)


@pytest.mark.parametrize("cfg,dmax", [("C1", 0.1), ("C1", 0.05)])
def test_region_backward_default_ksigma_error_budget(cfg, dmax):
    """Default k-sigma through the cell masks: the truncation moves no gradient by more than 1e-4 of the largest one
    (tolerance of the backward parity tests: 1e-3)."""
    _, s, c, k, h, w = fields.make(cfg)
    s, c, k = s.numpy(), c.numpy(), k.numpy()
    g = np.random.default_rng(3).uniform(0, 1, (h, w, 3)).astype(np.float32)
    want = oracle.backward(s, c, k, g, dmax)
    got = emulate.emulate_backward_cells(s, c, k, g, h, w, dmax, 0.0)
    for a, b in zip(got, want):
        assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max()
