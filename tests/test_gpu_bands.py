"""Row bands of one image (gsr_forward_band / gsr_backward_band, the multi-GPU split of a single large
image): the bands of a partition stacked together must equal the whole-image render, with the FULL
image's pixel coordinates and dmax inclusion set.  Needs a GPU: `-m gpu`."""
import numpy as np
import pytest
import torch

from gsasr_b200 import fields, gscuda, sharding
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FWD_TOL = 1e-4   # vs the fp64 oracle (north_star tolerance)
SELF_TOL = 2e-6  # vs our own whole-image render: same terms, different summation order


def _field(h, w, n, seed, sig=(0.01, 0.12)):
    rng = np.random.default_rng(seed)
    s = np.stack([rng.uniform(*sig, n), rng.uniform(*sig, n), np.tanh(rng.normal(0, 1, n)) * 0.99], 1)
    c = rng.uniform(-1.1, 1.1, (n, 2))
    k = rng.uniform(0, 1, (n, 3))
    return tuple(torch.tensor(a, dtype=torch.float32, device=DEV) for a in (s, c, k))


def _full(s, c, k, h, w, dmax, ksigma=None):
    img = torch.zeros(h, w, 3, device=DEV)
    gscuda.gs_render(s, c, k, img, s.shape[0], h, w, 3, dmax, ksigma=ksigma)
    return img


@pytest.mark.parametrize("h,w,cuts,dmax", [
    (64, 48, [0, 32, 64], 0.3),          # two aligned bands
    (97, 70, [0, 8, 40, 95, 97], 0.07),  # ragged bands, unaligned cut, 2-row tail; dmax windows bind
    (50, 33, [0, 48, 50], float("inf")),
    (40, 40, [0, 16, 24, 40], 0.02),     # windows narrower than a band
])
@pytest.mark.parametrize("ksigma", [None, float("inf")])
def test_bands_stack_to_the_whole_image(h, w, cuts, dmax, ksigma):
    s, c, k = _field(h, w, 400, seed=h * w)
    want = _full(s, c, k, h, w, dmax, ksigma)
    ref = oracle.forward(s.cpu().numpy(), c.cpu().numpy(), k.cpu().numpy(), h, w, dmax)
    parts = []
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        band = torch.full((r1 - r0, w, 3), 7.0, device=DEV)  # overwrite mode must not read it
        gscuda.gs_render_band(s, c, k, band, s.shape[0], h, w, 3, r0, r1 - r0, dmax, ksigma=ksigma, flags=1)
        parts.append(band)
    got = torch.cat(parts, 0)
    torch.cuda.synchronize()
    assert float((got - want).abs().max()) <= SELF_TOL
    if ksigma is not None:  # exact mode: the full image's inclusion set, bit for bit
        assert np.abs(got.cpu().double().numpy() - ref).max() <= 2e-5
    assert np.abs(got.cpu().double().numpy() - ref).max() <= FWD_TOL


def test_band_inclusion_counts_are_exact():
    """Colours 1, huge sigma: every in-window pixel gets ~1 per Gaussian, so the rounded band images are
    the per-pixel inclusion COUNTS of the full image (gs.cu:40-50), compared exactly with the oracle."""
    h, w, n, dmax = 61, 45, 300, 0.11
    rng = np.random.default_rng(5)
    s = torch.tensor(np.stack([np.full(n, 1e3), np.full(n, 1e3), np.zeros(n)], 1), dtype=torch.float32, device=DEV)
    c = torch.tensor(rng.uniform(-1.05, 1.05, (n, 2)), dtype=torch.float32, device=DEV)
    k = torch.ones(n, 3, device=DEV)
    rr = oracle.ranges(c.cpu().numpy(), h, w, dmax)
    cnt = np.zeros((h, w))
    for x0, x1, y0, y1 in rr:
        if x1 >= x0 and y1 >= y0:
            cnt[y0:y1 + 1, x0:x1 + 1] += 1
    for r0, rows in ((0, 24), (24, 8), (32, 29)):
        band = torch.zeros(rows, w, 3, device=DEV)
        gscuda.gs_render_band(s, c, k, band, n, h, w, 3, r0, rows, dmax, ksigma=float("inf"))
        got = np.rint(band[..., 0].cpu().numpy())
        assert np.array_equal(got, cnt[r0:r0 + rows])


def test_band_gradients_sum_to_the_whole_image_gradients():
    h, w, dmax = 72, 56, 0.15
    s, c, k = _field(h, w, 300, seed=3)
    g = torch.rand(h, w, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    want = [torch.zeros_like(t) for t in (s, c, k)]
    gscuda.gs_render_backward(s, c, k, g, *want, s.shape[0], h, w, 3, dmax)
    acc = [torch.zeros_like(t) for t in (s, c, k)]
    for r0, rows in ((0, 16), (16, 40), (56, 16)):
        gscuda.gs_render_backward_band(s, c, k, g[r0:r0 + rows].contiguous(), *acc, s.shape[0], h, w, 3, r0, rows, dmax)
    torch.cuda.synchronize()
    ref = oracle.backward(s.cpu().numpy(), c.cpu().numpy(), k.cpu().numpy(), g.cpu().numpy(), dmax)
    for a, b, r in zip(acc, want, ref):
        scale = float(b.abs().max())
        # wide boxes are swept strip by strip inside the k-sigma ellipse, and the strips of a band start at the
        # band's first row: the two sweeps drop slightly different sub-exp(-12.5) terms
        assert float((a - b).abs().max()) <= 1e-4 * scale
        assert np.abs(a.cpu().double().numpy() - r).max() <= 1e-3 * np.abs(r).max()


def test_band_argument_errors():
    s, c, k = _field(32, 32, 8, seed=0)
    band = torch.zeros(8, 32, 3, device=DEV)
    with pytest.raises(RuntimeError):
        gscuda.gs_render_band(s, c, k, band, 8, 32, 32, 3, 28, 8, 0.1)   # band leaves the image
    with pytest.raises(RuntimeError):
        gscuda.gs_render_band(s, c, k, band[:1].contiguous(), 8, 32, 32, 3, 0, 1, 0.1)  # rows < 2


def test_single_process_band_api_equals_whole_image():
    """sharding.render_image_bands / backward_image_bands without a process group: one band."""
    _, s, c, k, h, w = fields.make("C1", 0)
    s, c, k = s.to(DEV), c.to(DEV), k.to(DEV)
    got = sharding.render_image_bands(s, c, k, h, w, 0.1)
    want = _full(s, c, k, h, w, 0.1)
    assert float((got - want).abs().max()) <= SELF_TOL
    g = torch.rand(h, w, 3, device=DEV, generator=torch.Generator(DEV).manual_seed(2))
    gs, gc, gk = sharding.backward_image_bands(s, c, k, g, h, w, 0.1)
    ws = [torch.zeros_like(t) for t in (s, c, k)]
    gscuda.gs_render_backward(s, c, k, g, *ws, s.shape[0], h, w, 3, 0.1)
    for a, b in zip((gs, gc, gk), ws):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_direct_stitching_equals_tile_buffers_and_paste():
    """split_and_joint_image(direct=True): tiles rendered straight into the canvas (gsr_forward_window with
    the ownership regions) against the reference's route (tile buffers + stitch_tiles)."""
    from gsasr_b200.split_and_joint_image import plan_tiles, split_and_joint_image

    for scale, split, overlap, crop, hw in ((4.0, 40, 8, 4, (72, 100)), (2.5, 32, 6, 2, (60, 60))):
        lq = torch.rand(1, 3, *hw, device=DEV)
        plan = plan_tiles(hw[0], hw[1], scale, split, overlap)
        raws = [fields.raw_field(2 * split, 2 * split, seed=i).to(DEV) for i in range(plan.n)]
        calls = [0]

        def fea2gs(feat, sv):
            calls[0] += 1
            return raws[(calls[0] - 1) % plan.n].unsqueeze(0)

        sm = torch.tensor([scale, scale])
        want = split_and_joint_image(lq, scale, split, overlap, lambda t: t, fea2gs, sm, crop_size=crop,
                                     if_dmax=True, dmax=0.1)
        calls[0] = 0
        got = split_and_joint_image(lq, scale, split, overlap, lambda t: t, fea2gs, sm, crop_size=crop,
                                    if_dmax=True, dmax=0.1, direct=True)
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 2e-6


def test_window_render_touches_only_its_clip_rectangles():
    s, c, k = _field(48, 40, 300, seed=9)
    full = _full(s, c, k, 48, 40, 0.2).permute(2, 0, 1).contiguous()          # (3,48,40)
    canvas = torch.full((3, 80, 64), -5.0, device=DEV)
    clips = [(4, 2, 30, 20), (0, 30, 39, 47)]
    gscuda.gs_render_window(s, c, k, canvas, 10 * 64 + 7, 64, 1, 80 * 64, clips, s.shape[0], 48, 40, 0.2, flags=1)
    torch.cuda.synchronize()
    want = torch.full((3, 80, 64), -5.0, device=DEV)
    for x0, y0, x1, y1 in clips:
        want[:, 10 + y0:10 + y1 + 1, 7 + x0:7 + x1 + 1] = full[:, y0:y1 + 1, x0:x1 + 1]
    assert float((canvas - want).abs().max()) <= 2e-6
    with pytest.raises(RuntimeError):
        gscuda.gs_render_window(s, c, k, canvas, 70 * 64, 64, 1, 80 * 64, [], s.shape[0], 48, 40, 0.2)  # leaves dst


@pytest.mark.parametrize("h,w", [(64, 48), (72, 100), (40, 70), (33, 64)])
def test_row_stores_write_the_same_pixels(h, w):
    """GSR_FLAG_ROW_STORES (finished regions leave as 128-bit stores of whole region rows, staged through shared
    memory: what render_image_bands_peer uses for an image in another GPU's memory) changes how pixels are stored,
    not their values: equal to the plain overwrite up to the summation order of two runs, incl. partial edge
    regions and widths that are not a multiple of 4 (where the flag must fall back)."""
    s, c, k = _field(h, w, 600, seed=h + w)
    a = torch.full((h, w, 3), 7.0, device=DEV)
    b = torch.full((h, w, 3), -3.0, device=DEV)
    gscuda.gs_render(s, c, k, a, s.shape[0], h, w, 3, 0.3, flags=0x1)
    gscuda.gs_render(s, c, k, b, s.shape[0], h, w, 3, 0.3, flags=0x1 | 0x10)
    torch.cuda.synchronize()
    assert float((a - b).abs().max()) <= 2 * SELF_TOL
    # the uint8 write-out takes the staged path whenever w % 16 == 0 and the per-pixel path otherwise
    u8 = torch.full((h, w, 3), 77, dtype=torch.uint8, device=DEV)
    gscuda.gs_render_u8(s, c, k, u8, s.shape[0], h, w, 0.3)
    want = (a.clamp(0, 1) * 255.0).round().to(torch.int16)
    diff = (u8.to(torch.int16) - want).abs()
    assert int(diff.max()) <= 1 and float((diff > 0).float().mean()) < 1e-3
