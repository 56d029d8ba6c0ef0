"""The oracle against the reference's own golden vectors and against itself (CPU only)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import oracle

CHECK_FILES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "tests/golden/check_*.npz")))


@pytest.mark.parametrize("name", CHECK_FILES)
def test_forward_matches_reference_check_py(name):
    """oracle == check.py torch_version (fp32 torch) within fp32 round-off."""
    g = golden(name)
    img = oracle.forward(g["sigmas"], g["coords"], g["colors"], int(g["h"]), int(g["w"]), float(g["dmax"]))
    ref = g["img"].astype(np.float64)
    assert np.abs(img - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", CHECK_FILES)
def test_backward_matches_reference_autograd(name):
    """oracle analytic backward (gs.cu:134-159) == autograd through check.py's torch_version."""
    g = golden(name)
    gs, gc, gk = oracle.backward(g["sigmas"], g["coords"], g["colors"], g["weight"], float(g["dmax"]))
    for got, key in ((gs, "g_sigmas"), (gc, "g_coords"), (gk, "g_colors")):
        ref = g[key].astype(np.float64)
        assert np.abs(got - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), key


@pytest.mark.parametrize("seed", [0, 1])
def test_c_oracle_matches_numpy_brute_force(seed):
    rng = np.random.default_rng(seed)
    s, h, w = 12, 13, 17
    sig = np.stack([rng.uniform(0.05, 0.6, s), rng.uniform(0.05, 0.6, s), rng.uniform(-0.95, 0.95, s)], 1)
    xy = rng.uniform(-1.1, 1.1, (s, 2))
    col = rng.uniform(0, 1, (s, 3))
    for dmax in (0.3, 100.0):
        a = oracle.forward(sig, xy, col, h, w, dmax)
        # the brute force evaluates coordinates in float64; mimic the fp32 rounding of the inputs only
        b = oracle.brute_force_numpy(sig.astype(np.float32), xy.astype(np.float32), col.astype(np.float32), h, w, dmax)
        assert np.abs(a - b).max() < 5e-6


def test_fp32_mode_close_to_exact_mode():
    rng = np.random.default_rng(3)
    s, h, w = 200, 40, 56
    sig = np.stack([rng.uniform(0.02, 0.2, s), rng.uniform(0.02, 0.2, s), rng.uniform(-0.9, 0.9, s)], 1)
    xy = rng.uniform(-1, 1, (s, 2))
    col = rng.uniform(0, 1, (s, 3))
    a = oracle.forward(sig, xy, col, h, w, 0.25, mode=0)
    b = oracle.forward(sig, xy, col, h, w, 0.25, mode=1)
    assert np.abs(a - b).max() < 1e-4


def test_inclusion_ranges_are_contiguous_and_inclusive():
    """|d| == dmax is INSIDE (the reference skips on > and <, gs.cu:41,48)."""
    # centre exactly on pixel 3 of a 9-pixel axis (coordinate -0.25), dmax = 0.25: pixels 2..4
    xy = np.array([[-0.25, -0.25]], np.float32)
    r = oracle.ranges(xy, 9, 9, 0.25)
    assert r.tolist() == [[2, 4, 2, 4]]
    assert oracle.ranges(xy, 9, 9, 0.0).tolist() == [[3, 3, 3, 3]]
    r = oracle.ranges(xy, 9, 9, -1.0)
    assert r[0, 1] < r[0, 0] and r[0, 3] < r[0, 2]


def test_accumulates_into_initial_image():
    g = golden("check_dmax_seed0.npz")
    h, w = int(g["h"]), int(g["w"])
    base = np.full((h, w, 3), 0.5)
    a = oracle.forward(g["sigmas"], g["coords"], g["colors"], h, w, float(g["dmax"]), init=base)
    b = oracle.forward(g["sigmas"], g["coords"], g["colors"], h, w, float(g["dmax"]))
    assert np.allclose(a, b + 0.5)


def test_frontend_fixture_matches_field_mapping():
    """gsasr_b200.fields.map_field restates the reference front end bit for bit (CPU torch)."""
    import torch

    from gsasr_b200 import fields

    for name in ("frontend_x4_fix.npz", "frontend_x2p5_dynamic.npz"):
        g = golden(name)
        sig, xy, col = fields.map_field(torch.from_numpy(g["raw"]), int(g["h"]), int(g["w"]), float(g["scale"]))
        assert np.array_equal(sig.numpy(), g["sigmas"])
        assert np.array_equal(xy.numpy(), g["coords"])
        assert np.array_equal(col.numpy(), g["colors"])


def test_crop_render_equals_the_same_rectangle_of_the_whole_render():
    """oracle.forward_crop is the checker of the full-size GPU tests: pin it to oracle.forward."""
    rng = np.random.default_rng(9)
    s, h, w = 500, 70, 90
    sig = np.stack([rng.uniform(0.01, 0.2, s), rng.uniform(0.01, 0.2, s), rng.uniform(-0.95, 0.95, s)], 1)
    xy = rng.uniform(-1.1, 1.1, (s, 2))
    col = rng.uniform(0, 1, (s, 3))
    for dmax in (0.12, float("inf")):
        whole = oracle.forward(sig, xy, col, h, w, dmax)
        for y0, x0, ch, cw in ((0, 0, 16, 24), (54, 66, 16, 24), (20, 31, 33, 7), (0, 0, h, w)):
            crop = oracle.forward_crop(sig, xy, col, h, w, dmax, y0, x0, ch, cw)
            assert np.array_equal(crop, whole[y0:y0 + ch, x0:x0 + cw])
