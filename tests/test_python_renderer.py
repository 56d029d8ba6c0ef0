"""the mirror of the reference's PyTorch renderer (rendering_python, cuda_rendering=False)
against fixtures recorded from the reference's own function (tests/golden/make_golden.py).  CPU."""
import numpy as np
import pytest
import torch

from conftest import golden
from gsasr_b200 import gaussian_splatting as gsp


@pytest.mark.parametrize("name", ["python_renderer_x2.npz", "python_renderer_x3p5.npz"])
def test_python_renderer_matches_the_reference(name):
    g = golden(name)
    h, w, scale = int(g["h"]), int(g["w"]), float(g["scale"])
    raw = torch.tensor(g["raw"])
    out = gsp.generate_2D_gaussian_splatting_step(
        sr_size=torch.tensor([h, w]), gs_parameters=raw.clone(), scale=scale,
        scale_modify=torch.tensor([scale, scale]), cuda_rendering=False)
    assert out.shape == (3, h, w) and out.dtype == torch.float32
    ref = g["img"]
    # fp32 round-off of a sum of ~N terms: relative to the image's scale
    assert np.abs(out.numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    buf = gsp.generate_2D_gaussian_splatting_step_buffer(
        sr_size=torch.tensor([h, w]), gs_parameters=raw.clone(), scale=scale,
        scale_modify=torch.tensor([scale, scale]), cuda_rendering=False, buffer_size=100)
    assert torch.equal(buf, out)


def test_python_renderer_sample_coords_and_mode_checks():
    g = golden("python_renderer_x2.npz")
    h, w, scale = int(g["h"]), int(g["w"]), float(g["scale"])
    raw = torch.tensor(g["raw"])
    pts = [(0, 0), (3, 5), (h - 1, w - 1)]
    out = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), raw.clone(), scale, torch.tensor([scale, scale]),
                                                  sample_coords=pts, cuda_rendering=False)
    assert out.shape == (3, 3)
    assert np.allclose(out[:, 1].numpy(), g["img"][:, 3, 5], atol=2e-5 * np.abs(g["img"]).max())
    with pytest.raises(ValueError):
        gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), raw, scale, torch.tensor([scale, scale]),
                                                cuda_rendering=False, mode="nope")
    with pytest.raises(AssertionError):
        gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), raw, scale, torch.tensor([scale, scale + 1]),
                                                cuda_rendering=False)
    # the fused path validates like the unfused one (it used to accept any mode silently)
    with pytest.raises(ValueError):
        gsp._step_size(scale, torch.tensor([scale, scale]), 1.2, "nope")
    with pytest.raises(AssertionError):
        gsp._step_size(scale, torch.tensor([scale, scale + 1]), 1.2, "scale_modify")
