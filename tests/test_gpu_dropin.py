"""Drop-in proof: the reference's OWN python files of the path -- utils/gs_cuda_dmax/gswrapper.py (GSCUDA,
gaussiansplatting_render; its `import gscuda` resolves to this repository's top-level gscuda.py) and
utils/gaussian_splatting.py (generate_2D_gaussian_splatting_step[_buffer]) -- staged unmodified under
oracle/_ref/ref_py by the build recipe (oracle/ref_py.py), running on the B200-native kernels.  `-m gpu`."""
import numpy as np
import pytest
import torch

from conftest import golden
from gsasr_b200 import fields
from gsasr_b200 import gaussian_splatting as gsp
from oracle import oracle, ref_py

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_py.have(), reason="oracle/_ref/ref_py not staged (needs /root/reference at build time)")]
DEV = "cuda:0"


def test_reference_gswrapper_runs_on_our_gscuda_forward_and_backward():
    ref_wrap = ref_py.load("utils.gs_cuda_dmax.gswrapper")
    import gscuda as ours

    assert ref_wrap.GSWrapper is ours                     # the reference file bound OUR module
    g = golden("check_dmax_narrow.npz")
    h, w, dmax = int(g["h"]), int(g["w"]), float(g["dmax"])
    s = torch.tensor(g["sigmas"], device=DEV, requires_grad=True)
    c = torch.tensor(g["coords"], device=DEV, requires_grad=True)
    k = torch.tensor(g["colors"], device=DEV, requires_grad=True)
    img = ref_wrap.gaussiansplatting_render(s, c, k, (h, w), dmax)   # reference code: zeros + GSCUDA.apply
    assert img.shape == (h, w, 3)
    (torch.tensor(g["weight"], device=DEV) * img).sum().backward()  # reference GSCUDA.backward -> gs_render_backward
    assert np.abs(img.detach().cpu().numpy() - g["img"]).max() <= 1e-4
    for t, key in ((s, "g_sigmas"), (c, "g_coords"), (k, "g_colors")):
        ref = g[key]
        assert np.abs(t.grad.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max(), key


@pytest.mark.parametrize("scale,dmax_mode,dmax", [(4.0, "fix", 0.1), (2.5, "dynamic", 25), (1.5, "fix", 0.5)])
def test_reference_front_end_runs_unmodified_and_equals_the_mirror_and_the_oracle(scale, dmax_mode, dmax):
    ref_gsp = ref_py.load("utils.gaussian_splatting")
    lr = 40
    h, w = int(lr * scale) + 3, int(lr * scale)
    raw = fields.raw_field(2 * lr, 2 * lr, seed=7).to(DEV)
    args = dict(sr_size=torch.tensor([h, w]), scale=scale, scale_modify=torch.tensor([scale, scale]),
                default_step_size=1.2, cuda_rendering=True, mode="scale_modify", if_dmax=True, dmax_mode=dmax_mode,
                dmax=dmax)
    raw_a = raw.clone().requires_grad_(True)
    out_ref = ref_gsp.generate_2D_gaussian_splatting_step(gs_parameters=raw_a, **args)   # the reference's own code
    raw_b = raw.clone().requires_grad_(True)
    out_mir = gsp.generate_2D_gaussian_splatting_step(gs_parameters=raw_b, **args)       # this repo's mirror
    assert out_ref.shape == (3, h, w)
    assert float((out_ref - out_mir).abs().max()) <= 2e-5
    wgt = torch.rand(3, h, w, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    (out_ref * wgt).sum().backward()
    (out_mir * wgt).sum().backward()
    assert float((raw_a.grad - raw_b.grad).abs().max()) <= 1e-4 * float(raw_b.grad.abs().max())
    # and against the oracle, through the mapping the reference applied
    s, c, k = fields.map_field(raw.cpu(), h, w, scale)
    dm = (dmax + 2) / min(h, w) if dmax_mode == "dynamic" else dmax
    want = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, dm)
    assert np.abs(out_ref.detach().permute(1, 2, 0).cpu().double().numpy() - want).max() <= 1e-4
    buf = ref_gsp.generate_2D_gaussian_splatting_step_buffer(gs_parameters=raw.clone(), buffer_size=1500, **args)
    assert float((buf - out_mir.detach()).abs().max()) <= 2e-5


def test_reference_python_renderer_equals_the_mirror_on_the_gpu():
    """cuda_rendering=False: the reference's rendering_python and this repo's restatement, both on CUDA tensors."""
    ref_gsp = ref_py.load("utils.gaussian_splatting")
    raw = fields.raw_field(32, 32, seed=2).to(DEV)
    args = dict(sr_size=torch.tensor([34, 32]), scale=2.0, scale_modify=torch.tensor([2.0, 2.0]), cuda_rendering=False)
    a = ref_gsp.generate_2D_gaussian_splatting_step(gs_parameters=raw.clone(), **args)
    b = gsp.generate_2D_gaussian_splatting_step(gs_parameters=raw.clone(), **args)
    assert float((a - b).abs().max()) <= 2e-5 * max(1.0, float(a.abs().max()))
