"""Generates the committed golden fixtures from the REFERENCE's own code, run in this container
(CPU only).  Run once from the repository root:  python tests/golden/make_golden.py

  check_dmax_seed*.npz / check_plain_seed*.npz
      The reference's only parity check is utils/gs_cuda*/check.py: `torch_version`, a brute-force
      per-pixel evaluation of the closed form with the `<= dmax` mask, plus autograd gradients of
      loss = sum(weight * img).  We import that function unmodified (its module also imports the
      CUDA wrapper, which is stubbed out) and run it on seeded inputs drawn exactly like
      check.py:35-46 (dmax variant: s=4, 10x10, dmax=0.5) and gs_cuda/check.py:31-38 (s=40, 49x49).

  frontend_*.npz
      The tensors the reference's front end hands to its CUDA kernel.  We import
      utils/gaussian_splatting.py unmodified, replace the CUDA autograd function by a recorder,
      and call generate_2D_gaussian_splatting_step on a seeded raw (N,9) tensor: the recorded
      (sigmas, coords, colors, dmax) pin activations + unit/coordinate mapping (:174-180,121-123).

  python_renderer_*.npz
      The image of the reference's PyTorch renderer (rendering_python, cuda_rendering=False) on a seeded field.

Nothing here runs at test time; tests only read the .npz files.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_check(path):
    stub = types.ModuleType("gswrapper")
    stub.gaussiansplatting_render = None
    sys.modules["gswrapper"] = stub
    spec = importlib.util.spec_from_file_location("ref_check_" + str(abs(hash(path))), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_check(mod, s, size, dmax, seed, sigma_scale):
    g = torch.Generator().manual_seed(seed)
    sigmas = 0.999 * torch.rand(s, 3, generator=g)
    sigmas[:, :2] = sigma_scale * sigmas[:, :2]
    coords = 2 * torch.rand(s, 2, generator=g) - 1.0
    colors = torch.rand(s, 3, generator=g)
    weight = torch.rand(size[0], size[1], 3, generator=g)
    sigmas.requires_grad_(True)
    coords.requires_grad_(True)
    colors.requires_grad_(True)
    if dmax is None:
        img = mod.torch_version(sigmas, coords, colors, size)
    else:
        img = mod.torch_version(sigmas, coords, colors, size, dmax)
    loss = torch.sum(weight * img)
    loss.backward()
    return dict(sigmas=sigmas.detach().numpy(), coords=coords.detach().numpy(),
                colors=colors.detach().numpy(), weight=weight.numpy(), img=img.detach().numpy(),
                g_sigmas=sigmas.grad.numpy(), g_coords=coords.grad.numpy(),
                g_colors=colors.grad.numpy(), h=size[0], w=size[1],
                dmax=np.float32(np.inf if dmax is None else dmax))


def frontend_fixture(n_side, scale, seed, dmax, dmax_mode):
    sys.path.insert(0, REF)
    rec = {}

    class Recorder:
        @staticmethod
        def apply(sigmas, coords, colors, rendered_img, dmax=None):
            rec.update(sigmas=sigmas.clone().numpy(), coords=coords.clone().numpy(),
                       colors=colors.clone().numpy(),
                       dmax=np.float32(np.inf if dmax is None else float(dmax)))
            return rendered_img

    for name in ("utils.gs_cuda_dmax.gswrapper", "utils.gs_cuda.gswrapper"):
        m = types.ModuleType(name)
        m.GSCUDA = Recorder
        sys.modules[name] = m
    import utils.gaussian_splatting as gsp  # the reference's own front end

    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(n_side * n_side, 9, generator=g)
    raw[:, 7:9] = torch.rand(n_side * n_side, 2, generator=g)
    lr = n_side // 2
    h, w = int(np.floor(lr * scale)) + 3, int(np.floor(lr * scale))
    out = gsp.generate_2D_gaussian_splatting_step(
        sr_size=torch.tensor([h, w]), gs_parameters=raw.clone(), scale=scale,
        scale_modify=torch.tensor([scale, scale]), default_step_size=1.2, cuda_rendering=True,
        mode='scale_modify', if_dmax=True, dmax_mode=dmax_mode, dmax=dmax)
    assert tuple(out.shape) == (3, h, w)
    rec.update(raw=raw.numpy(), h=h, w=w, scale=np.float32(scale), dmax_in=np.float32(dmax),
               dmax_mode=dmax_mode)
    return rec


def python_renderer_fixture(n_side, scale, seed):
    """The reference's PyTorch renderer (utils/gaussian_splatting.py:11-84, cuda_rendering=False) run unmodified on
    the CPU on a seeded raw (N,9) tensor: pins the mirror gsasr_b200.gaussian_splatting.rendering_python."""
    sys.path.insert(0, REF)
    for name in ("utils.gs_cuda_dmax.gswrapper", "utils.gs_cuda.gswrapper"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import utils.gaussian_splatting as gsp

    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(n_side * n_side, 9, generator=g)
    jj, ii = torch.meshgrid(torch.arange(n_side, dtype=torch.float32), torch.arange(n_side, dtype=torch.float32),
                            indexing="xy")
    raw[:, 7] = (jj.reshape(-1) + 0.5) / n_side + raw[:, 7] * (0.5 / n_side)
    raw[:, 8] = (ii.reshape(-1) + 0.5) / n_side + raw[:, 8] * (0.5 / n_side)
    lr = n_side // 2
    h, w = int(np.floor(lr * scale)) + 2, int(np.floor(lr * scale))
    out = gsp.generate_2D_gaussian_splatting_step(
        sr_size=torch.tensor([h, w]), gs_parameters=raw.clone(), scale=scale,
        scale_modify=torch.tensor([scale, scale]), default_step_size=1.2, cuda_rendering=False,
        mode='scale_modify')
    assert tuple(out.shape) == (3, h, w)
    return dict(raw=raw.numpy(), img=out.numpy(), h=h, w=w, scale=np.float32(scale))


def stitch_fixture(h_lq, w_lq, scale, split, overlap, crop, seed):
    """The reference's split_and_joint_image (utils/split_and_joint_image.py:98-232) run unmodified
    on CPU with stub networks and a stub renderer that returns seeded random tiles: pins the tile
    geometry and every paste rule (integer and non-integer scale branches)."""
    sys.path.insert(0, REF)
    import utils.split_and_joint_image as sj

    calls = []

    def fake_render(sr_size, gs_parameters, **kw):
        g = torch.Generator().manual_seed(1000 * seed + len(calls))
        calls.append(int(sr_size[0]))
        return torch.rand(3, int(sr_size[0]), int(sr_size[1]), generator=g)

    sj.generate_2D_gaussian_splatting_step = fake_render
    lq = torch.rand(1, 3, h_lq, w_lq, generator=torch.Generator().manual_seed(seed))
    out = sj.split_and_joint_image(lq, scale, split, overlap, lambda t: t, lambda f, s: torch.zeros(1, 4, 9),
                                   torch.tensor([scale, scale]), crop_size=crop)
    return dict(out=out.numpy(), h_lq=h_lq, w_lq=w_lq, scale=np.float32(scale), split=split, overlap=overlap,
                crop=crop, seed=seed, n_tiles=len(calls))


def main():
    cd = load_check(os.path.join(REF, "utils/gs_cuda_dmax/check.py"))
    cp = load_check(os.path.join(REF, "utils/gs_cuda/check.py"))
    for seed in range(3):
        np.savez(os.path.join(OUT, f"check_dmax_seed{seed}.npz"), **run_check(cd, 4, (10, 10), 0.5, seed, 5.0))
        np.savez(os.path.join(OUT, f"check_plain_seed{seed}.npz"), **run_check(cp, 40, (49, 49), None, seed, 1.0))
    # narrower Gaussians and a binding window on a non-square image
    np.savez(os.path.join(OUT, "check_dmax_narrow.npz"), **run_check(cd, 24, (37, 45), 0.2, 7, 0.15))
    np.savez(os.path.join(OUT, "frontend_x4_fix.npz"), **frontend_fixture(16, 4.0, 0, 0.1, 'fix'))
    np.savez(os.path.join(OUT, "frontend_x2p5_dynamic.npz"), **frontend_fixture(12, 2.5, 1, 25, 'dynamic'))
    np.savez(os.path.join(OUT, "python_renderer_x2.npz"), **python_renderer_fixture(24, 2.0, 3))
    np.savez(os.path.join(OUT, "python_renderer_x3p5.npz"), **python_renderer_fixture(16, 3.5, 4))
    np.savez(os.path.join(OUT, "stitch_x4.npz"), **stitch_fixture(40, 52, 4.0, 16, 4, 2, 0))
    np.savez(os.path.join(OUT, "stitch_x2p5.npz"), **stitch_fixture(37, 45, 2.5, 14, 3, 2, 1))
    np.savez(os.path.join(OUT, "stitch_x3p3.npz"), **stitch_fixture(50, 31, 3.3, 12, 2, 3, 2))
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
