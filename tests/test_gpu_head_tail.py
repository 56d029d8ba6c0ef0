"""Head-tail fusion (gsr_head_tail_forward: the five MLPs of the fea2gs head on the tcgen05 tensor cores) against the
reference's expression (utils/fea2gs.py:496-541, 611-633) evaluated with torch.  Needs a GPU: `-m gpu`."""
import pytest
import torch
import torch.nn as nn

from gsasr_b200 import _lib, head_tail

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _blocks(c, seed):
    torch.manual_seed(seed)
    return [nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, 4 * c), nn.ReLU(), nn.Linear(4 * c, k)).to(DEV)
            for k in (2, 1, 1, 3, 2)]   # mlp_block_sigma / rho / alpha / rgb / mean, gs_up_factor = 1


def _reference(query, blks, emulate):
    """fea2gs.py:611-633 on the (b, H, W, C) map.  emulate=True: the kernel's arithmetic (bf16 operands, fp32
    accumulation, first hidden layer rounded to bf16, last Linear in fp32)."""
    b, gh, gw, _ = query.shape
    outs = []
    for blk in blks:
        if emulate:
            x = query.to(torch.bfloat16).float()
            h1 = torch.relu(x @ blk[0].weight.to(torch.bfloat16).float().t() + blk[0].bias).to(torch.bfloat16).float()
            h2 = torch.relu(h1 @ blk[2].weight.to(torch.bfloat16).float().t() + blk[2].bias)
            outs.append(h2 @ blk[4].weight.t() + blk[4].bias)
        else:
            outs.append(blk(query))
    sig, rho, al, rgb, mean = [o.reshape(b, -1, o.shape[-1]) for o in outs]
    mean = mean / torch.tensor([gw, gh], device=query.device)[None, None]
    sy, sx = 1 / gh, 1 / gw
    ry, rx = torch.meshgrid(torch.linspace(sy / 2, 1 - sy / 2, gh, dtype=torch.float32, device=query.device),
                            torch.linspace(sx / 2, 1 - sx / 2, gw, dtype=torch.float32, device=query.device), indexing="ij")
    mean = mean + torch.stack((rx.reshape(-1), ry.reshape(-1)), -1)[None]
    return torch.cat([sig, rho, al, rgb, mean], -1)


def test_umma_building_block_matches_matmul():
    """One 128 x n tile per CTA: TMA (128-byte swizzle) -> tcgen05.mma (kind::f16, bf16 -> fp32 in TMEM) -> tcgen05.ld."""
    L = _lib.load()
    torch.manual_seed(0)
    for m, n in ((128, 192), (512, 192), (256, 128)):
        a = torch.randn(m, 192, device=DEV).to(torch.bfloat16)
        b = torch.randn(n, 192, device=DEV).to(torch.bfloat16)
        c = torch.full((m, n), float("nan"), device=DEV)
        _lib.check(L.gsr_test_umma_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, 192, None))
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        assert float((c - ref).abs().max()) <= 1e-3  # fp32 accumulation order only


@pytest.mark.parametrize("c,b,gh,gw", [(192, 1, 16, 8), (192, 2, 48, 40), (180, 1, 24, 24), (192, 1, 64, 75),
                                       (180, 3, 5, 7)])
def test_fused_head_tail_matches_the_reference_expression(c, b, gh, gw):
    """Against the same arithmetic in torch (tight) and against the reference's fp32 modules (bf16 operand rounding:
    the tolerance of the reference's own bf16-autocast training path); partial last tile, C = 180 padding, the
    reference-point columns exactly."""
    blks = _blocks(c, seed=c + gh)
    q = torch.randn(b, gh, gw, c, device=DEV)
    pk = head_tail.PackedHeadTail(blks, DEV)
    out = head_tail.fused_head_tail(q, pk)
    assert out.shape == (b, gh * gw, 9)
    with torch.no_grad():
        emu, ref = _reference(q, blks, True), _reference(q, blks, False)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            amp = _reference(q, blks, False).float()
    assert float((out - emu).abs().max()) <= 2e-3   # a hidden activation on a bf16 rounding boundary
    assert float((out - ref).abs().max()) <= 1e-2   # bf16 operands vs fp32
    assert float((out - ref).abs().max()) <= 1.5 * float((amp - ref).abs().max()) + 1e-3   # no worse than autocast
    # means: zero weights -> exactly the reference points (torch.linspace's formula)
    for blk in blks:
        for lin in (blk[0], blk[2], blk[4]):
            nn.init.zeros_(lin.weight), nn.init.zeros_(lin.bias)
    z = head_tail.fused_head_tail(q, head_tail.PackedHeadTail(blks, DEV))
    with torch.no_grad():
        assert torch.equal(z, _reference(q, blks, False))


def test_fused_head_tail_feeds_the_renderer():
    """raw (b, N, 9) is what generate_2D_gaussian_splatting_step takes: head tail -> fused front end -> raster."""
    from gsasr_b200 import gaussian_splatting as gsp

    blks = _blocks(192, seed=1)
    q = torch.randn(1, 32, 32, 192, device=DEV)
    raw = head_tail.fused_head_tail(q, head_tail.PackedHeadTail(blks, DEV))
    with torch.no_grad():
        ref = _reference(q, blks, False)
    a = gsp.generate_2D_gaussian_splatting_step(torch.tensor([64, 64]), raw[0], 2.0, torch.tensor([2.0, 2.0]), dmax=0.3, fused=True)
    b = gsp.generate_2D_gaussian_splatting_step(torch.tensor([64, 64]), ref[0], 2.0, torch.tensor([2.0, 2.0]), dmax=0.3, fused=True)
    assert a.shape == (3, 64, 64) and float((a - b).abs().max()) <= 5e-2


def test_reference_head_with_fused_tail_is_a_drop_in():
    """The reference's OWN Fea2GS module (utils/fea2gs.py, staged unmodified by oracle/ref_py.py) end to end:
    head(srcs, scale) against forward_fused_tail(head, srcs, scale) -- same body, fused tail -- and the images both
    parameter sets render to."""
    from oracle import ref_py
    if not ref_py.have():
        pytest.skip("reference python files not staged (oracle/_ref/ref_py)")
    from gsasr_b200 import gaussian_splatting as gsp

    fea2gs = ref_py.load("utils.fea2gs")
    torch.manual_seed(0)
    head = fea2gs.Fea2GS(inchannel=64, channel=180, num_heads=6, num_gs_seed=64, window_size=8, num_crossattn_blocks=1,
                         num_crossattn_layers=1, num_selfattn_blocks=1, num_selfattn_layers=1).to(DEV).eval()
    srcs = torch.randn(2, 64, 16, 24, device=DEV)
    scale = torch.tensor([2.0, 2.0], device=DEV)
    with torch.no_grad():
        want = head(srcs, scale)                      # (2, N, 9), fp32 modules
    got = head_tail.forward_fused_tail(head, srcs, scale)
    assert got.shape == want.shape
    assert isinstance(head.mlp_block_sigma, torch.nn.Sequential)   # the module is left as it was
    assert float((got - want).abs().max()) <= 2e-2                 # bf16 operands vs fp32 modules, |params| ~ 1
    assert float((got[..., 7:9] - want[..., 7:9]).abs().max()) <= 1e-3   # means: divided by the grid size
    h, w = 32, 48
    for b in range(2):
        a = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), got[b], 2.0, torch.tensor([2.0, 2.0]), dmax=0.3, fused=True)
        r = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), want[b], 2.0, torch.tensor([2.0, 2.0]), dmax=0.3, fused=True)
        assert float((a - r).abs().max()) <= 0.05 * max(1.0, float(r.abs().max()))


def test_pipeline_from_encoder_features_to_uint8_image():
    """inference_paper.py:117-138 after the encoder: the reference's own decoder + its own
    generate_2D_gaussian_splatting_step (running on this repo's gscuda) + numpy post-processing, against
    pipeline.render_from_features (fused tail, fused front end, fused uint8 write-out)."""
    import numpy as np
    from oracle import ref_py
    if not ref_py.have():
        pytest.skip("reference python files not staged (oracle/_ref/ref_py)")
    from gsasr_b200 import pipeline

    fea2gs = ref_py.load("utils.fea2gs")
    ref_gsp = ref_py.load("utils.gaussian_splatting")
    torch.manual_seed(1)
    head = fea2gs.Fea2GS(inchannel=64, channel=180, num_heads=6, num_gs_seed=64, window_size=8, num_crossattn_blocks=1,
                         num_crossattn_layers=1, num_selfattn_blocks=1, num_selfattn_layers=1).to(DEV).eval()
    feats = torch.randn(1, 64, 16, 16, device=DEV)
    scale, size = 2.0, torch.tensor([32, 32])
    with torch.no_grad():
        params = head(feats, torch.tensor([scale], device=DEV))[0]
        out = ref_gsp.generate_2D_gaussian_splatting_step(gs_parameters=params, sr_size=size, scale=scale, sample_coords=None,
                                                          scale_modify=torch.tensor([scale, scale]), default_step_size=1.2,
                                                          cuda_rendering=True, mode='scale_modify', if_dmax=True,
                                                          dmax_mode='fix', dmax=0.3)
    want = out.float().cpu().clamp_(0, 1).numpy()
    want = (np.transpose(want[[2, 1, 0], :, :], (1, 2, 0)) * 255.0).round().astype(np.uint8)
    got = pipeline.render_from_features(head, feats, scale, size, dmax=0.3)[0].cpu().numpy()
    assert got.shape == want.shape == (32, 32, 3)
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 6 and diff.mean() <= 1.0     # bf16 tail vs fp32 modules: a few grey levels at most
