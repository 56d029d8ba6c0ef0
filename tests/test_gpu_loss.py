"""Fused crop + L1 + gradient (gsr_l1_crop_loss) against the reference training loop's per-sample expression
(gsasr_model.py:212-234: pad, crop, L1Loss(mean), / batch).  Needs a GPU: `-m gpu`."""
import pytest
import torch
import torch.nn.functional as F

from gsasr_b200 import losses

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(sr, gt, sizes, weight):
    total = 0
    for b, (h, w) in enumerate(sizes):
        total = total + weight * F.l1_loss(sr[b:b + 1, :, :h, :w], gt[b:b + 1, :, :h, :w], reduction="mean")
    return total / len(sizes)


@pytest.mark.parametrize("channels_last", [True, False])
def test_l1_crop_loss_equals_the_per_sample_loop(channels_last):
    g = torch.Generator().manual_seed(3)
    sizes = [(40, 64), (37, 51), (8, 8), (1, 64)]
    b, hmax, wmax = len(sizes), 40, 64
    sr = torch.rand(b, 3, hmax, wmax, generator=g).to(DEV)
    gt = torch.rand(b, 3, 48, 72, generator=g).to(DEV)      # ground truth padded to its own maximum
    gt[0, :, :4, :4] = sr[0, :, :4, :4]                     # exact zeros: sign(0) = 0
    if channels_last:
        sr = sr.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)   # the padded batch render's layout
    a = sr.clone().requires_grad_(True)
    want = _reference(a, gt, sizes, 0.7)
    want.backward()
    x = sr.clone().requires_grad_(True)
    got = losses.l1_crop_loss_padded(x, gt, sizes, 0.7)
    (got * 2.0).backward()                                  # the incoming gradient scales dloss/dsr
    assert abs(float(got) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))
    assert x.grad.stride() == x.stride() or x.grad.shape == x.shape
    assert float((x.grad - 2.0 * a.grad).abs().max()) <= 1e-9
    for i, (h, w) in enumerate(sizes):                      # nothing flows into the padding
        assert float(x.grad[i, :, h:, :].abs().sum()) == 0 and float(x.grad[i, :, :, w:].abs().sum()) == 0
    # fixed-order reduction: the same bits every time
    again = losses.l1_crop_loss_padded(sr, gt, sizes, 0.7)
    assert float(again) == float(got)


def test_l1_crop_loss_drives_the_padded_batch_render():
    """End to end: padded batch render -> fused loss -> backward to the raw head output, against the reference
    loop's expression on the same render."""
    from gsasr_b200 import gaussian_splatting as gsp

    g = torch.Generator().manual_seed(5)
    sizes = [(64, 64), (56, 48)]
    raw = torch.randn(2, 1024, 9, generator=g)
    raw[..., 7:9] = torch.rand(2, 1024, 2, generator=g)
    raw = raw.to(DEV)
    gt = torch.rand(2, 3, 64, 64, generator=g).to(DEV)
    outs = []
    for fn in (lambda sr: losses.l1_crop_loss_padded(sr, gt, sizes), lambda sr: _reference(sr, gt, sizes, 1.0)):
        p = raw.clone().requires_grad_(True)
        sr = gsp.generate_2D_gaussian_splatting_step_batch_padded(sizes, p, [2.0, 1.7], dmax=0.3, fused=True)
        loss = fn(sr)
        loss.backward()
        outs.append((float(loss), p.grad.clone()))
    assert abs(outs[0][0] - outs[1][0]) <= 1e-5
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 1e-6 + 1e-4 * float(outs[1][1].abs().max())
