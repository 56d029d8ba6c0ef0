"""Small end-to-end exercise of every kernel path for compute-sanitizer (memcheck / racecheck / initcheck):
forward (region path, window-binding path, bucket-overflow fallback incl. its grid barriers), backward, bands, batch,
window, uint8, deterministic mode, row stores, fused loss."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda, sharding
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
def field(n, sig):
    s = np.stack([rng.uniform(*sig, n), rng.uniform(*sig, n), np.tanh(rng.normal(0, 1, n)) * 0.99], 1)
    return tuple(torch.tensor(a, dtype=torch.float32, device=dev) for a in (s, rng.uniform(-1.1, 1.1, (n, 2)), rng.uniform(0, 1, (n, 3))))
for (h, w, n, sig, dmax) in ((97, 70, 500, (0.01, 0.1), 0.07), (64, 64, 300, (0.2, 2.0), 10.0), (33, 129, 200, (0.005, 0.05), 0.02)):
    s, c, k = field(n, sig)
    img = torch.zeros(h, w, 3, device=dev)
    gscuda.gs_render(s, c, k, img, n, h, w, 3, dmax)
    g = torch.rand(h, w, 3, device=dev)
    gs, gc, gk = torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)
    gscuda.gs_render_backward(s, c, k, g, gs, gc, gk, n, h, w, 3, dmax)
    band = torch.zeros(16, w, 3, device=dev)
    gscuda.gs_render_band(s, c, k, band, n, h, w, 3, 8, 16, dmax, flags=1)
    gscuda.gs_render_backward_band(s, c, k, g[8:24].contiguous(), gs, gc, gk, n, h, w, 3, 8, 16, dmax)
    u8 = torch.empty(h, w, 3, dtype=torch.uint8, device=dev)
    gscuda.gs_render_u8(s, c, k, u8, n, h, w, dmax, bgr=True)
    canvas = torch.zeros(3, h + 20, w + 12, device=dev)
    gscuda.gs_render_window(s, c, k, canvas, 5 * (w + 12) + 3, w + 12, 1, (h + 20) * (w + 12), [(0, 0, w - 1, h // 2), (2, h // 2 + 1, w - 3, h - 1)], n, h, w, dmax, flags=1)
# clustered field: buckets overflow -> home-bin fallback
n, h, w = 4000, 128, 128
s = torch.tensor(np.stack([np.full(n, 0.02), np.full(n, 0.02), np.zeros(n)], 1), dtype=torch.float32, device=dev)
c = torch.tensor(rng.normal(0, 0.01, (n, 2)), dtype=torch.float32, device=dev)
k = torch.rand(n, 3, device=dev)
img = torch.zeros(h, w, 3, device=dev)
gscuda.gs_render(s, c, k, img, n, h, w, 3, 0.5)
# uniform batch
sb = torch.stack([field(256, (0.01, 0.1))[0] for _ in range(3)]); cb = torch.rand(3, 256, 2, device=dev) * 2 - 1; kb = torch.rand(3, 256, 3, device=dev)
imgs = torch.zeros(3, 40, 48, 3, device=dev)
gscuda.gs_render_batch(sb, cb, kb, imgs, 0.2)
gscuda.gs_render_backward_batch(sb, cb, kb, torch.rand_like(imgs), torch.zeros_like(sb), torch.zeros_like(cb), torch.zeros_like(kb), 0.2)
# padded (ragged) batch, forward + backward, and its fused front end
from gsasr_b200 import gaussian_splatting as gsp
sizes = [(40, 48), (29, 37), (16, 50)]
imgs = torch.zeros(3, 40, 52, 3, device=dev)
gscuda.gs_render_batch_padded(sb, cb, kb, imgs, sizes, [0.2, 0.05, 0.4], flags=1)
gscuda.gs_render_backward_batch_padded(sb, cb, kb, torch.rand_like(imgs), torch.zeros_like(sb), torch.zeros_like(cb), torch.zeros_like(kb), sizes, [0.2, 0.05, 0.4])
raw = torch.randn(3, 256, 9, device=dev); raw[..., 7:9] = torch.rand(3, 256, 2, device=dev)
raw.requires_grad_(True)
out = gsp.generate_2D_gaussian_splatting_step_batch_padded([torch.tensor(z) for z in sizes], raw, [2.0, 1.5, 2.5], dmax=0.3, fused=True)
out.sum().backward()
# round 2: deterministic mode (bucket sort, short and long buckets), row stores, staged uint8, fused loss
from gsasr_b200 import losses
_, s, c, k, h, w = fields.make("C1", 1)
sd, cd, kd = s.to(dev), c.to(dev), k.to(dev)
img = torch.zeros(h, w, 3, device=dev)
gscuda.gs_render(sd, cd, kd, img, s.shape[0], h, w, 3, 0.1, flags=0x1 | 0x10 | 0x20)
big = field(2000, (0.5, 2.0))
img2 = torch.zeros(37, 53, 3, device=dev)
gscuda.gs_render(*big, img2, 2000, 37, 53, 3, flags=0x20)
u8 = torch.empty(h, w, 3, dtype=torch.uint8, device=dev)
gscuda.gs_render_u8(sd, cd, kd, u8, s.shape[0], h, w, 0.1)
srb = torch.rand(3, 3, 40, 52, device=dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
losses.l1_crop_loss_padded(srb, torch.rand(3, 3, 44, 60, device=dev), sizes, 0.5).backward()
# head tail (tcgen05 / TMEM / TMA): two tiles, the second partial; C = 180 padding
import torch.nn as nn
from gsasr_b200 import head_tail
blks = [nn.Sequential(nn.Linear(180, 180), nn.ReLU(), nn.Linear(180, 720), nn.ReLU(), nn.Linear(720, kk)).to(dev) for kk in (2, 1, 1, 3, 2)]
head_tail.fused_head_tail(torch.randn(1, 12, 15, 180, device=dev), head_tail.PackedHeadTail(blks, dev))
_, s, c, k, h, w = fields.make("C1", 0)
sharding.render_image_bands(s.to(dev), c.to(dev), k.to(dev), h, w, 0.1)
torch.cuda.synchronize()
print("sanitize_smoke done")
