"""Profiling driver: a few forward (+ optional backward) calls on one config, for ncu."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda
name = sys.argv[1] if len(sys.argv) > 1 else "HL"
do_bwd = len(sys.argv) > 2 and sys.argv[2] == "bwd"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
p, s, c, k, h, w = fields.make(name)
sd, cd, kd = s.to(dev), c.to(dev), k.to(dev)
n = s.shape[0]
img = torch.zeros(h, w, 3, device=dev)
ws = gscuda.workspace(n, h, w, dev)
grd = torch.rand(h, w, 3, device=dev)
gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
for _ in range(iters):
    gscuda.gs_render(sd, cd, kd, img, n, h, w, 3, 0.1, workspace_buf=ws)
    if do_bwd:
        gscuda.gs_render_backward(sd, cd, kd, grd, gs, gc, gk, n, h, w, 3, 0.1, workspace_buf=ws)
torch.cuda.synchronize()
