import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda
dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "HL"
_, s, c, k, h, w = fields.make(cfg)
sd, cd, kd = s.to(dev), c.to(dev), k.to(dev); n = s.shape[0]
ws = gscuda.workspace(n, h, w, dev)
grd = torch.rand(h, w, 3, device=dev)
gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
for _ in range(2):
    gscuda.gs_render_backward(sd, cd, kd, grd, gs, gc, gk, n, h, w, 3, 0.1, workspace_buf=ws)
torch.cuda.synchronize()
