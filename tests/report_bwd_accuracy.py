"""Error of the two backward kernels against the fp64 oracle (same field, same gradient image; lives under tests/ because it runs the oracle):
   python tests/report_bwd_accuracy.py [C1|C2]      -> max |err| / max |gradient| per output, region vs Gaussian-centric."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda
from oracle import oracle

cfg = sys.argv[1] if len(sys.argv) > 1 else "C1"
dev = torch.device("cuda:0")
_, s, c, k, h, w = fields.make(cfg, 0)
g = torch.rand(h, w, 3, generator=torch.Generator().manual_seed(1))
want = oracle.backward(s.numpy(), c.numpy(), k.numpy(), g.numpy(), 0.1)
sd, cd, kd, gd = s.to(dev), c.to(dev), k.to(dev), g.to(dev)
res = {"config": cfg, "h": h, "w": w, "gaussians": int(s.shape[0])}
for name, flags in (("region", 0), ("gaussian_centric", 0x20)):
    out = [torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)]
    gscuda.gs_render_backward(sd, cd, kd, gd, *out, s.shape[0], h, w, 3, 0.1, flags=flags)
    torch.cuda.synchronize()
    res[name] = {n: float(np.abs(o.cpu().double().numpy() - b).max() / max(np.abs(b).max(), 1e-30))
                 for o, b, n in zip(out, want, ("sigmas", "coords", "colors"))}
print(json.dumps(res))
