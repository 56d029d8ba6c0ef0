import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda, _lib
L = _lib.load(); dev = torch.device("cuda:0")
cfg = sys.argv[1] if len(sys.argv) > 1 else "HL"
_, s, c, k, h, w = fields.make(cfg)
sd, cd, kd = s.to(dev), c.to(dev), k.to(dev); n = s.shape[0]
img = torch.zeros(h, w, 3, device=dev); ws = gscuda.workspace(n, h, w, dev)
sp = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    L.gsr_forward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), img.data_ptr(), n, h, w, 3, 0.1, 0.0, 1, ws.data_ptr(), ws.numel(), sp)
torch.cuda.synchronize()
