import sys, os, json, numpy as np, torch
sys.path.insert(0, "/root/repo")
from gsasr_b200 import fields, gscuda, _lib
L = _lib.load(); dev = torch.device("cuda:0")
res = {}
for cfg, dmax in (("C1", 0.1), ("C1", 0.05), ("C1", float("inf")), ("C2", 0.1)):
    _, s, c, k, h, w = fields.make(cfg)
    sd, cd, kd = s.to(dev), c.to(dev), k.to(dev); n = s.shape[0]
    img = torch.zeros(h, w, 3, device=dev); ws = gscuda.workspace(n, h, w, dev)
    sp = torch.cuda.current_stream().cuda_stream
    L.gsr_prepare(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), n, h, w, dmax, 0.0, ws.data_ptr(), ws.numel(), sp)
    ts = []
    for i in range(23):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); L.gsr_forward_prepared(img.data_ptr(), n, h, w, 0.0, 1, ws.data_ptr(), ws.numel(), sp); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    res[f"{cfg}_{dmax}_raster_us"] = round(1e3 * float(np.median(ts[3:])), 1)
    st = torch.zeros(16, dtype=torch.int32, device=dev)
print(os.environ.get("GSR_FR_WS"), json.dumps(res))
