import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from gsasr_b200 import gscuda
from test_gpu_bands import _field
for (h, w, dmax) in ((64, 48, 0.3), (97, 70, 0.07), (50, 33, float("inf"))):
    s, c, k = _field(h, w, 400, seed=h * w)
    img = torch.zeros(h, w, 3, device="cuda:0")
    gscuda.gs_render(s, c, k, img, s.shape[0], h, w, 3, dmax)
    torch.cuda.synchronize()
    print("full", h, w, float(img.abs().max()))
    for r0, r1 in ((0, 48), (48, 50)):
        band = torch.full((r1 - r0, w, 3), 7.0, device="cuda:0")
        gscuda.gs_render_band(s, c, k, band, s.shape[0], h, w, 3, r0, r1 - r0, dmax, flags=1)
        torch.cuda.synchronize()
        print("band", r0, r1, float(band.abs().max()))
