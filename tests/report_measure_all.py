"""Measurement table for every BASELINE config that fits one GPU: our forward / backward (device
time, CUDA events, median of 20 after 5 warm-ups, whole call = set-up + raster) next to the
reference's own kernels rebuilt for sm_100a (oracle/_ref, wall time of one synchronous call --
they launch on the legacy default stream), with max-abs differences where both ran.
Writes gpurun_out/measure_all.json and prints a markdown table."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda, _lib
from gsasr_b200 import gaussian_splatting as gsp
from oracle import oracle

dev = torch.device("cuda:0")
L = _lib.load()

def ev(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

rows = []
REF_LIMIT = {"C1": True, "C2": True, "C2d": True, "C3": False, "HL": False, "T480": False}
for name, dmax in (("C1", 0.1), ("C1", 0.05), ("C2", 0.1), ("C2d", 0.1), ("T480", 0.1), ("C3", 0.1), ("C3", 0.05), ("HL", 0.1)):
    p, s, c, k, h, w = fields.make(name)
    sd, cd, kd = s.to(dev), c.to(dev), k.to(dev); n = s.shape[0]
    img = torch.zeros(h, w, 3, device=dev); ws = gscuda.workspace(n, h, w, dev)
    grd = torch.rand(h, w, 3, device=dev, generator=torch.Generator(dev).manual_seed(1))
    gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
    sp = torch.cuda.current_stream().cuda_stream
    fwd = lambda: L.gsr_forward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), img.data_ptr(), n, h, w, 3, dmax, 0.0, 1, ws.data_ptr(), ws.numel(), sp)
    bwd = lambda: L.gsr_backward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), grd.data_ptr(), gs.data_ptr(), gc.data_ptr(), gk.data_ptr(), n, h, w, 3, dmax, 0.0, 0, ws.data_ptr(), ws.numel(), sp)
    t_f, t_b = ev(fwd), ev(bwd)
    raw = p.to(dev)
    fe = lambda fused: gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), raw, fields.CONFIGS[name].scale, torch.tensor([fields.CONFIGS[name].scale] * 2), dmax=dmax, fused=fused)
    t_fe, t_fef = ev(lambda: fe(False), 10, 3), ev(lambda: fe(True), 10, 3)
    row = dict(config=name, n=n, h=h, w=w, dmax=dmax, fwd_ms=t_f, bwd_ms=t_b, mp_s=h * w / 1e6 / (t_f * 1e-3),
               frontend_ms=t_fe, frontend_fused_ms=t_fef)
    if REF_LIMIT[name] and oracle.have_ref():
        R = oracle.RefKernels(True)
        out = torch.zeros(h, w, 3, device=dev); fwd(); torch.cuda.synchronize(); ours = img.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter(); R.forward(sd, cd, kd, out, dmax); row["ref_fwd_ms"] = 1e3 * (time.perf_counter() - t0)
        row["fwd_maxabs_vs_ref"] = float((ours - out).abs().max())
        rs, rc, rk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
        t0 = time.perf_counter(); R.backward(sd, cd, kd, grd, rs, rc, rk, dmax); row["ref_bwd_ms"] = 1e3 * (time.perf_counter() - t0)
        gs.zero_(); gc.zero_(); gk.zero_(); bwd(); torch.cuda.synchronize()
        row["bwd_maxrel_vs_ref"] = max(float((a - b).abs().max() / b.abs().max()) for a, b in ((gs, rs), (gc, rc), (gk, rk)))
    rows.append(row); print(row, flush=True)
json.dump(rows, open("gpurun_out/measure_all.json", "w"), indent=1)
print("| config | N | HxW | dmax | fwd ms | MP/s | bwd ms | front end ms (torch ops / fused) | ref fwd ms | ref bwd ms | fwd max-abs vs ref | bwd max-rel vs ref |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    g = lambda k, f="{:.3g}": f.format(r[k]) if k in r else "-"
    print(f"| {r['config']} | {r['n']} | {r['h']}x{r['w']} | {r['dmax']} | {r['fwd_ms']:.3f} | {r['mp_s']:.0f} | {r['bwd_ms']:.3f} | {r['frontend_ms']:.3f} / {r['frontend_fused_ms']:.3f} | {g('ref_fwd_ms')} | {g('ref_bwd_ms')} | {g('fwd_maxabs_vs_ref','{:.2e}')} | {g('bwd_maxrel_vs_ref','{:.2e}')} |")
