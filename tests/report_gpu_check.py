"""Quick GPU sanity run: parity of forward/backward vs the CPU oracle and the reference kernels,
plus first timings.  Development aid (the judged artefacts are tests/ and bench.py)."""
import sys, time, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda, _lib
from oracle import oracle

dev = torch.device("cuda:0")
torch.cuda.init()
print(torch.cuda.get_device_name(0))

def ev_time(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def run(name, dmax, ks_list=(0.0,), check_oracle=True, check_ref=True, bwd=True, seed=0, compact=False):
    p, s, c, k, h, w = fields.make(name, seed, compact)
    sd, cd, kd = s.to(dev), c.to(dev), k.to(dev)
    n = s.shape[0]
    print(f"== {name} N={n} {h}x{w} dmax={dmax} compact={compact}")
    ref_o = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, dmax) if check_oracle else None
    ref_g = None
    if check_ref and oracle.have_ref():
        R = oracle.RefKernels(True)
        img = torch.zeros(h, w, 3, device=dev)
        t0 = time.time(); R.forward(sd, cd, kd, img, dmax); t_ref = time.time() - t0
        ref_g = img.cpu().double().numpy()
        print(f"   reference kernel fwd: {t_ref*1e3:.1f} ms (wall, 1 call)")
        if ref_o is not None:
            print(f"   reference vs oracle: maxabs {np.abs(ref_g-ref_o).max():.3e}")
    for ks in ks_list:
        img = torch.zeros(h, w, 3, device=dev)
        gscuda.gs_render(sd, cd, kd, img, n, h, w, 3, dmax, ksigma=ks)
        torch.cuda.synchronize()
        out = img.cpu().double().numpy()
        msg = f"   ours k={ks}:"
        if ref_o is not None: msg += f" vs oracle {np.abs(out-ref_o).max():.3e}"
        if ref_g is not None: msg += f" vs reference {np.abs(out-ref_g).max():.3e}"
        ws = gscuda.workspace(n, h, w, dev)
        t = ev_time(lambda: gscuda.gs_render(sd, cd, kd, img, n, h, w, 3, dmax, ksigma=ks, workspace_buf=ws))
        msg += f"  fwd {t*1e3:.1f} us  {h*w/1e6/(t*1e-3):.0f} MP/s"
        print(msg)
    if bwd:
        g = torch.Generator().manual_seed(seed + 100)
        gr = torch.rand(h, w, 3, generator=g)
        grd = gr.to(dev)
        gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
        gscuda.gs_render_backward(sd, cd, kd, grd, gs, gc, gk, n, h, w, 3, dmax)
        torch.cuda.synchronize()
        if check_oracle:
            os_, oc, ok = oracle.backward(s.numpy(), c.numpy(), k.numpy(), gr.numpy(), dmax)
            for nm, a, b in (("sigmas", gs, os_), ("coords", gc, oc), ("colors", gk, ok)):
                a = a.cpu().double().numpy()
                scale = np.abs(b).max()
                print(f"   bwd {nm}: maxabs {np.abs(a-b).max():.3e} (scale {scale:.3e}) rel {np.abs(a-b).max()/scale:.3e}")
        if check_ref and oracle.have_ref():
            R = oracle.RefKernels(True)
            rs, rc_, rk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
            t0 = time.time(); R.backward(sd, cd, kd, grd, rs, rc_, rk, dmax); t_ref = time.time() - t0
            print(f"   reference kernel bwd: {t_ref*1e3:.1f} ms")
            for nm, a, b in (("sigmas", gs, rs), ("coords", gc, rc_), ("colors", gk, rk)):
                d = (a - b).abs().max().item(); sc = b.abs().max().item()
                print(f"   bwd {nm} vs reference: maxabs {d:.3e} rel {d/sc:.3e}")
        ws = gscuda.workspace(n, h, w, dev)
        t = ev_time(lambda: gscuda.gs_render_backward(sd, cd, kd, grd, gs, gc, gk, n, h, w, 3, dmax, workspace_buf=ws))
        print(f"   bwd {t*1e3:.1f} us")

run("C1", 0.1, ks_list=(0.0, 4.5, 6.0, float("inf")))
run("C1", 0.05, ks_list=(0.0, float("inf")))
run("C2", 0.1, ks_list=(0.0, 6.0, float("inf")))
run("C2", 0.1, ks_list=(0.0,), compact=True, check_ref=False)
run("HL", 0.1, ks_list=(0.0, 4.0, 6.0), check_oracle=False, check_ref=False, bwd=True)
run("HL", 0.1, ks_list=(0.0,), check_oracle=False, check_ref=False, bwd=False, compact=True)
run("C3", 0.1, ks_list=(0.0,), check_oracle=False, check_ref=False, bwd=False)
run("C2d", 0.1, ks_list=(0.0,), check_oracle=False, check_ref=False, bwd=True)
