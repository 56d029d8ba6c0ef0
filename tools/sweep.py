"""Tuning sweep: build geometry variants locally (no GPU), time them on the GPU box.
  python tools/sweep.py build     # here: writes gpurun_out/../variants/*.so  (in-tree, travels)
  python tools/sweep.py run       # on the GPU box: times forward/backward on HL for each variant
"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "variants")
VARIANTS = {
    "base": [],
    # examples of what has been swept: GSR_CFG_FR_MIN_CTAS, GSR_CFG_FR_LW=4, GSR_CFG_FR_TAIL=0,
    # GSR_CFG_FR_UNROLL8=0, GSR_CFG_RB2_MIN_CTAS, GSR_CFG_MASK_PER_BAND=1, GSR_CFG_FALLBACK_COOP=0, GSH_CFG_EPI_PARTS=2
    "ws_sync_arrive": ["GSR_CFG_WS_SYNC_ARRIVE=1"],  # racecheck aid (profiles/r02_sanitizer.md)
    "br_c4": ["GSR_CFG_BR_MIN_CTAS=4"],   # region backward: 4 CTAs per SM (121 registers): HL 654 vs 657 us, C2d 390 vs 345 us
}
if sys.argv[1] == "build":
    from gsasr_b200 import build
    os.makedirs(VDIR, exist_ok=True)
    for name, d in VARIANTS.items():
        print(name, build.build_variant(d, os.path.join(VDIR, f"libgsraster_{name}.so")))
elif sys.argv[1] == "run":
    for name in VARIANTS:
        env = dict(os.environ, GSR_LIB_PATH=os.path.join(VDIR, f"libgsraster_{name}.so"))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep.py"), "one"], env=env, capture_output=True, text=True)
        print(name, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:], flush=True)
else:
    import torch, numpy as np
    from gsasr_b200 import fields, gscuda, _lib
    L = _lib.load(); dev = torch.device("cuda:0")
    res = {}
    for cfg in (("HL",) if os.environ.get("GSR_SWEEP_FAST") else ("HL", "C2d", "C3", "C2")):
        _, s, c, k, h, w = fields.make(cfg)
        sd, cd, kd = s.to(dev), c.to(dev), k.to(dev); n = s.shape[0]
        img = torch.zeros(h, w, 3, device=dev); ws = gscuda.workspace(n, h, w, dev)
        sp = torch.cuda.current_stream().cuda_stream
        L.gsr_prepare(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), n, h, w, 0.1, 0.0, ws.data_ptr(), ws.numel(), sp)
        ts = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); L.gsr_forward_prepared(img.data_ptr(), n, h, w, 0.0, 1, ws.data_ptr(), ws.numel(), sp); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[cfg + "_fwd_us"] = round(1e3 * float(np.median(ts[3:])), 1)
        ts = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); L.gsr_forward_prepared(img.data_ptr(), n, h, w, 0.0, 0, ws.data_ptr(), ws.numel(), sp); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[cfg + "_fwd_acc_us"] = round(1e3 * float(np.median(ts[3:])), 1)
        ts = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); L.gsr_forward_prepared(img.data_ptr(), n, h, w, 0.0, 1 | 0x20, ws.data_ptr(), ws.numel(), sp); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[cfg + "_fwd_det_us"] = round(1e3 * float(np.median(ts[3:])), 1)
        ts = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); L.gsr_forward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), img.data_ptr(), n, h, w, 3, 0.1, 0.0, 1, ws.data_ptr(), ws.numel(), sp); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[cfg + "_full_us"] = round(1e3 * float(np.median(ts[3:])), 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); L.gsr_prepare(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), n, h, w, 0.1, 0.0, ws.data_ptr(), ws.numel(), sp); b.record()
        torch.cuda.synchronize(); res[cfg + "_prep_us"] = round(1e3 * a.elapsed_time(b), 1)
        grd = torch.rand(h, w, 3, device=dev); gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); L.gsr_backward_prepared(sd.data_ptr(), grd.data_ptr(), gs.data_ptr(), gc.data_ptr(), gk.data_ptr(), n, h, w, 0, ws.data_ptr(), ws.numel(), sp); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        res[cfg + "_bwd_us"] = round(1e3 * float(np.median(ts[2:])), 1)
    print(json.dumps(res))
