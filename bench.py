#!/usr/bin/env python
"""bench.py -- the headline benchmark of the render path (BASELINE.json):

    metric   HR megapixels/sec rasterized (fwd) at x4, 2M Gaussians; % HBM roofline
    workload "HL": 512x1024 LR -> x4 = 2048x4096 HR (8.39 MP), 2,097,152 synthetic Gaussians
             (seeded raw head tensor, model-like sigma, the reference's own activation/mapping),
             dmax 0.1, library-default k-sigma truncation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload HL]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE for N > 1).  The path has no
exchange step, so at N GPUs every rank rasterises its own image of the same shape (weak scaling,
no collective in the timed region); `value` is the whole-job MP/s: N * MP / max-over-ranks time.

A step is one forward pass of the hot path (set-up + raster kernels) with the Gaussian tensors
and the image resident in HBM.  `e2e` is the same metric through the plugin call a GSASR user
makes (gscuda.gs_render) from pinned HOST buffers, host<->device copies inside the timed region.
`roofline` is for the dominant kernel (gsr_forward_region_kernel), timed with CUDA events on the launch
stream via the split-phase C ABI.  `cpu_baseline` / `--impl reference` time the CPU restatement
of the reference algorithm (oracle/, OpenMP over all host cores) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "HR megapixels/sec rasterized (fwd) at x4, 2M Gaussians"
UNIT = "MP/s"
DMAX = 0.1
KERNELS_PER_STEP = 7  # table, region_build, forward_region + the guarded fallback (bin, scan, scatter, forward_bins)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled every 20 ms (NVML) while the timed region runs."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.reasons, self.mx, self._stop, self._thr = index, [], set(), None, False, None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(
                            pynvml, "nvmlDeviceGetCurrentClocksEventReasons") else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for bit, name in self.BAD.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.02)

            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception as exc:  # no NVML: report that, do not fake numbers
            self.reasons.add(f"nvml unavailable: {exc}")

    def stop(self):
        self._stop = True
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(sm)}


class CpuPort:
    """The CPU restatement of the reference algorithm (oracle/gs_oracle.c, fp32 mode, dmax window
    semantics, OpenMP over all host cores) on an evenly strided sample of the workload's Gaussians."""

    def __init__(self, cfg_name: str, threads: int | None = None):
        import numpy as np

        from gsasr_b200 import fields
        from oracle import oracle

        oracle.build()
        if threads:
            oracle.set_num_threads(threads)
        self.oracle, self.np, self.name = oracle, np, cfg_name
        self.cores = oracle.num_threads()
        _, s, c, k, self.h, self.w = fields.make(cfg_name, 0)
        self.s, self.c, self.k = s.numpy(), c.numpy(), k.numpy()
        self.n = self.s.shape[0]

    def run(self, m: int) -> float:
        idx = self.np.linspace(0, self.n - 1, m).astype(self.np.int64)  # every row band gets work
        t0 = time.perf_counter()
        self.oracle.forward(self.s[idx], self.c[idx], self.k[idx], self.h, self.w, DMAX, mode=1)
        return time.perf_counter() - t0

    def calibrate(self, seconds: float) -> int:
        """Sample size whose render takes about `seconds`."""
        m = min(self.n, 32 * self.cores)
        while True:
            t = self.run(m)
            if t >= 0.3 * seconds or m >= self.n:
                return max(1, min(self.n, int(m * seconds / max(t, 1e-4))))
            m = min(self.n, max(m + 1, int(m * min(8.0, 0.8 * seconds / max(t, 1e-4)))))

    def mps(self, m: int, t: float) -> float:
        return (self.h * self.w / 1e6) / (t * self.n / m)

    def describe(self, m: int, t: float) -> str:
        return (f"{m} of {self.n} Gaussians of {self.name} (evenly strided; {self.h}x{self.w}, dmax {DMAX}) rendered in "
                f"{t:.2f} s by oracle/gs_oracle.c (fp32 mode, OpenMP {self.cores} threads); MP/s = MP / (t * {self.n}/{m})")


def cpu_sample(cfg_name: str, seconds: float = 12.0):
    port = CpuPort(cfg_name)
    m = port.calibrate(seconds)
    t = port.run(m)
    return port.mps(m, t), port.cores, port.describe(m, t), t


def run_reference(args, rank):
    """--impl reference: every step renders the same bounded sample; the whole run is sized to ~90 s."""
    if rank != 0:
        return
    port = CpuPort(args.workload)
    total = args.warmup + args.steps
    m = port.calibrate(max(0.05, 90.0 / max(total, 1)))
    ts = []
    for i in range(total):
        t = port.run(m)
        if i >= args.warmup:
            ts.append(t)
    t = sum(ts) / len(ts)
    value = port.mps(m, t)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args.workload), split=args.split),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": port.cores, "kind": "port",
                         "sample": port.describe(m, t)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's own implementation of this path is CUDA-only (gs_cuda_dmax); its CPU arm is the "
                "oracle port of that algorithm on the host cores; a step renders the bounded sample described in "
                "cpu_baseline.sample and value extrapolates it to the whole image",
    }
    print(json.dumps(out), flush=True)


def workload_config(name):
    from gsasr_b200 import fields
    cfg = fields.CONFIGS[name]
    h, w = cfg.hr
    return {"workload": name, "lr": [cfg.lr_h, cfg.lr_w], "scale": cfg.scale, "hr": [h, w],
            "gaussians": cfg.n, "dmax": DMAX, "sigma": "model-like: 0.99999*sigmoid(N(0,1))+1e-6",
            "ksigma": "library default (5)",
            "l2": "per-step working set (params 67 MB + image 101 MB + workspace 564 MB) exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="HL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compact", action="store_true",
                    help="compact-sigma field (sigma_px ~ 0.6 instead of the model-like 1.67): few pixels per "
                         "Gaussian, the regime where the raster approaches its HBM roofline (SURVEY 8d)")
    ap.add_argument("--split", default="images", choices=["images", "bands", "bands-peer"],
                    help="N > 1: 'images' = every rank its own image (weak scaling, the default); "
                         "'bands' = one image split into row bands + all-gather (strong scaling); 'bands-peer' = the "
                         "bands stored straight into rank 0's image over NVLink by the raster kernel")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from gsasr_b200 import _lib, fields, gscuda, sharding
    from gsasr_b200 import build as gbuild

    gbuild.build()
    L = _lib.load()  # raises if the CUDA library is missing: there is no fallback
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the communicator is created: keep stdout for the
        # one JSON line (the banner goes to stderr instead)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    cfg = fields.CONFIGS[args.workload]
    # every rank its own image (weak scaling); the same image on every rank when it is split into bands
    _, s, c, k, h, w = fields.make(cfg, seed=0 if args.split != "images" else rank, compact=args.compact)
    n = s.shape[0]
    mp_img = h * w / 1e6
    s_h, c_h, k_h = (t.pin_memory() for t in (s, c, k))
    sd, cd, kd = s_h.to(dev), c_h.to(dev), k_h.to(dev)
    img = torch.zeros(h, w, 3, device=dev)
    ws = gscuda.workspace(n, h, w, dev)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def step_bands():
        # --split bands (N > 1): ONE image, every rank renders its row band (gsr_forward_band), then the
        # bands are gathered on every rank -- strong scaling, the gather is inside the timed region
        sharding.render_image_bands(sd, cd, kd, h, w, DMAX, gather_to=None)

    def step():
        if args.split == "bands-peer" and world > 1:
            return sharding.render_image_bands_peer(sd, cd, kd, h, w, DMAX, gather_to=0)
        if args.split == "bands" and world > 1:
            return step_bands()
        _lib.check(L.gsr_forward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), img.data_ptr(), n, h, w, 3,
                                 DMAX, 0.0, _lib.GSR_FLAG_OVERWRITE, ws.data_ptr(), ws.numel(), sptr))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, device-timed, max over ranks ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    strong = args.split in ("bands", "bands-peer") and world > 1
    value = (1 if strong else world) * mp_img / (ms_step * 1e-3)

    # ---- dominant kernel, timed alone with events through the split-phase ABI ----
    kern_ms = []
    for _ in range(min(args.steps, 30)):
        _lib.check(L.gsr_prepare(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), n, h, w, DMAX, 0.0, ws.data_ptr(),
                                 ws.numel(), sptr))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.gsr_forward_prepared(img.data_ptr(), n, h, w, 0.0, _lib.GSR_FLAG_OVERWRITE, ws.data_ptr(),
                                          ws.numel(), sptr))
        b.record()
        torch.cuda.synchronize()
        kern_ms.append(a.elapsed_time(b))
    kern_ms = sum(kern_ms) / len(kern_ms)
    alg_bytes = 32 * n + 12 * h * w
    peak, peak_src = peaks()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9

    # ---- end to end through the plugin call, host buffers ----
    # Every step copies ITS inputs host->device, renders through gscuda.gs_render (the call GSASR's
    # gswrapper makes) and copies ITS image device->host.  Consecutive steps rotate over NSTREAM
    # CUDA streams (own device buffers and pinned result buffers each), so the D2H of step i overlaps
    # the H2D + raster of step i+1 on the full-duplex PCIe link; wall clock over the whole loop.
    NSTREAM = int(os.environ.get("GSR_E2E_STREAMS", "3"))  # 3 keep the D2H engine busy: 4,290 vs 4,110 MP/s with 2
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSTREAM)]
    outs_h = [torch.empty(h, w, 3, dtype=torch.float32).pin_memory() for _ in range(NSTREAM)]

    def e2e_step(i):
        st = streams[i % NSTREAM]
        with torch.cuda.stream(st):
            a, b, cc = s_h.to(dev, non_blocking=True), c_h.to(dev, non_blocking=True), k_h.to(dev, non_blocking=True)
            o = torch.zeros(h, w, 3, device=dev)
            gscuda.gs_render(a, b, cc, o, n, h, w, 3, DMAX)
            outs_h[i % NSTREAM].copy_(o, non_blocking=True)

    for i in range(4):
        e2e_step(i)
    barrier()
    ksteps = max(6, min(args.steps, 40))
    t0 = time.perf_counter()
    for i in range(ksteps):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / ksteps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * mp_img / (e2e_ms * 1e-3)

    # ---- the same end-to-end loop with the post-processing of inference_paper.py:136-138 fused into the
    # raster kernel (gs_render_u8: clamp, x255, round, uint8 HWC): 3 bytes per pixel come back, not 12.
    # An extra figure for the inference pipeline; `e2e` above stays the reference-shaped fp32 call.
    outs_u8 = [torch.empty(h, w, 3, dtype=torch.uint8).pin_memory() for _ in range(NSTREAM)]

    def e2e_u8_step(i):
        st = streams[i % NSTREAM]
        with torch.cuda.stream(st):
            a, b, cc = s_h.to(dev, non_blocking=True), c_h.to(dev, non_blocking=True), k_h.to(dev, non_blocking=True)
            o = torch.empty(h, w, 3, dtype=torch.uint8, device=dev)
            gscuda.gs_render_u8(a, b, cc, o, n, h, w, DMAX, bgr=True)
            outs_u8[i % NSTREAM].copy_(o, non_blocking=True)

    for i in range(4):
        e2e_u8_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(ksteps):
        e2e_u8_step(i)
    barrier()
    e2e_u8_ms = (time.perf_counter() - t0) * 1e3 / ksteps
    if world > 1:
        t = torch.tensor([e2e_u8_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_u8_ms = float(t.item())

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(workload_config(args.workload), split=args.split,
                           **({"sigma": "compact: 0.99999*sigmoid(N(-1.5,0.5^2))+1e-6"} if args.compact else {})),
            "clocks": clocks, "gpu_launches": KERNELS_PER_STEP * args.steps,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 12 * h * w,
                    "note": f"gscuda.gs_render from pinned host tensors, steps rotate over {NSTREAM} CUDA streams"},
            "e2e_u8": {"value": world * mp_img / (e2e_u8_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_u8_ms,
                       "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 3 * h * w,
                       "note": "extra: gscuda.gs_render_u8 (fused clamp/x255/round/uint8 post-processing of "
                               "inference_paper.py:136-138), uint8 HWC image copied back"},
            "roofline": {"bound": "hbm", "kernel": "gsr_forward_region_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms,
                         "kernel_share_of_step": kern_ms / ms_step, "traffic": None if args.compact else TRAFFIC.get(args.workload),
                         "note": "the kernel is bound by the MUFU.EX2 / FP32 issue pipes, not by HBM: see DESIGN.md"},
        }
        if not args.no_cpu_baseline and world == 1:
            mps, cores, sample, _ = cpu_sample(args.workload)
            out["cpu_baseline"] = {"value": mps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of gsr_forward_kernel per launch, from the
# `ncu --set full` capture summarised under profiles/ (bytes); None where not captured.
TRAFFIC = {"HL": 216.8e6}  # profiles/r01_fwd_halfwarp_HL_ncu_full.txt: 150.0 MB read + 66.8 MB written

if __name__ == "__main__":
    main()
