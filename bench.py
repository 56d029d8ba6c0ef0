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
and the image resident in HBM.  What the JSON line carries besides the contract's keys:

  e2e            the same metric through the plugin call a GSASR user makes (gscuda.gs_render) from pinned HOST
                 buffers, host<->device copies inside the timed region; every rank bound to its GPU's NUMA node.
  roofline       the dominant kernel (gsr_forward_region_kernel), timed with CUDA events on the launch stream via
                 the split-phase C ABI; `step_breakdown` gives set-up / raster / backward beside it.
  sustained      the same step looped for >= 1 s: ms per step and the clocks sampled under that load.
  strong         (N > 1) ONE image over the N GPUs in row bands: all-gather and peer-store forward, backward with
                 its all-reduce, each against the single-GPU time measured in the same run.
  cpu_baseline   the CPU restatement of the reference algorithm (oracle/, OpenMP over all host cores) on a bounded
                 sample; `cpu_reference_python` = the reference's OWN PyTorch-CPU renderer (rendering_python,
                 unmodified, BASELINE config 1) timed on the same host cores.
  gpu_reference  the reference's own CUDA kernels (oracle/_ref, rebuilt for sm_100a) on BASELINE config 2, same box.
  head_tail      (beside the headline) the five MLPs that produce the headline field's 2,097,152 raw
                 Gaussians from (N, 192) features, as ONE tcgen05 kernel: ms, TFLOP/s, fraction of the measured bf16
                 tensor peak, and torch's bf16-autocast time for the same modules.
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
KERNELS_PER_STEP = 4  # table, region_build2, forward_region + the device-guarded fallback (home-bin set-up + raster in one kernel)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores() -> int:
    """Cores this process may use: the affinity mask if there is one, else the machine's count.  torchrun
    exports OMP_NUM_THREADS=1, which must not decide how many cores the CPU arms are given."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def numa_memory_policy(node):
    """set_mempolicy(2) for the calling thread: MPOL_PREFERRED on `node` (None: back to MPOL_DEFAULT).  Placement of the
    pinned staging buffers without touching any thread's CPU affinity.  Returns True if the kernel took it."""
    try:
        import ctypes, platform

        nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
        if nr is None:
            return False
        libc = ctypes.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(nr, 0, None, 0) == 0
        mask = ctypes.c_ulong(1 << int(node))
        return libc.syscall(nr, 1, ctypes.byref(mask), 64) == 0
    except Exception:
        return False


def bind_to_gpu_numa(local: int, bind: bool = True):
    """Pin this rank's threads to the NUMA node its GPU hangs off, BEFORE the pinned host buffers are allocated
    (first touch places them on that node).  bind=False (a single rank, whose CPU arms should keep every core): only
    the memory policy of the calling thread prefers that node while the pinned buffers are allocated.
    Best effort: returns what was done for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        how = "sysfs"
        if node < 0:
            # virtualised boxes hide the PCI device's node: NVML's own CPU affinity for the GPU, else an even split of
            # the GPUs over the host's memory nodes (GPUs 0..n/2-1 on node 0, ... -- the usual 8-GPU board)
            nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
            try:
                words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
                aff = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
            except Exception:
                aff = set()
            node = None
            if aff and len(aff) < (os.cpu_count() or 1):
                for nd in nodes:
                    cl = set()
                    for part in open(f"/sys/devices/system/node/node{nd}/cpulist").read().strip().split(","):
                        a, _, b = part.partition("-")
                        cl.update(range(int(a), int(b or a) + 1))
                    if aff <= cl:
                        node, how = nd, "nvml cpu affinity"
            if node is None:
                if len(nodes) < 2:
                    return {"numa_node": None, "note": "no NUMA information for this GPU (one memory node visible)"}
                ngpu = max(pynvml.nvmlDeviceGetCount(), 1)
                node, how = nodes[min(local * len(nodes) // ngpu, len(nodes) - 1)], "assumed: GPUs split evenly over the memory nodes"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if not bind:
            ok = numa_memory_policy(node)
            return {"numa_node": node, "how": how, "threads": "not pinned",
                    "memory_policy": "preferred on this node while the pinned buffers are allocated" if ok else "unchanged (set_mempolicy refused)"}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "how": how}
    except Exception as exc:
        return {"numa_node": None, "note": f"not bound: {exc}"}


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

    def __init__(self, cfg_name: str):
        import numpy as np

        from gsasr_b200 import fields
        from oracle import oracle

        oracle.build()
        oracle.set_num_threads(host_cores())  # explicitly: torchrun's OMP_NUM_THREADS=1 does not apply to this arm
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


def cpu_reference_python():
    """The reference's OWN CPU path: utils/gaussian_splatting.py rendering_python (cuda_rendering=False), unmodified
    (staged by oracle/ref_py.py), on BASELINE config 1 (64x64 LR -> x2, 16,384 Gaussians, 128x128), all host cores.
    A different function of the parameters than the CUDA path: a timing comparator, not a parity one."""
    try:
        import torch

        from gsasr_b200 import fields
        from oracle import ref_py

        if not ref_py.have():
            return {"unavailable": "oracle/_ref/ref_py not staged (built where /root/reference exists)"}
        gsp = ref_py.load("utils.gaussian_splatting")
        cores = host_cores()
        torch.set_num_threads(cores)
        cfg = fields.CONFIGS["C1"]
        raw = fields.raw_field(*cfg.grid, seed=0)
        h, w = cfg.hr
        args = dict(sr_size=torch.tensor([h, w]), scale=cfg.scale, scale_modify=torch.tensor([cfg.scale, cfg.scale]),
                    cuda_rendering=False)
        ts = []
        for i in range(4):  # 1 warm-up + 3
            t0 = time.perf_counter()
            gsp.generate_2D_gaussian_splatting_step(gs_parameters=raw.clone(), **args)
            ts.append(time.perf_counter() - t0)
        t = sorted(ts[1:])[1]
        return {"value": h * w / 1e6 / t, "unit": UNIT, "seconds_per_call": t, "cores": cores, "kind": "reference",
                "config": "C1: 64x64 LR -> x2, 16,384 Gaussians, 128x128 (BASELINE configs[0])",
                "sample": "whole image, median of 3 calls after 1 warm-up, torch CPU threads = cores"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def head_tail_bench(dev):
    """gsr_head_tail_forward at the headline field's size (1024 x 2048 Gaussian grid, C = 192) beside torch running the
    same five nn.Sequential MLPs under bf16 autocast (a quarter of the rows, extrapolated: the unfused path
    materialises rows x 768 activations per head)."""
    try:
        import torch
        import torch.nn as nn

        from gsasr_b200 import head_tail

        c, gh, gw = 192, 1024, 2048
        torch.manual_seed(0)
        blks = [nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, 4 * c), nn.ReLU(), nn.Linear(4 * c, k)).to(dev)
                for k in (2, 1, 1, 3, 2)]
        pk = head_tail.PackedHeadTail(blks, dev)
        q = torch.randn(1, gh, gw, c, device=dev, dtype=torch.bfloat16)

        def ev(fn, reps):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        ms = ev(lambda: head_tail.fused_head_tail(q, pk), 5)
        flop = 2.0 * gh * gw * (5 * (c * c + c * 4 * c) + 4 * c * 9)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            qs = q[:, :gh // 4]
            ref_ms = 4 * ev(lambda: [blk(qs) for blk in blks], 3)
        peak = None
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
        tf = flop / (ms * 1e-3) / 1e12
        return {"rows": gh * gw, "channels": c, "ms": ms, "tflops": tf, "flop": flop,
                "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s",
                             "frac": (tf / peak) if peak else None,
                             "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst)"},
                "torch_bf16_autocast_ms": ref_ms,
                "note": "gsr_head_tail_forward: TMA weight ring -> tcgen05.mma (bf16, fp32 in TMEM) -> epilogue "
                        "(ReLU, next operand in shared memory, last Linear on the CUDA cores); includes the bf16 "
                        "cast of the features"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def gpu_reference(dev):
    """The reference's own CUDA kernels (oracle/_ref/libgsref_dmax.so: gs_cuda_dmax rebuilt for sm_100a, unmodified) on
    BASELINE config 2 (256x256 LR -> x4, 262,144 Gaussians, 1024x1024), beside this library on the same tensors."""
    try:
        import torch

        from gsasr_b200 import fields, gscuda
        from oracle import oracle

        if not oracle.have_ref():
            return {"unavailable": "oracle/_ref not built"}
        _, s, c, k, h, w = fields.make("C2", 0)
        sd, cd, kd = s.to(dev), c.to(dev), k.to(dev)
        n = sd.shape[0]
        R = oracle.RefKernels(True)
        g = torch.rand(h, w, 3, device=dev)

        def wall(fn, reps):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e3

        ref_f = wall(lambda: R.forward(sd, cd, kd, torch.zeros(h, w, 3, device=dev), DMAX), 2)
        ref_b = wall(lambda: R.backward(sd, cd, kd, g, torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd), DMAX), 1)
        img = torch.zeros(h, w, 3, device=dev)
        ws = gscuda.workspace(n, h, w, dev)
        gs, gc, gk = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)
        our_f = wall(lambda: gscuda.gs_render(sd, cd, kd, img, n, h, w, 3, DMAX, workspace_buf=ws), 50)
        our_b = wall(lambda: gscuda.gs_render_backward(sd, cd, kd, g, gs, gc, gk, n, h, w, 3, DMAX, workspace_buf=ws), 50)
        return {"config": "C2: 256x256 LR -> x4, 262,144 Gaussians, 1024x1024, dmax 0.1 (BASELINE configs[1])",
                "reference_fwd_ms": ref_f, "reference_bwd_ms": ref_b, "ours_fwd_ms": our_f, "ours_bwd_ms": our_b,
                "reference_fwd_mps": h * w / 1e6 / (ref_f * 1e-3), "ours_fwd_mps": h * w / 1e6 / (our_f * 1e-3),
                "note": "gscuda-level calls (accumulate contract), wall clock around synchronised calls; the reference "
                        "walks the exact dmax window, this library also culls at k = 5 sigma"}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


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
        "cpu_reference_python": cpu_reference_python(),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's own implementation of this path at this shape is CUDA-only (gs_cuda_dmax); its CPU arm "
                "is the oracle port of that algorithm on the host cores (all of them, whatever OMP_NUM_THREADS says); a "
                "step renders the bounded sample described in cpu_baseline.sample and value extrapolates it to the "
                "whole image.  cpu_reference_python times the reference's PyTorch-CPU renderer on the one config it "
                "can hold in memory (BASELINE configs[0])",
    }
    print(json.dumps(out), flush=True)


def workload_config(name):
    from gsasr_b200 import _lib, fields
    cfg = fields.CONFIGS[name]
    h, w = cfg.hr
    try:
        ws_mb = _lib.load().gsr_workspace_bytes(cfg.n, h, w) / 1e6
    except Exception:
        ws_mb = float("nan")
    return {"workload": name, "lr": [cfg.lr_h, cfg.lr_w], "scale": cfg.scale, "hr": [h, w],
            "gaussians": cfg.n, "dmax": DMAX, "sigma": "model-like: 0.99999*sigmoid(N(0,1))+1e-6",
            "ksigma": "library default (5)",
            "image_mode": "value: GSR_FLAG_OVERWRITE (one plain store per pixel, what the front-end mirror uses); "
                          "`accumulate` and `e2e` use the reference's accumulate-into-rendered_img contract",
            "l2": f"per-step working set (params {32 * cfg.n / 1e6:.0f} MB + image {12 * h * w / 1e6:.0f} MB + workspace "
                  f"{ws_mb:.0f} MB) exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="HL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_reference / strong / sustained (quick runs)")
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

    numa = bind_to_gpu_numa(local, bind=world > 1 or bool(os.environ.get("GSR_BIND_NUMA")))

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

    def fwd(flags):
        _lib.check(L.gsr_forward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), img.data_ptr(), n, h, w, 3,
                                 DMAX, 0.0, flags, ws.data_ptr(), ws.numel(), sptr))

    def step():
        if args.split == "bands-peer" and world > 1:
            return sharding.render_image_bands_peer(sd, cd, kd, h, w, DMAX, gather_to=0)
        if args.split == "bands" and world > 1:
            return step_bands()
        fwd(_lib.GSR_FLAG_OVERWRITE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world > 1:
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def timed(fn, reps, warm=3):
        """Device time of `reps` back-to-back calls (CUDA events on the launch stream), ms per call."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

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
    ms_total = maxr(ms_total)
    ms_step = ms_total / args.steps
    strong_main = args.split in ("bands", "bands-peer") and world > 1
    value = (1 if strong_main else world) * mp_img / (ms_step * 1e-3)

    # ---- the same step for >= 1 s: the clocks under sustained load (the K-step region above lasts milliseconds)
    sustained = None
    if not args.no_extras:
        reps = max(args.steps, int(1200.0 / max(ms_step, 1e-3)))
        s2 = ClockSampler(local)
        if rank == 0:
            s2.start()
        barrier()
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        barrier()
        sus_ms = maxr(e0.elapsed_time(e1)) / reps
        sustained = {"steps": reps, "ms_per_step": sus_ms,
                     "value": (1 if strong_main else world) * mp_img / (sus_ms * 1e-3), "unit": UNIT,
                     "clocks": s2.stop() if rank == 0 else None}

    # ---- the reference's contract: accumulate into the caller's image (RED instead of plain stores)
    acc_ms = maxr(timed(lambda: fwd(0), min(args.steps, 50)))

    # ---- dominant kernel, timed alone with events through the split-phase ABI ----
    def prepare():
        _lib.check(L.gsr_prepare(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), n, h, w, DMAX, 0.0, ws.data_ptr(),
                                 ws.numel(), sptr))

    def raster():
        _lib.check(L.gsr_forward_prepared(img.data_ptr(), n, h, w, 0.0, _lib.GSR_FLAG_OVERWRITE, ws.data_ptr(),
                                          ws.numel(), sptr))

    prepare()
    kern_ms = timed(raster, min(args.steps, 30))
    prep_ms = timed(prepare, min(args.steps, 30))  # region build + the home-bin sort the backward needs
    grd = torch.rand(h, w, 3, device=dev)
    gs_, gc_, gk_ = torch.zeros_like(sd), torch.zeros_like(cd), torch.zeros_like(kd)

    def bwd():
        _lib.check(L.gsr_backward(sd.data_ptr(), cd.data_ptr(), kd.data_ptr(), grd.data_ptr(), gs_.data_ptr(),
                                  gc_.data_ptr(), gk_.data_ptr(), n, h, w, 3, DMAX, 0.0, 0, ws.data_ptr(), ws.numel(),
                                  sptr))

    bwd_ms = timed(bwd, min(args.steps, 20))
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

    def e2e_loop(fn):
        for i in range(4):
            fn(i)
        barrier()
        ksteps = max(6, min(args.steps, 40))
        t0 = time.perf_counter()
        for i in range(ksteps):
            fn(i)
        barrier()
        return maxr((time.perf_counter() - t0) * 1e3 / ksteps)

    e2e_ms = e2e_loop(e2e_step)
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

    e2e_u8_ms = e2e_loop(e2e_u8_step)

    # ---- host<->device link of this rank, alone and with all ranks copying at once (what limits e2e scaling)
    def copy_gbs():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            s_h.to(dev, non_blocking=True), c_h.to(dev, non_blocking=True), k_h.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        up = 5 * 32 * n / (time.perf_counter() - t0) / 1e9
        t0 = time.perf_counter()
        for _ in range(5):
            outs_h[0].copy_(img, non_blocking=True)
        torch.cuda.synchronize()
        return up, 5 * 12 * h * w / (time.perf_counter() - t0) / 1e9

    barrier()
    h2d_gbs, d2h_gbs = copy_gbs()
    if world > 1:
        t = torch.tensor([h2d_gbs, d2h_gbs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        h2d_gbs, d2h_gbs = float(t[0]), float(t[1])

    # ---- strong scaling: ONE image over the N GPUs in row bands ----
    strong = None
    if world > 1 and not args.no_extras:
        _, s0, c0, k0, _, _ = fields.make(cfg, seed=0)  # the same field on every rank
        s0, c0, k0 = s0.to(dev), c0.to(dev), k0.to(dev)
        g0 = torch.rand(h, w, 3, device=dev, generator=torch.Generator(dev).manual_seed(7))

        def coll(fn, reps=20):
            for _ in range(3):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            barrier()
            return maxr((time.perf_counter() - t0) * 1e3 / reps)

        def dev_ms(fn, reps=20):
            """device time per call: CUDA events on the launch stream around `reps` calls, max over ranks"""
            barrier()
            return maxr(timed(fn, reps))

        one_f = maxr(timed(lambda: gscuda.gs_render(s0, c0, k0, img, n, h, w, 3, DMAX, flags=1, workspace_buf=ws), 20))
        gsz = [torch.zeros_like(s0), torch.zeros_like(c0), torch.zeros_like(k0)]
        one_b = maxr(timed(lambda: gscuda.gs_render_backward(s0, c0, k0, g0, *gsz, n, h, w, 3, DMAX, workspace_buf=ws), 10))
        ag = coll(lambda: sharding.render_image_bands(s0, c0, k0, h, w, DMAX, gather_to=None))
        try:
            pr = coll(lambda: sharding.render_image_bands_peer(s0, c0, k0, h, w, DMAX, gather_to=0))
        except Exception as exc:  # symmetric memory unavailable: report, do not fail the bench
            pr = None
            peer_note = f"{type(exc).__name__}: {exc}"
        try:
            pr8 = coll(lambda: sharding.render_image_bands_peer(s0, c0, k0, h, w, DMAX, gather_to=0, u8=True))
            pr8_dev = dev_ms(lambda: sharding.render_image_bands_peer(s0, c0, k0, h, w, DMAX, gather_to=0, u8=True))
            u8img = torch.empty(h, w, 3, dtype=torch.uint8, device=dev)
            one_u8 = maxr(timed(lambda: gscuda.gs_render_u8(s0, c0, k0, u8img, n, h, w, DMAX, workspace_buf=ws), 20))
        except Exception as exc:
            pr8 = pr8_dev = one_u8 = None
            peer_note = f"{type(exc).__name__}: {exc}"
        bb = coll(lambda: sharding.backward_image_bands(s0, c0, k0, g0, h, w, DMAX), 10)
        ag_dev = dev_ms(lambda: sharding.render_image_bands(s0, c0, k0, h, w, DMAX, gather_to=None))
        pr_dev = dev_ms(lambda: sharding.render_image_bands_peer(s0, c0, k0, h, w, DMAX, gather_to=0)) if pr else None
        bb_dev = dev_ms(lambda: sharding.backward_image_bands(s0, c0, k0, g0, h, w, DMAX), 10)
        r0_, rows_ = sharding.band_rows(h, rank, world)
        band_img = torch.zeros(rows_, w, 3, device=dev)
        band_only = dev_ms(lambda: gscuda.gs_render_band(s0, c0, k0, band_img, n, h, w, 3, r0_, rows_, DMAX, flags=1))
        strong = {
            "what": f"ONE {h}x{w} image ({n} Gaussians) split into {world} row bands; wall clock per call incl. the "
                    "collective and the stream synchronisation, max over ranks",
            "single_gpu_fwd_ms": one_f, "single_gpu_bwd_ms": one_b,
            "bands_allgather_fwd_ms": ag, "bands_allgather_mps": mp_img / (ag * 1e-3),
            "bands_allgather_efficiency": one_f / (world * ag),
            "bands_peer_fwd_ms": pr, "bands_peer_mps": (mp_img / (pr * 1e-3)) if pr else None,
            "bands_peer_efficiency": (one_f / (world * pr)) if pr else None,
            "bands_peer_u8": None if pr8 is None else {
                "what": "the inference result: the (h,w,3) uint8 image of inference_paper.py:136-138 written by the band "
                        "kernels into the stitching rank's memory (3 bytes per pixel over NVLink)",
                "single_gpu_u8_fwd_ms": one_u8, "fwd_ms": pr8, "device_ms": pr8_dev,
                "efficiency": one_u8 / (world * pr8), "device_efficiency": one_u8 / (world * pr8_dev),
                "peer_store_bytes_per_rank": 3 * h * w // world},
            "bands_bwd_allreduce_ms": bb, "bands_bwd_efficiency": one_b / (world * bb),
            "device_time_ms": {"band_kernels_only": band_only, "bands_allgather": ag_dev, "bands_peer": pr_dev,
                               "bands_bwd_allreduce": bb_dev,
                               "note": "CUDA-event time on the launch stream (max over ranks): what the GPUs need; the "
                                       "wall-clock figures above add the host's launch path of a sub-millisecond call"},
            "device_efficiency": {"band_kernels_only": one_f / (world * band_only),
                                  "bands_allgather": one_f / (world * ag_dev),
                                  "bands_peer": (one_f / (world * pr_dev)) if pr_dev else None,
                                  "bands_bwd_allreduce": one_b / (world * bb_dev)},
            "collectives": {"forward_allgather_bytes_per_rank": 12 * h * w // world,
                            "forward_peer_store_bytes_per_rank": 12 * h * w // world,
                            "backward_allreduce_bytes": 32 * n},
        }
        if pr is None:
            strong["bands_peer_note"] = peer_note

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if strong_main else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.workload), split=args.split,
                           **({"sigma": "compact: 0.99999*sigmoid(N(-1.5,0.5^2))+1e-6"} if args.compact else {})),
            "clocks": clocks, "gpu_launches": KERNELS_PER_STEP * args.steps,
            "accumulate": {"ms_per_step": acc_ms, "value": world * mp_img / (acc_ms * 1e-3), "unit": UNIT,
                           "note": "gsr_forward with flags = 0: the reference's accumulate-into-rendered_img contract "
                                   "(vector RED per pixel pair instead of plain stores)"},
            "sustained": sustained,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 12 * h * w,
                    "h2d_gbs_per_rank_min": h2d_gbs, "d2h_gbs_per_rank_min": d2h_gbs, "numa": numa,
                    "note": f"gscuda.gs_render from pinned host tensors, steps rotate over {NSTREAM} CUDA streams"},
            "e2e_u8": {"value": world * mp_img / (e2e_u8_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_u8_ms,
                       "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 3 * h * w,
                       "note": "the inference pipeline's e2e: gscuda.gs_render_u8 (fused clamp/x255/round/uint8 "
                               "post-processing of inference_paper.py:136-138), uint8 HWC image copied back"},
            "roofline": {"bound": "hbm", "kernel": "gsr_forward_region_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms,
                         "kernel_share_of_step": kern_ms / ms_step, "traffic": None if args.compact else TRAFFIC.get(args.workload),
                         "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / peak,
                         "note": "the kernel is bound by the MUFU.EX2 / FP32 issue pipes, not by HBM: see DESIGN.md"},
            "step_breakdown": {"raster_ms": kern_ms, "prepare_ms": prep_ms, "forward_step_ms": ms_step,
                               "backward_call_ms": bwd_ms,
                               "note": "prepare = region build + home-bin sort (gsr_prepare); backward_call = region "
                                       "build + gsr_backward_region_kernel + chain-rule kernel (gsr_backward), "
                                       "B_bwd = 64 N + 12 H W bytes",
                               "backward_frac": (64 * n + 12 * h * w) / (bwd_ms * 1e-3) / 1e9 / peak},
        }
        if strong is not None:
            out["strong"] = strong
        if world == 1 and not args.no_extras:
            out["gpu_reference"] = gpu_reference(dev)
            out["head_tail"] = head_tail_bench(dev)
        if not args.no_cpu_baseline and world == 1:
            numa_memory_policy(None)  # the CPU arms place their memory as usual
            mps, cores, sample, _ = cpu_sample(args.workload)
            out["cpu_baseline"] = {"value": mps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
            out["cpu_reference_python"] = cpu_reference_python()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of gsr_forward_region_kernel per launch, from the
# `ncu --set full` capture summarised under profiles/ (bytes); None where not captured.
TRAFFIC = {"HL": 276.6e6}  # profiles/r02_fwd_final_HL_ncu_full.txt: 213.5 MB read + 63.1 MB written

if __name__ == "__main__":
    main()
