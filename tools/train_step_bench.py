"""BASELINE config 5 shaped measurement: a training-step render of a batch of B samples
(256x256 LR -> x4, 262,144 Gaussians each = EDSR-baseline fea2gs shapes), forward + backward
through the front-end mirror with autograd, samples dealt to the ranks in blocks (4 per GPU at
B = 32 on 8 GPUs).  The fea2gs head is out of scope: its output is replaced by seeded random raw
(N,9) tensors.  dmax = 0.5 as in the training YAMLs, plus 0.1.

  python tools/train_step_bench.py [--batch 32] [--steps 20]            (1 GPU: renders its share)
  torchrun --nproc-per-node N tools/train_step_bench.py --batch 32
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, sharding
from gsasr_b200 import gaussian_splatting as gsp

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--world", type=int, default=0, help="pretend world size when run on one GPU (share = batch/world)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
eff_world = args.world or world
lo, hi = sharding.shard_range(args.batch, rank if world > 1 else 0, eff_world)
cfg = fields.CONFIGS["C2"]
h, w = cfg.hr
raws = [fields.raw_field(*cfg.grid, seed=i).to(dev) for i in range(lo, hi)]
gts = [torch.rand(3, h, w, device=dev) for _ in range(lo, hi)]
out = {}
for dmax in (0.5, 0.1):
    for fused in (False, True):
        def step():
            loss = 0.0
            for raw, gt in zip(raws, gts):
                p = raw.clone().requires_grad_(True)
                img = gsp.generate_2D_gaussian_splatting_step(torch.tensor([h, w]), p, cfg.scale, torch.tensor([cfg.scale] * 2), dmax=dmax, fused=fused)
                l = (img - gt).abs().mean()          # L1, as gsasr_model.py:213-229
                l.backward()
                loss += float(l.detach()) if False else 0.0
        for _ in range(3): step()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps): step()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
        out[f"dmax{dmax}_{'fused' if fused else 'torchops'}_ms_per_step"] = round(ms, 3)
    # the whole share as ONE uniform batch: one set-up + one raster launch each way (ground truth kept
    # channels-last like the render, so the loss needs no transpose)
    rawb, gtb = torch.stack(raws), torch.stack(gts).contiguous(memory_format=torch.channels_last)
    for fused in (False, True):
        def step_batch():
            p = rawb.clone().requires_grad_(True)
            img = gsp.generate_2D_gaussian_splatting_step_batch(torch.tensor([h, w]), p, cfg.scale, torch.tensor([cfg.scale] * 2), dmax=dmax, fused=fused)
            ((img - gtb).abs().mean(dim=(1, 2, 3))).sum().backward()
        for _ in range(3): step_batch()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps): step_batch()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
        out[f"dmax{dmax}_batch_{'fused' if fused else 'torchops'}_ms_per_step"] = round(ms, 3)
# ragged: every sample at its own scale in [3.5, 4] (HR 896..1024), the reference loop's shape, vs ONE padded launch
import torch.nn.functional as F
scs = [3.5 + 0.5 * ((i * 7) % 8) / 7.0 for i in range(len(raws))]
szs = [(int(256 * sc) // 8 * 8, int(256 * sc) // 8 * 8) for sc in scs]
gtp = torch.rand(len(raws), 3, 1024, 1024, device=dev).contiguous(memory_format=torch.channels_last)
def step_loop_ragged():
    for raw, gt, sc, (hh, ww) in zip(raws, gtp, scs, szs):
        p = raw.clone().requires_grad_(True)
        img = gsp.generate_2D_gaussian_splatting_step(torch.tensor([hh, ww]), p, sc, torch.tensor([sc] * 2), dmax=0.1, fused=True)
        (img - gt[:, :hh, :ww]).abs().mean().backward()
def step_padded(fused):
    p = rawb.clone().requires_grad_(True)
    img = gsp.generate_2D_gaussian_splatting_step_batch_padded([torch.tensor(z) for z in szs], p, scs, dmax=0.1, hmax=1024, wmax=1024, fused=fused)
    ((img - gtp).abs().mean(dim=(1, 2, 3))).sum().backward()
for name, fn in (("ragged_loop_fused", step_loop_ragged), ("ragged_padded_batch_torchops", lambda: step_padded(False)),
                 ("ragged_padded_batch_fused", lambda: step_padded(True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps): fn()
    b.record(); torch.cuda.synchronize()
    out[f"{name}_ms_per_step"] = round(a.elapsed_time(b) / args.steps, 3)
if rank == 0:
    print(json.dumps({"config": "C5-shaped: batch %d x (256x256 LR -> x4, 262144 Gaussians), fwd+bwd render + L1" % args.batch,
                      "samples_per_gpu": hi - lo, "world": eff_world, **out}))
if world > 1: dist.destroy_process_group()
