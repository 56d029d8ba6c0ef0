"""BASELINE config 4 shaped measurement: 1024x1024 LR -> x4 through split_and_joint_image (tile 480,
overlap 8, crop 4 -> 3x3 LR tiles, each 921,600 Gaussians -> 1920x1920 HR), the tiles sharded over the
ranks, gathered to rank 0 and stitched.  The encoder and the fea2gs head are out of scope: model_g is
the identity and model_fea2gs returns seeded random raw (N,9) tensors of the head's shape.

  python tools/c4_tiles_bench.py [--steps 10]
  torchrun --nproc-per-node N tools/c4_tiles_bench.py
"""
import argparse, json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields
from gsasr_b200.split_and_joint_image import split_and_joint_image, plan_tiles
from gsasr_b200.sharding import shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--direct", action="store_true", help="tiles written straight into the (peer) canvas")
ap.add_argument("--fused", action="store_true", help="with --direct: fused front end")
ap.add_argument("--check", action="store_true", help="compare with the tile-buffer route")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
scale, split, overlap, crop = 4.0, 480, 8, 4
lq = torch.rand(1, 3, 1024, 1024, device=dev)
plan = plan_tiles(1024, 1024, scale, split, overlap)
lo, hi = shard_range(plan.n, rank, world)
raws = {i: fields.raw_field(2 * split, 2 * split, seed=i).to(dev) for i in range(lo, hi)}   # 4 Gaussians per LR pixel
order = list(range(lo, hi))
calls = [0]
def model_fea2gs(feat, scale_vector):
    i = order[calls[0] % len(order)]; calls[0] += 1
    return raws[i].unsqueeze(0)
sm = torch.tensor([scale, scale])
def step():
    return split_and_joint_image(lq, scale, split, overlap, lambda t: t, model_fea2gs, sm, crop_size=crop,
                                 if_dmax=True, dmax=0.1, gather_to=0, direct=args.direct, fused=args.fused)
if args.check:
    calls[0] = 0
    ref = split_and_joint_image(lq, scale, split, overlap, lambda t: t, model_fea2gs, sm, crop_size=crop,
                                if_dmax=True, dmax=0.1, gather_to=0, direct=False)
    calls[0] = 0
    got = step()
    if rank == 0:
        print("direct=%s vs tile buffers: max-abs %.2e" % (args.direct, float((got - ref).abs().max())), flush=True)
    calls[0] = 0
for _ in range(3): out = step()
if world > 1: dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.steps): out = step()
b.record()
if world > 1: dist.barrier()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
if world > 1:
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
if rank == 0:
    h, w = out.shape[-2:]
    print(json.dumps({"config": "C4-shaped: 1024x1024 LR -> x4 via split_and_joint_image, %dx%d tiles of %d^2 LR, %d Gaussians each"
                      % (plan.tiles_h, plan.tiles_w, split, raws[lo].shape[0]), "world": world, "direct": args.direct, "fused": args.fused, "sr": [h, w],
                      "ms_per_image": round(ms, 3), "mp_per_s": round(h * w / 1e6 / (ms * 1e-3), 1)}))
if world > 1: dist.destroy_process_group()
