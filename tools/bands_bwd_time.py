"""torchrun --nproc-per-node N tools/bands_bwd_time.py : the backward of ONE headline image split into N row bands
(band kernels + one all-reduce of the parameter gradients), device time per call, max over ranks; rank 0 prints JSON."""
import os, sys, json
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda, sharding

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
_, s, c, k, h, w = fields.make(sys.argv[1] if len(sys.argv) > 1 else "HL", 0)
s, c, k = s.to(dev), c.to(dev), k.to(dev)
n = s.shape[0]
g = torch.rand(h, w, 3, device=dev, generator=torch.Generator(dev).manual_seed(1))

def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

gz = [torch.zeros_like(s), torch.zeros_like(c), torch.zeros_like(k)]
ws = gscuda.workspace(n, h, w, dev)
one = timed(lambda: gscuda.gs_render_backward(s, c, k, g, *gz, n, h, w, 3, 0.1, workspace_buf=ws), 10)
bands = timed(lambda: sharding.backward_image_bands(s, c, k, g, h, w, 0.1), 10) if world > 1 else one
if rank == 0:
    print(json.dumps({"world": world, "image": [h, w], "gaussians": n, "single_gpu_bwd_ms": one,
                      "bands_bwd_allreduce_device_ms": bands, "efficiency": one / (world * bands)}))
if world > 1:
    dist.destroy_process_group()
