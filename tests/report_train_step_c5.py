"""BASELINE config 5, for real: batch-32 256x256 LR -> x4 training step, the batch sharded 4 samples per GPU over
8 GPUs (torchrun, one process per GPU, NCCL).  Per rank and step:

    random encoder features (4, 64, 256, 256)
      -> the reference's OWN Fea2GS_ROPE_AMP head (utils/fea2gsropeamp.py:518-719, unmodified, staged by
         oracle/ref_py.py; bf16 autocast like gsasr_amp_model.py:208-271), optionally wrapped in DDP
      -> (4, 16*256*256, 9) raw Gaussians
      -> render forward (this library), L1 loss against random ground truth, backward through render and head,
         Adam step.

Reported per variant: ms per step (max over ranks), the render's own forward+backward time on the same tensors and
its share of the step.  Render variants: `reference-loop` = the reference's generate_2D_gaussian_splatting_step
(its own file, running on this repo's gscuda) once per sample, as gsasr_model.py:191-233 does; `batch` = one
uniform-batch call of this library (one set-up + one raster launch each way); `batch-fused` = the same with the front
end folded into the set-up kernel and the L1 loss + its gradient as one kernel (losses.l1_crop_loss_padded).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/report_train_step_c5.py [--ddp] [--steps 5]
    python tests/report_train_step_c5.py --per-gpu 1 --depth tiny          (single GPU smoke run)
"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import gaussian_splatting as gsp
from gsasr_b200 import losses
from oracle import ref_py

ap = argparse.ArgumentParser()
ap.add_argument("--per-gpu", type=int, default=4)
ap.add_argument("--lr", type=int, default=256)
ap.add_argument("--scale", type=float, default=4.0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--dmax", type=float, default=0.5, help="training YAMLs use 0.5")
ap.add_argument("--ddp", action="store_true")
ap.add_argument("--no-checkpoint", action="store_true",
                help="the head's own activation checkpointing (use_checkpoint, fea2gsropeamp.py:520) is ON by default: "
                     "4 samples of 16 x 256^2 Gaussians do not fit 180 GB without it")
ap.add_argument("--depth", default="default", choices=["default", "yaml", "tiny"],
                help="head depth: class defaults (1x2 cross, 6x6 self), the HATL YAML's (4x4, 8x6), or 1x1/1x1")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev); dist.barrier()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

head_mod = ref_py.load("utils.fea2gsropeamp")
ref_gsp = ref_py.load("utils.gaussian_splatting")
depth = {"default": dict(num_crossattn_blocks=1, num_crossattn_layers=2, num_selfattn_blocks=6, num_selfattn_layers=6),
         "yaml": dict(num_crossattn_blocks=4, num_crossattn_layers=4, num_selfattn_blocks=8, num_selfattn_layers=6),
         "tiny": dict(num_crossattn_blocks=1, num_crossattn_layers=1, num_selfattn_blocks=1, num_selfattn_layers=1)}[args.depth]
torch.manual_seed(0)
head = head_mod.Fea2GS_ROPE_AMP(inchannel=64, channel=192, num_heads=6, num_gs_seed=256, window_size=16,
                                shuffle_scale1=2, shuffle_scale2=2, use_checkpoint=not args.no_checkpoint, **depth).to(dev)
model = torch.nn.parallel.DistributedDataParallel(head, device_ids=[local]) if (args.ddp and world > 1) else head
opt = torch.optim.Adam(model.parameters(), lr=1e-5)
B, lr, sc = args.per_gpu, args.lr, args.scale
H = W = int(lr * sc)
g = torch.Generator(dev).manual_seed(100 + rank)
feats = torch.randn(B, 64, lr, lr, device=dev, generator=g)
gt = torch.rand(B, 3, H, W, device=dev, generator=g)
scale_vec = torch.full((B,), sc, device=dev)
sr_size, scale_modify = torch.tensor([H, W]), torch.tensor([sc, sc])


def render_reference_loop(params):  # the reference's own front end, once per sample (gsasr_model.py:191-233)
    outs = [ref_gsp.generate_2D_gaussian_splatting_step(sr_size=sr_size, gs_parameters=params[i], scale=sc,
                                                        scale_modify=scale_modify, default_step_size=1.2,
                                                        cuda_rendering=True, mode="scale_modify", if_dmax=True,
                                                        dmax_mode="fix", dmax=args.dmax).unsqueeze(0) for i in range(B)]
    return torch.cat(outs)


def render_batch(params):  # this library: the whole share in one set-up + one raster launch each way
    return gsp.generate_2D_gaussian_splatting_step_batch(sr_size, params, sc, scale_modify, dmax=args.dmax)


def head_forward():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return model(feats, scale_vec).float()


def ev(fn, steps):
    fn(); fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
    return ms


out = {"config": f"C5: {world} GPU(s) x {B} samples, {lr}x{lr} LR -> x{sc:g} ({H}x{W}), Fea2GS_ROPE_AMP head "
                 f"(window 16, 256 seeds, shuffle 2x2 = 16 Gaussians per LR pixel: {16 * lr * lr} per sample; depth "
                 f"'{args.depth}' {depth}, use_checkpoint={not args.no_checkpoint}), bf16 autocast, dmax {args.dmax}, "
                 f"ddp={bool(args.ddp and world > 1)}",
       "n_gpus": world, "global_batch": world * B}
with torch.no_grad():
    n_per = int(head_forward().shape[1])
out["gaussians_per_sample"] = n_per
def render_batch_fused(params):  # + the front end folded into the set-up kernel
    return gsp.generate_2D_gaussian_splatting_step_batch(sr_size, params, sc, scale_modify, dmax=args.dmax, fused=True)


def loss_torch(sr):
    return (sr - gt).abs().mean()


def loss_fused(sr):  # crop + L1 + gradient in one kernel (equal sizes: the crop is the whole image)
    return losses.l1_crop_loss_padded(sr, gt, [(H, W)] * B)


for name, render, loss_fn in (("reference-loop", render_reference_loop, loss_torch), ("batch", render_batch, loss_torch),
                              ("batch-fused", render_batch_fused, loss_fused)):
    def step():
        opt.zero_grad(set_to_none=True)
        params = head_forward()
        loss = loss_fn(render(params))
        loss.backward()
        opt.step()
    params0 = head_forward().detach()

    def render_only():
        p = params0.clone().requires_grad_(True)
        loss_fn(render(p)).backward()

    def head_only():
        opt.zero_grad(set_to_none=True)
        head_forward().sum().backward()
        opt.step()
    t_step, t_render = ev(step, args.steps), ev(render_only, args.steps)
    out[name] = {"step_ms": round(t_step, 2), "render_fwd_bwd_ms": round(t_render, 2),
                 "render_share": round(t_render / t_step, 3),
                 "samples_per_s": round(world * B / (t_step * 1e-3), 2)}
out["head_only_step_ms"] = round(ev(head_only, args.steps), 2)
out["max_memory_gb"] = round(torch.cuda.max_memory_allocated() / 2**30, 1)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
