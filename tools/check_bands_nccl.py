"""torchrun --nproc-per-node N tools/check_bands_nccl.py : one image split into row bands over N GPUs
(NCCL) against the whole-image render on every rank; forward and backward."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import fields, gscuda, sharding

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for name in ("C1", "C2"):
    _, s, c, k, h, w = fields.make(name, 0)
    s, c, k = s.to(dev), c.to(dev), k.to(dev)
    img = sharding.render_image_bands(s, c, k, h, w, 0.1)
    full = torch.zeros(h, w, 3, device=dev)
    gscuda.gs_render(s, c, k, full, s.shape[0], h, w, 3, 0.1)
    g = torch.rand(h, w, 3, device=dev, generator=torch.Generator(dev).manual_seed(1))
    got = sharding.backward_image_bands(s, c, k, g, h, w, 0.1)
    want = [torch.zeros_like(t) for t in (s, c, k)]
    gscuda.gs_render_backward(s, c, k, g, *want, s.shape[0], h, w, 3, 0.1)
    peer = sharding.render_image_bands_peer(s, c, k, h, w, 0.1, gather_to=0)
    if rank == 0:
        assert float((peer - full).abs().max()) <= 2e-6, "peer-written bands differ"
    torch.cuda.synchronize()
    peer8 = sharding.render_image_bands_peer(s, c, k, h, w, 0.1, gather_to=0, u8=True)
    if rank == 0:  # the uint8 image of inference_paper.py:136-138, from every rank's band kernel
        want8 = (full.clamp(0, 1) * 255.0).round().to(torch.int16)
        d8 = (peer8.to(torch.int16) - want8).abs()
        assert int(d8.max()) <= 1 and float((d8 > 0).float().mean()) < 1e-3, "peer-written uint8 bands differ"
    err = float((img - full).abs().max())
    rel = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(got, want))
    print(f"[rank {rank}/{dist.get_world_size()}] {name} {h}x{w}: bands vs whole image max-abs {err:.2e}, "
          f"gradients max rel {rel:.2e}", flush=True)
    assert err <= 2e-6 and rel <= 1e-4, (err, rel)  # (gradients: fp32 summation order, atomics both ways)
dist.destroy_process_group()
