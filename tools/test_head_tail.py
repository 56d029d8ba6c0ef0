"""Stage-2 check of the fused head tail against torch (run on the GPU box)."""
import ctypes, os, sys, time, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import head_tail, _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
def blocks(c):
    ks = (2, 1, 1, 3, 2)
    return [nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, 4 * c), nn.ReLU(), nn.Linear(4 * c, k)).to(dev) for k in ks]
def reference(query, blks, emulate):
    b, gh, gw, c = query.shape
    outs = []
    for blk in blks:
        if emulate:  # the kernel's arithmetic: bf16 operands, fp32 accumulation, bf16 first hidden layer, fp32 tail
            x = query.to(torch.bfloat16).float()
            h1 = torch.relu(x @ blk[0].weight.to(torch.bfloat16).float().t() + blk[0].bias).to(torch.bfloat16).float()
            h2 = torch.relu(h1 @ blk[2].weight.to(torch.bfloat16).float().t() + blk[2].bias)
            outs.append(h2 @ blk[4].weight.t() + blk[4].bias)
        else:
            outs.append(blk(query))
    sig, rho, al, rgb, mean = [o.reshape(b, -1, o.shape[-1]) for o in outs]
    mean = mean / torch.tensor([gw, gh], device=dev)[None, None]
    sy, sx = 1 / gh, 1 / gw
    ry, rx = torch.meshgrid(torch.linspace(sy / 2, 1 - sy / 2, gh, dtype=torch.float32, device=dev),
                            torch.linspace(sx / 2, 1 - sx / 2, gw, dtype=torch.float32, device=dev), indexing="ij")
    mean = mean + torch.stack((rx.reshape(-1), ry.reshape(-1)), -1)[None]
    return torch.cat([sig, rho, al, rgb, mean], -1)
for c, b, gh, gw in ((192, 1, 16, 8), (192, 2, 48, 40), (180, 1, 24, 24), (192, 1, 64, 75)):
    blks = blocks(c)
    q = torch.randn(b, gh, gw, c, device=dev)
    pk = head_tail.PackedHeadTail(blks, dev)
    out = head_tail.fused_head_tail(q, pk)
    torch.cuda.synchronize()
    with torch.no_grad():
        r_em, r_32 = reference(q, blks, True), reference(q, blks, False)
    e1 = float((out - r_em).abs().max()); e2 = float((out - r_32).abs().max())
    print(f"C={c} b={b} grid {gh}x{gw}: vs emulated {e1:.3e}  vs fp32 {e2:.3e}  (|ref| max {float(r_32.abs().max()):.2f}) nan={int(torch.isnan(out).sum())}", flush=True)
# timing at the headline size: 2,097,152 rows
c, gh, gw = 192, 1024, 2048
blks = blocks(c); pk = head_tail.PackedHeadTail(blks, dev)
q = torch.randn(1, gh, gw, c, device=dev, dtype=torch.bfloat16)
for _ in range(2): out = head_tail.fused_head_tail(q, pk)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): out = head_tail.fused_head_tail(q, pk)
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 3
flop = 2.0 * gh * gw * 5 * (c * c + c * 4 * c + 4 * c * 9 / 5)
print(f"fused head tail, {gh*gw} rows: {ms:.2f} ms  ({flop / ms / 1e9:.0f} TFLOP/s)", flush=True)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    qs = q[:, :256]  # a quarter of the rows: the unfused reference materialises (rows x 768) per head
    for _ in range(2): [blk(qs) for blk in blks]
    torch.cuda.synchronize(); a.record()
    for _ in range(3): [blk(qs) for blk in blks]
    e.record(); torch.cuda.synchronize()
print(f"torch (cuBLAS, bf16 autocast), same rows: {a.elapsed_time(e) / 3 * 4:.2f} ms (extrapolated from a quarter)", flush=True)
