"""Timeline of one CTA of the head-tail kernel (GSH_CFG_TRACE build): prints per head / chunk the hand-over times (us)."""
import ctypes, os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GSR_LIB_PATH"] = sys.argv[1]
from gsasr_b200 import head_tail, _lib
L = _lib.load(); dev = torch.device("cuda:0"); torch.manual_seed(0); c = 192
blks = [nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, 4 * c), nn.ReLU(), nn.Linear(4 * c, k)).to(dev) for k in (2, 1, 1, 3, 2)]
pk = head_tail.PackedHeadTail(blks, dev)
q = torch.randn(1, 148 * 8, 128, c, device=dev, dtype=torch.bfloat16)
tr = torch.zeros(320, dtype=torch.int64, device=dev)
L.gsr_head_tail_set_trace.argtypes = [ctypes.c_void_p]; L.gsr_head_tail_set_trace(tr.data_ptr())
for _ in range(3): out = head_tail.fused_head_tail(q, pk)
torch.cuda.synchronize()
t = tr.cpu().numpy().astype("int64"); t0 = t[0]
us = lambda v: (v - t0) / 1e3
for h in range(5):
    b = h * 64
    print(f"head {h}: L1 issue {us(t[b]):7.2f}  committed {us(t[b+1]):7.2f} | epi sees D1 {us(t[b+3]):7.2f}  H1 written {us(t[b+4]):7.2f} | MMA sees H1 {us(t[b+2]):7.2f}")
    for cc in range(6):
        print(f"    chunk {cc}: MMA sees D2 empty {us(t[b+8+4*cc]):7.2f}  committed {us(t[b+9+4*cc]):7.2f} | epi sees D2 full {us(t[b+10+4*cc]):7.2f}  done {us(t[b+11+4*cc]):7.2f}")
