import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gsasr_b200 import head_tail
dev = torch.device("cuda:0"); torch.manual_seed(0)
c = 192
blks = [nn.Sequential(nn.Linear(c, c), nn.ReLU(), nn.Linear(c, 4 * c), nn.ReLU(), nn.Linear(4 * c, k)).to(dev) for k in (2, 1, 1, 3, 2)]
pk = head_tail.PackedHeadTail(blks, dev)
q = torch.randn(1, 148 * 4, 128, c, device=dev, dtype=torch.bfloat16)
for _ in range(3): out = head_tail.fused_head_tail(q, pk)
torch.cuda.synchronize()
