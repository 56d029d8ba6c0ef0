"""Head-tail fusion: the five per-Gaussian MLPs of the fea2gs head on the B200's tcgen05 tensor cores.

The reference's head ends with (utils/fea2gs.py:611-633, identical in utils/fea2gsropeamp.py:701-719)

    query_sigma = mlp_block_sigma(query); query_rho = mlp_block_rho(query); ... (five Linear-ReLU-Linear-ReLU-Linear
    stacks C -> C -> 4C -> k on the (b, H, W, C) feature map), query_mean / grid size + reference points, torch.cat

``fused_head_tail`` computes the same (b, H*W, 9) raw parameter tensor in ONE kernel (gsr_head_tail_forward): bf16
operands, fp32 accumulation in tensor memory, the first hidden layer kept in shared memory and the 4C-wide second
hidden layer consumed from tensor memory by the last Linear -- neither ever reaches HBM.  Inference only (no backward).
Arithmetic = the reference under bf16 autocast with the last Linear in fp32.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

CP, HP, HEADS = 192, 768, 5
_ORDER = ("mlp_block_sigma", "mlp_block_rho", "mlp_block_alpha", "mlp_block_rgb", "mlp_block_mean")
_KOUT = (2, 1, 1, 3, 2)


class PackedHeadTail:
    """Weights of the five MLPs, padded to (CP, HP) and laid out for the kernel.  Build once per checkpoint."""

    def __init__(self, blocks, device):
        # blocks: five nn.Sequential(Linear(C,C), ReLU, Linear(C,4C), ReLU, Linear(4C,k)) in _ORDER
        c = blocks[0][0].in_features
        if c > CP or blocks[0][2].out_features > HP:
            raise RuntimeError(f"channel {c} exceeds the kernel's padded size {CP}")
        self.channel = c
        w1 = torch.zeros(HEADS * CP, CP, dtype=torch.float32)
        b1 = torch.zeros(HEADS, CP, dtype=torch.float32)
        w2 = torch.zeros(HEADS * HP, CP, dtype=torch.float32)
        b2 = torch.zeros(HEADS, HP, dtype=torch.float32)
        w3 = torch.zeros(9, HP, dtype=torch.float32)
        b3 = torch.zeros(9, dtype=torch.float32)
        off = 0
        for h, (blk, k) in enumerate(zip(blocks, _KOUT)):
            l1, l2, l3 = blk[0], blk[2], blk[4]
            if l3.out_features != k:
                raise RuntimeError("gs_up_factor must be 1 (outputs 2, 1, 1, 3, 2)")
            h4 = l2.out_features
            w1[h * CP:h * CP + c, :c] = l1.weight.detach().float().cpu()
            b1[h, :c] = l1.bias.detach().float().cpu()
            w2[h * HP:h * HP + h4, :c] = l2.weight.detach().float().cpu()
            b2[h, :h4] = l2.bias.detach().float().cpu()
            w3[off:off + k, :h4] = l3.weight.detach().float().cpu()
            b3[off:off + k] = l3.bias.detach().float().cpu()
            off += k
        self.w1 = w1.to(device=device, dtype=torch.bfloat16).contiguous()
        self.w2 = w2.to(device=device, dtype=torch.bfloat16).contiguous()
        self.b1, self.b2, self.w3, self.b3 = (t.to(device).contiguous() for t in (b1, b2, w3, b3))

    @classmethod
    def from_module(cls, head, device=None):
        """From a reference Fea2GS / Fea2GS_ROPE_AMP instance (its mlp_block_* attributes)."""
        blocks = [getattr(head, n) for n in _ORDER]
        return cls(blocks, device or next(head.parameters()).device)


def fused_head_tail(query: torch.Tensor, packed: PackedHeadTail) -> torch.Tensor:
    """query: (b, H, W, C) feature map after UPNet + permute (fea2gs.py:606-607) -> (b, H*W, 9) raw Gaussian parameters
    (sigma_x, sigma_y, rho, alpha, r, g, b, mean_x, mean_y), mean already normalised and offset by the reference points."""
    if query.dim() != 4 or query.shape[-1] != packed.channel or not query.is_cuda:
        raise RuntimeError("query must be a CUDA tensor of shape (b, H, W, C)")
    L = _lib.load()
    b, gh, gw, c = query.shape
    m = b * gh * gw
    x = torch.zeros(m, CP, dtype=torch.bfloat16, device=query.device) if c < CP else None
    if x is None:
        x = query.reshape(m, c).to(torch.bfloat16).contiguous()
    else:
        x[:, :c] = query.reshape(m, c)
    raw = torch.empty(m, 9, dtype=torch.float32, device=query.device)
    with torch.cuda.device(query.device):
        rc = L.gsr_head_tail_forward(x.data_ptr(), packed.w1.data_ptr(), packed.b1.data_ptr(), packed.w2.data_ptr(),
                                     packed.b2.data_ptr(), packed.w3.data_ptr(), packed.b3.data_ptr(), raw.data_ptr(),
                                     m, gh, gw, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    return raw.view(b, gh * gw, 9)


class _CaptureTail(torch.nn.Module):
    """Stands in for one mlp_block_* while the reference head's own forward runs: remembers the feature map the tail is
    applied to and returns zeros of the right shape (the reference's concatenation of them is discarded)."""

    def __init__(self, k, store):
        super().__init__()
        self.k, self.store = k, store

    def forward(self, query):
        if not self.store:
            self.store.append(query)
        return query.new_zeros(*query.shape[:-1], self.k)


def forward_fused_tail(head, srcs, scale, packed: PackedHeadTail | None = None) -> torch.Tensor:
    """``head(srcs, scale)`` of a reference Fea2GS / Fea2GS_ROPE_AMP instance (utils/fea2gs.py:565-633,
    utils/fea2gsropeamp.py:655-719) with its tail -- the five MLPs, the mean normalisation, the reference points, the
    concatenation -- computed by the fused tensor-core kernel.  The body (embeddings, window cross-attention, Gaussian
    self-attention, UPNet) is the module's OWN forward, unmodified: the five ``mlp_block_*`` are swapped for capture
    stubs while it runs.  Inference only (no gradient flows through the fused tail)."""
    if packed is None:
        packed = getattr(head, "_gsr_packed_tail", None)
        if packed is None:
            packed = PackedHeadTail.from_module(head)
            head._gsr_packed_tail = packed  # (weights are read once: rebuild after loading another checkpoint)
    store, saved = [], {}
    try:
        for name, k in zip(_ORDER, _KOUT):
            saved[name] = head._modules[name]
            head._modules[name] = _CaptureTail(k, store)
        with torch.no_grad():
            head(srcs, scale)
    finally:
        for name, mod in saved.items():
            head._modules[name] = mod
    return fused_head_tail(store[0], packed)
