"""Multi-GPU sharding of the render path: one process per GPU, independent units, one gather.

The rasteriser has no exchange step: training-batch samples
(gsasr_model.py:191-233) and split_and_joint_image tiles (utils/split_and_joint_image.py:127-151)
are independent, so the units are dealt to the ranks in contiguous blocks, every rank renders its
block on its own GPU, and the only collective is the gather of the finished images to the rank
that stitches / computes the loss (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous block [lo, hi) of `n_units` owned by `rank` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(unit: int, n_units: int, world: int) -> int:
    base, rem = divmod(n_units, world)
    edge = rem * (base + 1)
    return unit // (base + 1) if unit < edge else rem + (unit - edge) // max(base, 1)


def render_units_sharded(n_units: int, render_unit: Callable[[int], torch.Tensor], *,
                         gather_to: int | None = 0, group=None) -> List[torch.Tensor] | None:
    """Render units [0, n_units) across the ranks of `group`.

    render_unit(i) -> image tensor of unit i (any shape, on this rank's device).
    Returns the full list of images on rank `gather_to` (every rank if gather_to is None),
    None elsewhere.  Shapes may differ between units (ragged tiles / samples): they are exchanged
    first, then the pixels, one broadcast per unit from its owner -- tiny next to the raster time
    (a 4096x4096x3 fp32 image is 201 MB, ~0.3 ms over NVLink)."""
    if not dist.is_initialized():
        return [render_unit(i) for i in range(n_units)]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_units, rank, world)
    mine = {i: render_unit(i) for i in range(lo, hi)}
    device = next(iter(mine.values())).device if mine else torch.device("cpu")
    backend = dist.get_backend(group)
    if backend == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    # torch.distributed takes GLOBAL ranks for src / dst whatever the group: translate group-local ones
    glob = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    shapes = torch.zeros(n_units, 9, dtype=torch.int64, device=device)  # [ndim, up to 8 extents]
    for i, t in mine.items():
        shapes[i, 0] = t.dim()
        shapes[i, 1:1 + t.dim()] = torch.tensor(list(t.shape), dtype=torch.int64)
    dist.all_reduce(shapes, op=dist.ReduceOp.SUM, group=group)
    out: List[torch.Tensor] = []
    for i in range(n_units):
        src = owner_of(i, n_units, world)
        nd = int(shapes[i, 0])
        shape = [int(v) for v in shapes[i, 1:1 + nd]]
        numel = 1
        for v in shape:
            numel *= v
        if numel == 0:  # an empty unit (e.g. a rank without rows): nothing to exchange
            if gather_to is None or rank == gather_to:
                out.append(mine[i] if src == rank else torch.empty(shape, dtype=torch.float32, device=device))
            continue
        if gather_to is None:
            buf = mine[i].contiguous() if src == rank else torch.empty(shape, dtype=torch.float32, device=device)
            dist.broadcast(buf, src=glob(src), group=group)
            out.append(buf)
        else:
            if src == gather_to:
                if rank == gather_to:
                    out.append(mine[i])
            elif rank == src:
                dist.send(mine[i].contiguous(), dst=glob(gather_to), group=group)
            elif rank == gather_to:
                buf = torch.empty(shape, dtype=torch.float32, device=device)
                dist.recv(buf, src=glob(src), group=group)
                out.append(buf)
    if gather_to is None or rank == gather_to:
        return out
    return None


def render_batch_sharded(gs_parameters: Sequence[torch.Tensor], sr_sizes, scales, *, dmax=0.1,
                         render_fn=None, gather_to: int | None = 0, group=None):
    """Config-5 shaped helper: B raw head outputs (N_i,9) with per-sample (H_i,W_i) and scale,
    rendered by the ranks in blocks.  render_fn defaults to the library's front end."""
    if render_fn is None:
        from .gaussian_splatting import generate_2D_gaussian_splatting_step as render_fn

    def unit(i):
        sc = float(scales[i])
        return render_fn(sr_size=sr_sizes[i], gs_parameters=gs_parameters[i], scale=sc,
                         scale_modify=torch.tensor([sc, sc]), dmax=dmax)

    return render_units_sharded(len(gs_parameters), unit, gather_to=gather_to, group=group)


# ---- one large image over several GPUs: row bands -------------------------------
# Every rank holds all Gaussians (32 B each: 67 MB for the 2M-Gaussian headline field) and renders
# the rows of its band with gsr_forward_band; the bands are then gathered.  Bands start on multiples
# of 8 rows -- the rasteriser's region height -- so no region is split between two ranks.  Backward:
# every rank turns its band of dL/dimg into partial parameter gradients (the gradients are linear in
# the pixels), summed with one all-reduce of the (N,3)+(N,2)+(N,3) block.
BAND_ALIGN = 8


def band_rows(h: int, rank: int, world: int):
    """(row0, rows) of `rank`'s band of an h-row image; rows == 0 if there are more ranks than
    8-row strips.  Every non-empty band has at least 2 rows (the C ABI's minimum)."""
    strips = (h + BAND_ALIGN - 1) // BAND_ALIGN
    starts = [min(BAND_ALIGN * shard_range(strips, r, world)[0], h) for r in range(world)] + [h]
    # a 1-row tail (h % 8 == 1 and the last strip alone in its band) takes a strip from the band above
    last = max(r for r in range(world) if starts[r] < h)
    if h - starts[last] < 2 and last > 0:
        starts[last] -= BAND_ALIGN
        for r in range(last):
            starts[r] = min(starts[r], starts[last])
    return starts[rank], starts[rank + 1] - starts[rank]


def _band_forward_cuda(sigmas, coords, colors, h, w, row0, rows, dmax, out=None):
    """Renders rows [row0, row0 + rows) into `out` (a contiguous (rows,w,3) view, e.g. the band's rows of the
    full image: no band buffer, no copy) or into a fresh tensor; every pixel is written (GSR_FLAG_OVERWRITE)."""
    from . import gscuda

    band = out if out is not None else torch.empty(rows, w, 3, dtype=torch.float32, device=sigmas.device)
    gscuda.gs_render_band(sigmas, coords, colors, band, sigmas.shape[0], h, w, 3, row0, rows, dmax,
                          flags=1)  # GSR_FLAG_OVERWRITE
    return band


_SYMM_IMAGES = {}


def _peer_image(numel: int, device, root: int, group, slot: int = 0, dtype=torch.float32):
    """Symmetric-memory image buffer: (this rank's buffer, view of `root`'s buffer, handle); cached."""
    import torch.distributed._symmetric_memory as symm_mem

    key = (numel, str(device), id(group), slot, dtype)
    if key not in _SYMM_IMAGES:
        buf = symm_mem.empty(numel, dtype=dtype, device=device)
        _SYMM_IMAGES[key] = (buf, symm_mem.rendezvous(buf, group=group if group is not None else dist.group.WORLD))
    buf, hdl = _SYMM_IMAGES[key]
    return buf, hdl.get_buffer(root, (numel,), dtype), hdl


_PEER_CALLS = {}
_PEER_VIEWS = {}


def render_image_bands_peer(sigmas, coords, colors, h: int, w: int, dmax: float, *, gather_to: int = 0,
                            group=None, u8: bool = False, bgr: bool = False):
    """render_image_bands with the gather folded into the raster kernel: the (h,w,3) image lives in
    symmetric memory on `gather_to` and every rank's band kernel stores its rows straight into it over
    NVLink (peer writes) -- no collective, the transfer overlaps the raster.  Returns a view of the
    symmetric buffer on `gather_to`, None elsewhere.  CUDA + NCCL only.

    Lifetime of the returned view: calls alternate between TWO symmetric images, and every call ends with a
    stream-ordered barrier across the ranks.  The image of call k is therefore overwritten no earlier than call
    k+2, whose kernels run behind the barrier of call k+1 on every rank -- and that barrier completes only after
    everything the stitching rank queued on its current stream before it (the readers of image k).  Work queued on
    OTHER streams must be ordered by the caller; a caller that keeps the image longer clones it.

    u8=True: the inference pipeline's result instead -- the (h,w,3) uint8 image of inference_paper.py:136-138
    (clamp, x255, round; bgr=True: cv2's channel order), produced by the raster kernel's write-out
    (GSR_FLAG_U8): 3 bytes per pixel cross NVLink instead of 12."""
    from . import gscuda

    dtype = torch.uint8 if u8 else torch.float32
    flags = (0x1 | 0x4 | (0x8 if bgr else 0)) if u8 else (0x1 | 0x10)  # OVERWRITE | U8 [| BGR]  /  OVERWRITE | ROW_STORES
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        if not u8:
            return _band_forward_cuda(sigmas, coords, colors, h, w, 0, h, dmax)
        out = torch.empty(h, w, 3, dtype=torch.uint8, device=sigmas.device)
        gscuda.gs_render_band(sigmas, coords, colors, out, sigmas.shape[0], h, w, 3, 0, h, dmax, flags=flags)
        return out
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = (h, w, str(sigmas.device), id(group), gather_to, u8)
    slot = _PEER_CALLS.get(key, 0)
    _PEER_CALLS[key] = slot ^ 1
    if (key, slot) not in _PEER_VIEWS:  # peer views, the band slice and the workspace are set up once
        mine, target, hdl = _peer_image(h * w * 3, sigmas.device, gather_to, group, slot, dtype)
        row0, rows = band_rows(h, rank, world)
        band = target.view(h, w, 3)[row0:row0 + rows] if rows > 0 else None  # rows of the stitching rank's image
        ws = gscuda.workspace(sigmas.shape[0], max(rows, 2), w, sigmas.device)
        _PEER_VIEWS[(key, slot)] = (mine.view(h, w, 3), band, hdl, row0, rows, ws, sigmas.shape[0])
    mine, band, hdl, row0, rows, ws, n_ws = _PEER_VIEWS[(key, slot)]
    if n_ws < sigmas.shape[0]:
        ws = gscuda.workspace(sigmas.shape[0], max(rows, 2), w, sigmas.device)
        _PEER_VIEWS[(key, slot)] = (mine, band, hdl, row0, rows, ws, sigmas.shape[0])
    if rows > 0:
        gscuda.gs_render_band(sigmas, coords, colors, band, sigmas.shape[0], h, w, 3, row0, rows, dmax,
                              flags=flags,  # fp32: GSR_FLAG_ROW_STORES -- 16-byte packets over NVLink (uint8: always)
                              workspace_buf=ws)
    # Stream-ordered barrier (a kernel on the current stream that signals every peer and waits for all of them):
    # behind it, every rank's stores of this call have landed; no host synchronisation, the call returns at once.
    hdl.barrier()
    return mine if rank == gather_to else None


def render_image_bands(sigmas, coords, colors, h: int, w: int, dmax: float, *, render_band=None,
                       gather_to: int | None = None, group=None):
    """gaussiansplatting_render of ONE (h,w) image split into row bands over the ranks of `group`.
    Every rank passes the same (sigmas, coords, colors).  Returns the (h,w,3) image on `gather_to`
    (on every rank if None), None elsewhere.  render_band(sigmas, coords, colors, h, w, row0, rows,
    dmax) -> (rows,w,3) defaults to the CUDA band kernel.

    The band shapes follow from (h, world) alone, so nothing but pixels is exchanged: every rank
    renders its band and the bands -- contiguous row blocks of the (h,w,3) result -- are gathered
    with ONE in-place all-gather when they are equal-sized, else one broadcast / send per band."""
    direct = render_band is None  # the CUDA band kernel writes its rows straight into the full image
    render_band = render_band or _band_forward_cuda
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return render_band(sigmas, coords, colors, h, w, 0, h, dmax)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bands = [band_rows(h, r, world) for r in range(world)]
    row0, rows = bands[rank]
    if direct and gather_to is None:
        full = torch.empty(h, w, 3, dtype=torch.float32, device=sigmas.device)
        if rows > 0:
            _band_forward_cuda(sigmas, coords, colors, h, w, row0, rows, dmax, out=full[row0:row0 + rows])
        return _gather_bands(full, bands, rank, group)
    mine = render_band(sigmas, coords, colors, h, w, row0, rows, dmax) if rows > 0 else None
    glob = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    if gather_to is not None:  # point-to-point: only the stitching rank holds the image
        if rank != gather_to:
            if rows > 0:
                dist.send(mine.contiguous(), dst=glob(gather_to), group=group)
            return None
        full = torch.empty(h, w, 3, dtype=torch.float32, device=sigmas.device)
        for r, (r0, n) in enumerate(bands):
            if n == 0:
                continue
            if r == rank:
                full[r0:r0 + n].copy_(mine)
            else:
                dist.recv(full[r0:r0 + n], src=glob(r), group=group)
        return full
    full = torch.empty(h, w, 3, dtype=torch.float32, device=sigmas.device)
    if rows > 0:
        full[row0:row0 + rows].copy_(mine)
    return _gather_bands(full, bands, rank, group)


def _gather_bands(full, bands, rank, group):
    """All ranks end up with every band of `full` (each rank has filled its own): ONE in-place all-gather when the
    bands are equal-sized, else one broadcast per band."""
    glob = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    row0, rows = bands[rank]
    if all(n == bands[0][1] for _, n in bands):
        dist.all_gather_into_tensor(full, full[row0:row0 + rows], group=group)  # in place: band r at offset r
    else:
        for r, (r0, n) in enumerate(bands):
            if n > 0:
                dist.broadcast(full[r0:r0 + n], src=glob(r), group=group)
    return full


def _grad_views(sigmas, coords, colors):
    """One flat fp32 buffer of 8 values per Gaussian and the three contiguous gradient arrays that live in it
    ((N,3) | (N,2) | (N,3), one after the other): a single collective reduces all of them, no packing copies."""
    n = sigmas.shape[0]
    flat = torch.zeros(8 * n, device=sigmas.device, dtype=torch.float32)
    return flat, flat[:3 * n].view(n, 3), flat[3 * n:5 * n].view(n, 2), flat[5 * n:].view(n, 3)


def _band_backward_cuda(sigmas, coords, colors, grads_band, h, w, row0, rows, dmax, out=None):
    from . import gscuda

    gs, gc, gk = out if out is not None else (torch.zeros_like(sigmas), torch.zeros_like(coords), torch.zeros_like(colors))
    gscuda.gs_render_backward_band(sigmas, coords, colors, grads_band.contiguous(), gs, gc, gk,
                                   sigmas.shape[0], h, w, 3, row0, rows, dmax)
    return gs, gc, gk


def backward_image_bands(sigmas, coords, colors, grads, h: int, w: int, dmax: float, *,
                         backward_band=None, group=None):
    """Gradients of render_image_bands: `grads` is the full (h,w,3) dL/dimg (every rank reads only
    its band of it).  Returns (grads_sigmas, grads_coords, grads_colors), identical on every rank
    after one all-reduce(sum) of 32 bytes per Gaussian."""
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    row0, rows = band_rows(h, rank, world)
    if backward_band is None and sigmas.is_cuda and sigmas.dtype == torch.float32:
        # the three gradient arrays are views of ONE buffer: a single all-reduce, no packing or unpacking copies
        flat, gs, gc, gk = _grad_views(sigmas, coords, colors)
        if rows > 0:
            _band_backward_cuda(sigmas, coords, colors, grads[row0:row0 + rows], h, w, row0, rows, dmax, out=(gs, gc, gk))
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return gs, gc, gk
    backward_band = backward_band or _band_backward_cuda
    if rows > 0:
        gs, gc, gk = backward_band(sigmas, coords, colors, grads[row0:row0 + rows], h, w, row0, rows, dmax)
    else:
        gs, gc, gk = torch.zeros_like(sigmas), torch.zeros_like(coords), torch.zeros_like(colors)
    if world > 1:
        packed = torch.cat([gs, gc, gk], dim=1).contiguous()  # (N,8): one collective instead of three
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        gs, gc, gk = packed[:, 0:3].contiguous(), packed[:, 3:5].contiguous(), packed[:, 5:8].contiguous()
    return gs, gc, gk
