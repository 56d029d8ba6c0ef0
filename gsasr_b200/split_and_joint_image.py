"""Mirror of utils/split_and_joint_image.py:98-232 (tiled inference, `--tile_process`), with the
tiles sharded over the ranks of a process group.

    split_and_joint_image(lq, scale_factor, split_size, overlap_size, model_g, model_fea2gs,
                          scale_modify, crop_size=2, default_step_size=1.2, mode='scale_modify',
                          cuda_rendering=True, if_dmax=False, dmax_mode='fix', dmax=25)

Same arguments, same (B, C, H_pad, W_pad) result.  The reference runs encoder -> head -> render
for every LR tile sequentially on one GPU (:127-151) and pastes the SR tiles in row-major order,
later tiles overwriting earlier ones except for the first `crop_size` rows/columns of every
non-first tile (:166-225).  Tiles are independent, so with torch.distributed initialised each rank
processes a contiguous block of tiles on its own GPU and the finished tiles are gathered to
`gather_to` (NCCL over NVLink) where the stitching rules are applied; without a process group this
is the reference's sequential loop on the B200 rasteriser.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import torch
import torch.nn.functional as F

from .sharding import render_units_sharded


@dataclass(frozen=True)
class TilePlan:
    tiles_h: int
    tiles_w: int
    pad_h: int
    pad_w: int
    split: int
    overlap: int
    split_sr: int
    overlap_sr: int

    @property
    def n(self):
        return self.tiles_h * self.tiles_w

    @property
    def stride(self):
        return self.split - self.overlap

    @property
    def sr_h(self):
        return (self.tiles_h - 1) * (self.split_sr - self.overlap_sr) + self.split_sr

    @property
    def sr_w(self):
        return (self.tiles_w - 1) * (self.split_sr - self.overlap_sr) + self.split_sr


def plan_tiles(h_lq: int, w_lq: int, scale_factor: float, split_size: int, overlap_size: int) -> TilePlan:
    """Tile geometry of :108-125,153-163."""
    assert overlap_size > 0 and overlap_size < split_size // 2, "overlap size is wrong"
    stride = split_size - overlap_size
    th = math.ceil((h_lq - overlap_size) / stride)
    tw = math.ceil((w_lq - overlap_size) / stride)
    pad_h = th * stride + overlap_size - h_lq
    pad_w = tw * stride + overlap_size - w_lq
    assert pad_h < h_lq, f"pad_h_lq-{pad_h} should be smaller than h_lq-{h_lq}, please decrease the split_size-{split_size}"
    assert pad_w < w_lq, f"pad_w_lq-{pad_w} should be smaller than w_lq-{w_lq}, please decrease the split_size-{split_size}"
    return TilePlan(th, tw, pad_h, pad_w, split_size, overlap_size, math.ceil(split_size * scale_factor),
                    math.ceil(overlap_size * scale_factor))


def _paste_rule(plan: TilePlan, hn: int, wn: int, crop: int, integer_scale: bool):
    """(rows to skip, columns to skip) at the top/left of tile (hn, wn) when it is pasted.
    Integer scales: every non-first row/column of tiles skips `crop` (:213-225).  Non-integer
    scales follow the reference's case table (:168-205), in which an interior tile in the LAST
    column (or LAST row) skips only its columns (or only its rows)."""
    if integer_scale:
        return (crop if hn else 0), (crop if wn else 0)
    last_h, last_w = hn == plan.tiles_h - 1, wn == plan.tiles_w - 1
    if hn == 0:
        return 0, (crop if wn else 0)
    if wn == 0:
        return crop, 0
    if last_w and not last_h:
        return 0, crop
    if last_h and not last_w:
        return crop, 0
    return crop, crop


def stitch_tiles(tiles: List[torch.Tensor], plan: TilePlan, batch: int, channels: int, scale_factor: float,
                 crop_size: int) -> torch.Tensor:
    """Paste the (1,C,S,S) SR tiles (row-major) into the padded SR canvas (:160-227)."""
    s, step = plan.split_sr, plan.split_sr - plan.overlap_sr
    for t in tiles:
        assert t.shape[-2] == s and t.shape[-1] == s, f"tile {tuple(t.shape)} is not {s}x{s}"
    canvas = torch.zeros(batch, channels, plan.sr_h, plan.sr_w, device=tiles[0].device)
    integer_scale = scale_factor == int(scale_factor)
    idx = 0
    for hn in range(plan.tiles_h):
        for wn in range(plan.tiles_w):
            top, left = _paste_rule(plan, hn, wn, crop_size, integer_scale)
            y0, x0 = hn * step, wn * step
            y1, x1 = min(y0 + s, plan.sr_h), min(x0 + s, plan.sr_w)
            canvas[:, :, y0 + top:y1, x0 + left:x1] = tiles[idx][:, :, top:y1 - y0, left:x1 - x0]
            idx += 1
    return canvas


def paste_rect(plan: TilePlan, i: int, crop: int, integer_scale: bool):
    """(y0, y1, x0, x1), half-open, canvas coordinates: the block tile i writes when it is pasted."""
    hn, wn = divmod(i, plan.tiles_w)
    top, left = _paste_rule(plan, hn, wn, crop, integer_scale)
    step = plan.split_sr - plan.overlap_sr
    y0, x0 = hn * step, wn * step
    return y0 + top, min(y0 + plan.split_sr, plan.sr_h), x0 + left, min(x0 + plan.split_sr, plan.sr_w)


def tile_regions(plan: TilePlan, crop: int, integer_scale: bool):
    """For every tile the disjoint rectangles (y0, y1, x0, x1; half-open, canvas coordinates) of the
    canvas pixels it OWNS: those whose final value comes from this tile under the reference's paste
    order (row-major, later tiles overwrite earlier ones, :160-227).  With the regions known in advance
    the tiles can be written into the canvas in any order -- concurrently, from different GPUs."""
    rects = [paste_rect(plan, i, crop, integer_scale) for i in range(plan.n)]
    ys = sorted({v for r in rects for v in r[:2]})
    xs = sorted({v for r in rects for v in r[2:]})
    owner = {}
    for a, (ya, yb) in enumerate(zip(ys[:-1], ys[1:])):
        for b, (xa, xb) in enumerate(zip(xs[:-1], xs[1:])):
            for i in range(plan.n - 1, -1, -1):  # the last tile covering the cell wins
                y0, y1, x0, x1 = rects[i]
                if y0 <= ya and yb <= y1 and x0 <= xa and xb <= x1:
                    owner[(a, b)] = i
                    break
    regions = [[] for _ in range(plan.n)]
    for a in range(len(ys) - 1):  # merge each tile's cells of a strip into runs, then equal runs of strips
        b = 0
        while b < len(xs) - 1:
            i = owner.get((a, b))
            if i is None:
                b += 1
                continue
            e = b
            while e + 1 < len(xs) - 1 and owner.get((a, e + 1)) == i:
                e += 1
            run = [ys[a], ys[a + 1], xs[b], xs[e + 1]]
            prev = regions[i][-1] if regions[i] else None
            if prev is not None and prev[1] == run[0] and prev[2:] == run[2:]:
                prev[1] = run[1]
            else:
                regions[i].append(run)
            b = e + 1
    return [[tuple(r) for r in reg] for reg in regions]


_SYMM_CACHE = {}


def _peer_canvas(numel: int, device, gather_to: int, group):
    """A symmetric-memory canvas: (this rank's buffer, a view of `gather_to`'s buffer, the handle).
    Allocation and rendezvous cost milliseconds, so they are cached per (size, device, group)."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem

    key = (numel, str(device), id(group))
    if key not in _SYMM_CACHE:
        buf = symm_mem.empty(numel, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(buf, group=group if group is not None else dist.group.WORLD)
        _SYMM_CACHE[key] = (buf, hdl)
    buf, hdl = _SYMM_CACHE[key]
    return buf, hdl.get_buffer(gather_to, (numel,), torch.float32), hdl


def split_and_joint_image(lq, scale_factor, split_size, overlap_size, model_g, model_fea2gs, scale_modify,
                          crop_size=2, default_step_size=1.2, mode='scale_modify', cuda_rendering=True,
                          if_dmax=False, dmax_mode='fix', dmax=25, *, render_fn=None, gather_to=0, group=None,
                          direct=False, fused=False):
    """Returns the stitched (B,C,H_pad,W_pad) SR image on rank `gather_to` (every rank when it is
    None or when torch.distributed is not initialised); None on the other ranks.

    direct=True (inference, B = 1): no tile buffers and no paste pass -- the pixels every tile owns under
    the reference's paste order are known in advance (tile_regions), so the raster kernel writes each tile
    straight into the canvas (gsr_forward_window); with a process group the canvas lives in symmetric
    memory on `gather_to` and the other ranks' kernels store into it over NVLink (peer writes), the
    transfer overlapping the raster pixel by pixel instead of following it as a gather.  fused=True (with
    direct) also replaces the elementwise torch kernels of the front end by the library's fused one."""
    h_lq, w_lq = lq.shape[-2:]
    plan = plan_tiles(h_lq, w_lq, scale_factor, split_size, overlap_size)
    lq_pad = F.pad(input=lq, pad=(0, plan.pad_w, 0, plan.pad_h), mode='reflect')

    def tile_parameters(i: int):
        hn, wn = divmod(i, plan.tiles_w)
        y, x = hn * plan.stride, wn * plan.stride
        tile = lq_pad[:, :, y:y + split_size, x:x + split_size]
        feat = model_g(tile)
        scale_vector = scale_modify[0].unsqueeze(0).to(feat.device)
        return model_fea2gs(feat, scale_vector)[0, :]

    if direct:
        return _split_and_joint_direct(lq, plan, tile_parameters, scale_factor, scale_modify, crop_size,
                                       default_step_size, mode, cuda_rendering, if_dmax, dmax_mode, dmax,
                                       gather_to, group, fused)
    if render_fn is None:
        from .gaussian_splatting import generate_2D_gaussian_splatting_step as render_fn

    def render_tile(i: int) -> torch.Tensor:
        out = render_fn(sr_size=torch.tensor([plan.split_sr, plan.split_sr]), gs_parameters=tile_parameters(i),
                        scale=scale_factor, sample_coords=None, scale_modify=scale_modify,
                        default_step_size=default_step_size, mode=mode, cuda_rendering=cuda_rendering,
                        if_dmax=if_dmax, dmax_mode=dmax_mode, dmax=dmax)
        return out.unsqueeze(0)

    tiles = render_units_sharded(plan.n, render_tile, gather_to=gather_to, group=group)
    if tiles is None:
        return None
    return stitch_tiles(tiles, plan, lq.shape[0], lq.shape[1], scale_factor, crop_size)


def split_and_joint_image_buffer(lq, scale_factor, split_size, overlap_size, model_g, model_fea2gs, scale_modify,
                                 crop_size=2, default_step_size=1.2, mode='scale_modify', cuda_rendering=True,
                                 if_dmax=False, dmax_mode='fix', dmax=25, buffer_size=4000000, *, gather_to=0,
                                 group=None):
    """TrainTestGSASR/basicsr/utils/split_and_joint_image.py:142-285: split_and_joint_image whose tiles are
    rendered by generate_2D_gaussian_splatting_step_buffer -- the Gaussians of a tile are handed to the
    rasteriser in slices of `buffer_size`, accumulated into the same image.  Same tiling, paste rules and
    multi-GPU tile sharding as split_and_joint_image."""
    from .gaussian_splatting import generate_2D_gaussian_splatting_step_buffer

    def render_fn(**kw):
        return generate_2D_gaussian_splatting_step_buffer(buffer_size=buffer_size, **kw)

    return split_and_joint_image(lq, scale_factor, split_size, overlap_size, model_g, model_fea2gs, scale_modify,
                                 crop_size, default_step_size, mode, cuda_rendering, if_dmax, dmax_mode, dmax,
                                 render_fn=render_fn, gather_to=gather_to, group=group)


def _split_and_joint_direct(lq, plan, tile_parameters, scale_factor, scale_modify, crop_size,
                            default_step_size, mode, cuda_rendering, if_dmax, dmax_mode, dmax, gather_to, group,
                            fused=False):
    import torch.distributed as dist

    from .gaussian_splatting import render_into_canvas
    from .sharding import shard_range

    if not cuda_rendering:
        raise RuntimeError("direct=True writes tiles from the CUDA raster kernel; use direct=False with cuda_rendering=False")
    if lq.shape[0] != 1 or lq.shape[1] != 3:
        raise RuntimeError("direct=True renders one RGB image (B = 1, C = 3), like the reference's tile loop")
    regions = tile_regions(plan, crop_size, scale_factor == int(scale_factor))
    step = plan.split_sr - plan.overlap_sr
    numel = 3 * plan.sr_h * plan.sr_w
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    if multi:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        root = 0 if gather_to is None else gather_to
        mine, target, hdl = _peer_canvas(numel, lq.device, root, group)
        if rank == root:
            mine.zero_()
        torch.cuda.current_stream().synchronize()
        hdl.barrier()  # the canvas is cleared before anybody stores into it
        lo, hi = shard_range(plan.n, rank, world)
    else:
        rank = root = 0
        mine = target = torch.zeros(numel, dtype=torch.float32, device=lq.device)
        lo, hi = 0, plan.n
    canvas = target.view(1, 3, plan.sr_h, plan.sr_w)
    for i in range(lo, hi):
        hn, wn = divmod(i, plan.tiles_w)
        render_into_canvas(canvas, hn * step, wn * step, regions[i],
                           torch.tensor([plan.split_sr, plan.split_sr]), tile_parameters(i), scale_factor,
                           scale_modify, default_step_size, mode, if_dmax, dmax_mode, dmax, fused)
    if not multi:
        return canvas
    torch.cuda.current_stream().synchronize()
    hdl.barrier()  # every rank's stores have landed
    out = mine.view(1, 3, plan.sr_h, plan.sr_w).clone() if rank == root else None
    if gather_to is None:  # every rank wants the image: one broadcast from the stitching rank
        if out is None:
            out = torch.empty(1, 3, plan.sr_h, plan.sr_w, dtype=torch.float32, device=lq.device)
        dist.broadcast(out, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
    return out
