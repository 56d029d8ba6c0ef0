"""split_and_joint_image mirror: tile geometry + paste rules against fixtures produced by the
reference's own function (tests/golden/make_golden.py), single process and sharded over two gloo
ranks (CPU; the renderer is a stub -- the stitching/sharding logic is what is under test)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden
from gsasr_b200.split_and_joint_image import plan_tiles, split_and_joint_image

FIXTURES = ["stitch_x4.npz", "stitch_x2p5.npz", "stitch_x3p3.npz"]


def _run(name, **kw):
    """Runs the mirror on the fixture's inputs.  The stub renderer draws tile i from the same seeded
    stream the fixture generator used for its i-th call; the tile index is captured by spying on
    render_units_sharded (which also records which tiles THIS rank rendered)."""
    import gsasr_b200.split_and_joint_image as sj

    g = golden(name)
    seed, scale = int(g["seed"]), float(g["scale"])
    lq = torch.rand(1, 3, int(g["h_lq"]), int(g["w_lq"]), generator=torch.Generator().manual_seed(seed))
    plan = plan_tiles(lq.shape[2], lq.shape[3], scale, int(g["split"]), int(g["overlap"]))
    assert plan.n == int(g["n_tiles"])
    order, state = [], {"i": -1}
    real = sj.render_units_sharded

    def spy(n_units, render_unit, **kws):
        def unit(i):
            order.append(i)
            state["i"] = i
            return render_unit(i)

        return real(n_units, unit, **kws)

    def fake_render(sr_size, gs_parameters, **_):
        gen = torch.Generator().manual_seed(1000 * seed + state["i"])
        return torch.rand(3, int(sr_size[0]), int(sr_size[1]), generator=gen)

    sj.render_units_sharded = spy
    try:
        out = split_and_joint_image(lq, scale, int(g["split"]), int(g["overlap"]), lambda t: t,
                                    lambda f, sv: torch.zeros(1, 4, 9), torch.tensor([scale, scale]),
                                    crop_size=int(g["crop"]), render_fn=fake_render, **kw)
    finally:
        sj.render_units_sharded = real
    return out, g["out"], order


@pytest.mark.parametrize("name", FIXTURES)
def test_matches_reference_stitching(name):
    out, ref, order = _run(name)
    assert order == list(range(len(order)))
    assert out.shape == ref.shape
    assert np.array_equal(out.numpy(), ref)


def test_plan_rejects_bad_overlap():
    with pytest.raises(AssertionError):
        plan_tiles(40, 40, 4.0, 16, 8)
    with pytest.raises(AssertionError):
        plan_tiles(6, 40, 4.0, 16, 4)    # padding larger than the image


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out, ref, order = _run(name, gather_to=0)
        if rank == 0:
            ret[rank] = bool(out is not None and np.array_equal(out.numpy(), ref)) and len(order) < int(golden(name)["n_tiles"])
        else:
            ret[rank] = out is None and len(order) > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["stitch_x4.npz", "stitch_x2p5.npz"])
def test_sharded_over_two_ranks(name):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), name, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


# ---- ownership regions (direct / peer-write stitching): the paste order resolved in advance ------
@pytest.mark.parametrize("h_lq,w_lq,scale,split,overlap,crop", [
    (100, 130, 4.0, 48, 8, 4), (100, 130, 2.5, 48, 8, 2), (64, 64, 3.3, 40, 6, 3), (90, 50, 2.0, 40, 4, 2),
    (48, 48, 4.0, 48, 8, 4),
])
def test_tile_regions_reproduce_the_paste_order(h_lq, w_lq, scale, split, overlap, crop):
    """Pasting constant tiles (tile i filled with i+1) with the fixture-pinned stitch_tiles gives the
    owner of every canvas pixel; tile_regions must describe exactly those pixels, disjointly."""
    from gsasr_b200.split_and_joint_image import plan_tiles, stitch_tiles, tile_regions

    plan = plan_tiles(h_lq, w_lq, scale, split, overlap)
    tiles = [torch.full((1, 1, plan.split_sr, plan.split_sr), float(i + 1)) for i in range(plan.n)]
    owner = stitch_tiles(tiles, plan, 1, 1, scale, crop)[0, 0].numpy()
    got = np.zeros_like(owner)
    regs = tile_regions(plan, crop, scale == int(scale))
    for i, reg in enumerate(regs):
        assert len(reg) <= 8
        for y0, y1, x0, x1 in reg:
            assert np.all(got[y0:y1, x0:x1] == 0), "regions overlap"
            got[y0:y1, x0:x1] = i + 1
    assert np.array_equal(got, owner)
