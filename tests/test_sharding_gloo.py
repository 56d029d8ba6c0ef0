"""World-size-2 gloo test of the multi-GPU sharding logic (CPU; the render function is the
oracle here -- the sharding code itself never touches it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsasr_b200 import fields, sharding


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 9, 32):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                got += list(range(lo, hi))
                for u in range(lo, hi):
                    assert sharding.owner_of(u, n, world) == r
            assert got == list(range(n))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _unit_image(i):
    """Unit i: a small ragged render by the oracle."""
    from oracle import oracle

    h, w = 12 + 3 * i, 20 - i
    p = fields.raw_field(6, 6, seed=i)
    s, c, k = fields.map_field(p, h, w, 2.0)
    return torch.from_numpy(oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, 0.3).astype(np.float32))


def _worker(rank, world, port, n_units, gather_to, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def unit(i):
            calls.append(i)
            return _unit_image(i)

        out = sharding.render_units_sharded(n_units, unit, gather_to=gather_to)
        lo, hi = sharding.shard_range(n_units, rank, world)
        ok = calls == list(range(lo, hi))
        if out is not None:
            ok = ok and len(out) == n_units and all(torch.equal(o, _unit_image(i)) for i, o in enumerate(out))
        else:
            ok = ok and rank != gather_to
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units,gather_to", [(5, 0), (4, 1), (3, None), (1, 0)])
def test_units_sharded_over_two_ranks(n_units, gather_to):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_units, gather_to, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _subgroup_worker(rank, world, port, n_units, gather_to, ret):
    """World of 3, the render runs in the subgroup {1, 2}: group-local rank r is global rank r + 1, so a send /
    recv / broadcast that forgets the translation goes to the wrong peer (or hangs)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sub = dist.new_group([1, 2])  # every rank must take part in the creation
        ok = True
        if rank in (1, 2):
            out = sharding.render_units_sharded(n_units, _unit_image, gather_to=gather_to, group=sub)
            local = rank - 1
            if gather_to is None or local == gather_to:
                ok = out is not None and len(out) == n_units and all(
                    torch.equal(o, _unit_image(i)) for i, o in enumerate(out))
            else:
                ok = out is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units,gather_to", [(5, 0), (4, 1), (3, None)])
def test_units_sharded_in_a_subgroup_uses_global_ranks(n_units, gather_to):
    world = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_subgroup_worker, args=(world, _free_port(), n_units, gather_to, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True, 2: True}


# ---- one image split into row bands over the ranks (sharding.render_image_bands) ----------------
def test_band_rows_partition_the_image():
    for h in (2, 3, 8, 9, 16, 17, 25, 100, 2048):
        for world in (1, 2, 3, 4, 8):
            tot = 0
            for r in range(world):
                row0, rows = sharding.band_rows(h, r, world)
                assert rows == 0 or (row0 == tot and rows >= 2 and row0 % sharding.BAND_ALIGN == 0)
                tot += rows
            assert tot == h


def _band_field(h, w):
    p = fields.raw_field(10, 10, seed=3)
    s, c, k = fields.map_field(p, h, w, 2.0)
    return s, c, k


def _oracle_band(sigmas, coords, colors, h, w, row0, rows, dmax):
    """Stand-in for the CUDA band kernel in the CPU test: the oracle's whole image, cut."""
    from oracle import oracle

    img = oracle.forward(sigmas.numpy(), coords.numpy(), colors.numpy(), h, w, dmax)
    return torch.from_numpy(img[row0:row0 + rows].astype(np.float32))


def _oracle_band_backward(sigmas, coords, colors, grads_band, h, w, row0, rows, dmax):
    from oracle import oracle

    g = np.zeros((h, w, 3), np.float32)
    g[row0:row0 + rows] = grads_band.numpy()
    return tuple(torch.from_numpy(a.astype(np.float32)) for a in
                 oracle.backward(sigmas.numpy(), coords.numpy(), colors.numpy(), g, dmax))


def _band_worker(rank, world, port, h, w, gather_to, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle

        s, c, k = _band_field(h, w)
        img = sharding.render_image_bands(s, c, k, h, w, 0.3, render_band=_oracle_band, gather_to=gather_to)
        want = oracle.forward(s.numpy(), c.numpy(), k.numpy(), h, w, 0.3).astype(np.float32)
        ok = (img is None and rank != gather_to) if (gather_to is not None and rank != gather_to) else \
            (img is not None and tuple(img.shape) == (h, w, 3) and np.array_equal(img.numpy(), want))
        g = torch.rand(h, w, 3, generator=torch.Generator().manual_seed(7))
        got = sharding.backward_image_bands(s, c, k, g, h, w, 0.3, backward_band=_oracle_band_backward)
        ref = oracle.backward(s.numpy(), c.numpy(), k.numpy(), g.numpy(), 0.3)
        for a, b in zip(got, ref):
            ok = ok and np.abs(a.numpy() - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-12)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("h,w,gather_to", [(40, 24, None), (32, 16, None), (17, 12, 0), (9, 10, 1)])
def test_image_bands_over_two_ranks(h, w, gather_to):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_band_worker, args=(world, _free_port(), h, w, gather_to, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
