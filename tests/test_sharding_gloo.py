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
