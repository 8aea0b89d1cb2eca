"""world_size-2 gloo test (CPU) of the host-side N>1 logic: t-slab ownership, the halo ring of SURVEY.md 8e
(slot tloc <- next rank's first slice, slot tloc+1 <- previous rank's last slice) and the rank-ordered scalar
reduction, checked against the oracle on the assembled global lattice."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dims, q):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gf_oracle as oracle
    from gfb200.slabs import SlabDecomposition

    nx, ny, nz, nt = dims
    dec = SlabDecomposition(nt, world, rank)
    Ug = oracle.hot_start_philox(dims, 42)  # decomposition-independent: every rank can build the global field
    t0, t1 = dec.t_range()
    local = Ug[:, t0:t1].copy()
    # halo ring over gloo
    up, dn = dec.exchange_halo(local, dist)
    assert np.array_equal(up, Ug[:, (t1 % nt)]), "t+1 halo must be the next rank's first slice"
    assert np.array_equal(dn, Ug[:, (t0 - 1) % nt]), "t-1 halo must be the previous rank's last slice"
    # slab-local plaquette from local + halos equals this rank's share of the global sum
    ext = np.concatenate([dn[:, None], local, up[:, None]], axis=1)  # t = -1 .. tloc
    part = dec.local_plaquette_sum(ext)
    tot = dec.ordered_sum(part, dist)
    want = oracle.plaquette_sum(Ug, dims)
    ok = abs(tot - want) < 1e-10 * abs(want) + 1e-9
    q.put((rank, ok, tot, want))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_slab_halo_and_reduction():
    world, dims = 2, (4, 4, 2, 8)
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ok, tot, want in res:
        assert ok, (rank, tot, want)
    assert res[0][2] == res[1][2]  # every rank holds the same rank-ordered total


def test_slab_ranges():
    sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))
    from gfb200.slabs import SlabDecomposition

    for world in (1, 2, 4, 8):
        got = [SlabDecomposition(64, world, r).t_range() for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == 64
        assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
        assert SlabDecomposition(64, world, 0).neighbours() == ((world - 1) % world, 1 % world)
    with pytest.raises(ValueError):
        SlabDecomposition(10, 4, 0)
    with pytest.raises(ValueError):
        SlabDecomposition(4, 4, 0)  # a slab needs at least two time-slices
