"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch partition and the
one all-gather of the first-step policy that bench.py --gpus N performs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aligator_b200 import sharding


def test_shard_range_partitions_the_batch():
    for batch in (1, 7, 2048, 4096, 16384):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(batch, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _worker(rank, world, port, B, nu, nx, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    fb0 = torch.randn(B, nu + nx, nx, generator=g, dtype=torch.float64)
    ff0 = torch.randn(B, nu + nx, generator=g, dtype=torch.float64)
    pol = sharding.pack_first_step_policy(torch, fb0, ff0, nu, nx)
    allp = sharding.all_gather_policy(torch, dist, pol, world)
    ok = allp.shape == (world * B, nu, nx + 1)
    ok = ok and torch.equal(allp[rank * B:(rank + 1) * B], pol)
    # the other rank's block equals what that rank computed (same seeded generator)
    o = 1 - rank
    g2 = torch.Generator().manual_seed(100 + o)
    fbo = torch.randn(B, nu + nx, nx, generator=g2, dtype=torch.float64)
    ffo = torch.randn(B, nu + nx, generator=g2, dtype=torch.float64)
    ok = ok and torch.equal(allp[o * B:(o + 1) * B, :, :nx], fbo[:, :nu])
    ok = ok and torch.equal(allp[o * B:(o + 1) * B, :, nx], ffo[:, :nu])
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_policy_all_gather_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, 5, 3, 6, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _peer_setup_worker(rank, world, port, ret):
    """peer_gather_setup on a box where the peer buffers cannot be set up (here: no CUDA at all): every rank takes
    part in the same collectives and EVERY rank raises GarError -- nobody is left waiting in an all-gather."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import aligator_b200.gar as gar
    s = object.__new__(gar.CudaRiccatiBatch)  # no device: the handle cannot exist; the C entry point refuses a null one
    s.h = None
    try:
        s.peer_gather_setup(dist, rank, world)
        ret[rank] = "no error"
    except gar.GarError as e:
        ret[rank] = "GarError: " + str(e)
    dist.barrier()
    dist.destroy_process_group()


def test_peer_gather_setup_fails_on_all_ranks_together():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_setup_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0].startswith("GarError: peer gather unavailable") and ret[1].startswith("GarError: peer gather unavailable")
