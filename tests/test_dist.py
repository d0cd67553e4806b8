"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo run of the shard / pack / all-gather plumbing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_collisionavoidance_b200 import dist as D


def test_shard_ranges_cover_the_batch():
    for B in (1, 7, 4096, 131072, 10):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(B, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, N, nx, nu, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = D.shard_range(B, rank, world)
    g = torch.Generator().manual_seed(0)
    x = torch.rand((B, N + 1, nx), generator=g, dtype=torch.float64)
    u = torch.rand((B, N, nu), generator=g, dtype=torch.float64)
    st = torch.rand((B, 12), generator=g, dtype=torch.float64)
    packed = D.pack_results(torch, x[lo:hi], u[lo:hi], st[lo:hi])
    full = D.all_gather_results(torch, dist, packed, B, world)
    X, U, S = D.unpack_results(full, N, nx, nu)
    ok = bool(torch.equal(X, x) and torch.equal(U, u) and torch.equal(S, st[:, :7]))
    q.put((rank, ok, tuple(full.shape)))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_all_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    for B in (10, 7):  # even and ragged shards
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 5, 6, 2, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok, _ in res), res
        assert all(shape == (B, D.packed_width(5, 6, 2)) for _, _, shape in res)
