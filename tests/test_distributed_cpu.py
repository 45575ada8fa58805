"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: slab planning, term sharding and the
sum-over-ranks plumbing.  The local matvec is the CPU oracle standing in for the device kernel; what is under test is
that X-slab partial results, summed over ranks, reproduce the full matvec."""
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
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from carcassonne_b200 import distributed as cd
        from oracle import dense
        assert cd.world() == world and cd.rank() == rank
        rng = np.random.default_rng(7)          # same tensors on every rank
        D, X, d = 2, 11, 2                      # X not divisible by the world size: ragged slabs
        terms = []
        for t in range(3):
            A = rng.standard_normal((X + t, D, D, D, D)) + 1j * rng.standard_normal((X + t, D, D, D, D))
            B = rng.standard_normal((X + t, D, D, D, D)) + 1j * rng.standard_normal((X + t, D, D, D, D))
            O = None if t != 1 else rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))
            terms.append((A, B, O))
        v = rng.standard_normal((D, D, D, D, d)) + 1j * rng.standard_normal((D, D, D, D, d))
        full = sum(dense.stage3_multiply_joined(A, B, v, O) for A, B, O in terms)
        mine = cd.shard_terms(terms, rank, world)
        part = np.zeros_like(v)
        for A, B, O in mine:
            part += dense.stage3_multiply_joined(np.ascontiguousarray(A), np.ascontiguousarray(B), v, O)
        t = torch.from_numpy(part)
        cd.allreduce_sum_(t)
        err = np.linalg.norm(t.numpy() - full) / np.linalg.norm(full)
        rows = sum(A.shape[0] for A, _, _ in mine)
        q.put((rank, float(err), rows))
    finally:
        dist.destroy_process_group()


def test_slab_bounds_tile_exactly():
    sys.path.insert(0, ROOT)
    from carcassonne_b200.distributed import slab_bounds
    for X in (0, 1, 5, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            edges = [slab_bounds(X, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == X
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_bounds(10, 2, 2)


def test_sharded_matvec_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    for _, err, _ in results:
        assert err < 1e-13
    assert sum(r[2] for r in results) == 11 + 12 + 13
