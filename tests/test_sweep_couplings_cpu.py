"""Host logic of the coupling-sweep launcher (scripts/sweep_couplings.py, BASELINE north_star: "independent
coupling-sweep points additionally run one per GPU"): the work split and the merge, single-process and at world size 2
over gloo (the transport the launcher itself uses for the results).  The per-point device runs are covered by
tests/test_gpu_drivers.py."""
import os
import socket
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_assignments_partition_the_grid():
    import sweep_couplings as sc
    for n in (0, 1, 5, 8, 33):
        for world in (1, 2, 3, 8):
            owned = [sc.assign_points(n, r, world) for r in range(world)]
            assert sorted(i for o in owned for i in o) == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    with pytest.raises(ValueError):
        sc.assign_points(4, 2, 2)


def test_merge_detects_missing_and_duplicate_points():
    import sweep_couplings as sc
    assert sc.merge_results([[(0, "a"), (2, "c")], [(1, "b")]], 3) == ["a", "b", "c"]
    with pytest.raises(ValueError):
        sc.merge_results([[(0, "a")], [(0, "b")]], 1)
    with pytest.raises(ValueError):
        sc.merge_results([[(0, "a")]], 2)


def test_grid_parsing():
    import argparse
    import sweep_couplings as sc
    assert sc.parse_grid(argparse.Namespace(J="0.1,0.5", J_grid=None)) == [0.1, 0.5]
    grid = sc.parse_grid(argparse.Namespace(J="", J_grid="0:2:5"))
    assert grid == [0.0, 0.5, 1.0, 1.5, 2.0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sweep_couplings as sc
        grid = [0.1 * i for i in range(7)]
        mine = [(i, {"J": grid[i], "rank": rank}) for i in sc.assign_points(len(grid), rank, world)]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((mine, float(rank)), gathered, dst=0)
        if rank == 0:
            points = sc.merge_results([g[0] for g in gathered], len(grid))
            q.put([(p["J"], p["rank"]) for p in points])
    finally:
        dist.destroy_process_group()


def test_world2_gloo_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    points = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r for _, r in points] == [0, 1, 0, 1, 0, 1, 0]
    assert [round(j, 1) for j, _ in points] == [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6]
