#!/usr/bin/env python
"""Coupling sweep: independent ground-state runs over a grid of couplings J, one ``System`` per GPU.

BASELINE.json north_star: "independent coupling-sweep points additionally run one per GPU".  The J-dependent runs of the
reference are its transverse-Ising chain tests (tests/test_simulator_2d_in_1d.py:36-48, J = 0.01) and the closed form
they are compared with is scripts/computeTIinfinite.py:6-12; here the coupling is the swept parameter.  Points need no
communication: under torchrun (one rank per GPU) grid point i runs on rank ``i mod world`` (``assign_points``), each
rank drives its own device ``System`` to convergence with the reference's policies, and rank 0 gathers the results
over a gloo group and prints ONE JSON line with every point (energy per site, closed form, error, sweeps, iterations,
seconds), the job's wall time (max over ranks) and the aggregate throughput in sweep iterations per second.

    python scripts/sweep_couplings.py --J 0.02,0.1,0.3,0.5,0.8,1.0,1.4,2.0
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep_couplings.py --J-grid 0.05:0.8:32

``--model plane`` sweeps the infinite square lattice instead (bounded full-2D runs, see _drivers.run_tfim_plane).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def assign_points(n_points, rank, world):
    """Indices of the grid points rank ``rank`` runs: round-robin, so neighbouring couplings (similar cost: the bond
    dimension a run needs grows towards the critical coupling) land on different GPUs.  The assignments of all ranks
    partition range(n_points).  Pure host logic."""
    if not 0 <= rank < world:
        raise ValueError("rank {} outside world of {}".format(rank, world))
    return list(range(rank, n_points, world))


def parse_grid(args):
    if args.J_grid:
        lo, hi, n = args.J_grid.split(":")
        lo, hi, n = float(lo), float(hi), int(n)
        return [lo + (hi - lo) * i / max(n - 1, 1) for i in range(n)]
    return [float(x) for x in args.J.split(",") if x]


def merge_results(per_rank, n_points):
    """[(index, record), ...] lists of all ranks -> records in grid order; every point exactly once."""
    out = [None] * n_points
    for records in per_rank:
        for index, record in records:
            if out[index] is not None:
                raise ValueError("grid point {} was run twice".format(index))
            out[index] = record
    missing = [i for i, r in enumerate(out) if r is None]
    if missing:
        raise ValueError("grid points {} were not run".format(missing))
    return out


def run_point(model, J, args):
    import _drivers
    if model == "chain":
        counts = {}
        energy, seconds, bond, sweeps = _drivers.run_tfim_chain(J, seed=args.seed, sweep_tol=args.sweep_tol,
                                                                 run_tol=args.run_tol, counts=counts, max_bond=args.max_bond,
                                                                 max_iterations_per_sweep=args.max_iterations_per_sweep)
        # _drivers.tfim_infinite_chain_energy keeps the reference script's convention (computeTIinfinite.py:6-12: its
        # argument is twice the XX coupling, lam = J / 2); the run's Hamiltonian is -sum Z - J sum X X
        exact = float(_drivers.tfim_infinite_chain_energy(2.0 * J))
        return {"J": J, "energy_per_site": energy, "exact": exact, "error": abs(energy - exact), "bond": bond,
                "sweeps": sweeps, "iterations": counts.get("iterations"), "seconds": seconds}
    energies, seconds, bond, sweeps, note = _drivers.run_tfim_plane(J, args.chi, args.max_bandwidth, seed=args.seed)
    return {"J": J, "energies_by_bond": energies, "bond": bond, "sweeps": sweeps, "seconds": seconds, "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--J", default="0.02,0.1,0.3,0.5,0.8,1.0,1.4,2.0", help="comma-separated couplings")
    ap.add_argument("--J-grid", default=None, help="lo:hi:n evenly spaced couplings (overrides --J)")
    ap.add_argument("--model", default="chain", choices=["chain", "plane"])
    ap.add_argument("--chi", type=int, default=2)
    ap.add_argument("--max-bandwidth", type=int, default=2)
    ap.add_argument("--sweep-tol", type=float, default=1e-5)
    ap.add_argument("--run-tol", type=float, default=1e-7)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-bond", type=int, default=12,
                    help="chain runs stop growing the state bond here (the critical point J = 1 never converges otherwise)")
    ap.add_argument("--max-iterations-per-sweep", type=int, default=200)
    args = ap.parse_args()
    grid = parse_grid(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")          # results are a few hundred bytes of Python objects: host transport
        group = dist.group.WORLD
    # warm-up outside the timed region: library load, kernel images, allocator pools
    run_point(args.model, grid[0], args)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mine = [(i, run_point(args.model, grid[i], args)) for i in assign_points(len(grid), rank, world)]
    torch.cuda.synchronize()
    seconds = time.perf_counter() - t0
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((mine, seconds), gathered, dst=0, group=group)
        if rank != 0:
            dist.destroy_process_group()
            return
    else:
        gathered = [(mine, seconds)]
    points = merge_results([g[0] for g in gathered], len(grid))
    wall = max(g[1] for g in gathered)
    iterations = sum(p.get("iterations") or 0 for p in points)
    line = {"what": "coupling sweep, one System per GPU, no communication", "model": args.model, "n_gpus": world,
            "points": points, "wall_seconds_max_over_ranks": wall, "seconds_per_rank": [g[1] for g in gathered],
            "sweep_iterations_total": iterations,
            "sweep_iterations_per_second": iterations / wall if wall > 0 else None,
            "points_per_second": len(points) / wall if wall > 0 else None,
            "worst_error_vs_closed_form": max((p["error"] for p in points if "error" in p), default=None)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
