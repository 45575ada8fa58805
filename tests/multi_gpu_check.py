"""Run under torchrun (one rank per GPU): the X-slab sharded operator with the NVLink peer-memory all-reduce and
with NCCL against the unsharded operator, and a sharded relaxOver whose state stays bit-identical on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from carcassonne_b200 import distributed as cd
    from carcassonne_b200.data import DeviceData
    from carcassonne_b200.operator import Stage3Operator
    from carcassonne_b200.utils import Multiplier, relaxOver

    failures = []
    for D, X, d in ((2, 37, 2), (4, 300, 2), (8, 256, 2), (3, 5, 2)):
        rng = np.random.default_rng(100 + D)
        def crand(*shape):
            return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        terms = [(crand(X + t, D, D, D, D), crand(X + t, D, D, D, D), None if t % 2 == 0 else crand(d, d))
                 for t in range(3)]
        v = crand(D, D, D, D, d)
        full = Stage3Operator(v.shape)
        for A, B, O in terms:
            full.add_term(DeviceData.fromArray(A), DeviceData.fromArray(B), O)
        ref = full(DeviceData.fromArray(v)).toArray()
        comm = cd.PeerComm(v.size)
        for mode in ("peer", "nccl"):
            op = Stage3Operator(v.shape)
            for A, B, O in cd.shard_terms(terms, rank, world):
                op.add_term(DeviceData.fromArray(np.ascontiguousarray(A)), DeviceData.fromArray(np.ascontiguousarray(B)), O)
            op.finalize()
            if mode == "peer":
                cd.shard_operator(op, comm)
            for rep in range(5):                      # several epochs: exercises the double buffering
                out = op(DeviceData.fromArray(v))._t
                if mode == "nccl":
                    dist.all_reduce(out)
            err = np.linalg.norm(out.cpu().numpy() - ref) / np.linalg.norm(ref)
            gathered = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            identical = all(torch.equal(gathered[0], g) for g in gathered)
            if err > 1e-12 or (mode == "peer" and not identical):
                failures.append((D, X, mode, float(err), identical))
            if rank == 0:
                print("D=%d X=%d %s: relerr %.2e identical_across_ranks=%s" % (D, X, mode, err, identical), flush=True)
        if comm.timed_out():
            failures.append((D, X, "timeout"))
        # sharded relaxOver: H sharded, no normalization; every rank must hold the same state bit for bit
        h_terms = [(A + 0, B + 0, None) for A, B, _ in terms[:1]]
        op = Stage3Operator(v.shape)
        for A, B, O in cd.shard_terms(h_terms, rank, world):
            op.add_term(DeviceData.fromArray(np.ascontiguousarray(A)), DeviceData.fromArray(np.ascontiguousarray(B)), O)
        op.finalize()
        cd.shard_operator(op, comm)
        mult = Multiplier((v.size, v.size), op, 10 ** 12, None, 10 ** 15)
        mult.device_operator = op
        res = relaxOver(DeviceData.fromArray(v), mult, None, maximum_number_of_multiplications=12)._t
        gathered = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(gathered, res)
        if not all(torch.equal(gathered[0], g) for g in gathered):
            failures.append((D, X, "relax state differs across ranks"))
        comm.close()
    failures += system_level_checks(rank, world)
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok.item() == 1 else "FAIL", failures, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1 else 1)


def _identical_on_all_ranks(t, world):
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t.contiguous())
    return all(torch.equal(gathered[0], g) for g in gathered)


def system_level_checks(rank, world):
    """System-level multi-GPU mode: every rank holds the same seeded ``System``; with
    ``shard_environment()`` each builds only its X slab of the stage-2 halves and ``minimizeExpectation`` /
    ``computeExpectation`` run on all GPUs.  Checked: matvec, <H>, <N>, dense matrices and the minimised energy against
    the unsharded system on the same GPU (summation order differs, so not bit-for-bit: <= 1e-12 / 1e-10), and bit-identical
    states on all ranks (the replicated Arnoldi iteration relies on it)."""
    from copy import copy
    from carcassonne_b200 import distributed as cd
    from carcassonne_b200 import synthetic, utils
    failures = []

    def rel(a, b):
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

    # (chi, D, absorptions before the check, force GMRES for N^-1)
    # (chi = 1 without absorptions: the slow factor of X has ONE entry, so every rank but the last holds an EMPTY slab)
    cases = [(2, 2, (0, 1, 2, 3), False), (3, 3, (0, 1), False), (1, 2, (0, 1, 2, 3), False), (4, 4, (0, 2), False),
             (2, 3, (0, 1, 2, 3), True), (5, 2, (1,), False), (1, 2, (), False), (1, 3, (0,), False)]
    for chi, D, walk, force_gmres in cases:
        base = synthetic.device_system(chi, D, J=0.7, seed=10 + chi + D)
        for direction in walk:
            base.contractTowards(direction)
        v = base.state_center_data
        # ---- unsharded reference on this GPU
        cd.unshard_environment()
        H0, N0 = base.formExpectationAndNormalizationMultipliers()
        hv0, nv0 = H0(v).toArray(), N0(v).toArray()
        e0, n0 = base.computeExpectationAndNormalization()
        Hm0, Nm0 = H0.formMatrix().toArray(), N0.formMatrix().toArray()
        sub0 = base.formNormalizationSubmatrix().toArray()
        one = copy(base)
        saved = utils.Multiplier.isCheaperToFormMatrix
        if force_gmres:
            utils.Multiplier.isCheaperToFormMatrix = lambda self, n: False
        try:
            one.minimizeExpectation()
            energy0 = one.computeExpectation()
            # ---- sharded over all ranks
            sharding = cd.shard_environment()
            H1, N1 = base.formExpectationAndNormalizationMultipliers()
            hv1, nv1 = H1(v), N1(v)
            e1, n1 = base.computeExpectationAndNormalization()
            Hm1, Nm1 = H1.formMatrix(), N1.formMatrix()
            sub1 = base.formNormalizationSubmatrix()
            many = copy(base)
            stats = {}
            many.minimizeExpectation(statistics=stats)
            energy1 = many.computeExpectation()
            state = many.state_center_data._t
            checks = {
                "Hv": rel(hv1.toArray(), hv0), "Nv": rel(nv1.toArray(), nv0),
                "<H>": abs(e1 - e0) / abs(e0), "<N>": abs(n1 - n0) / abs(n0),
                "H matrix": rel(Hm1.toArray(), Hm0), "N matrix": rel(Nm1.toArray(), Nm0),
                "N submatrix": rel(sub1.toArray(), sub0),
            }
            worst = max(checks.values())
            same = all(_identical_on_all_ranks(t, world) for t in (hv1._t, nv1._t, Hm1._t, Nm1._t, state))
            de = abs(energy1 - energy0) / abs(energy0)
            if worst > 1e-12 or not same or de > 1e-10 or sharding.world != world:
                failures.append(("system", chi, D, walk, checks, same, de))
            if rank == 0:
                print("system chi=%d D=%d walk=%s gmres=%s: worst rel err %.2e, minimised energy rel diff %.2e, "
                      "N^-1 by %s, identical_across_ranks=%s" % (chi, D, walk, force_gmres, worst, de,
                                                                 stats.get("normalization"), same), flush=True)
        finally:
            utils.Multiplier.isCheaperToFormMatrix = saved
            cd.unshard_environment()
    # a short sharded sweep: minimise + absorb + compress, replicated parts must stay in lock step
    np.random.seed(5)
    system = synthetic.device_system(2, 2, J=0.5, seed=3)
    cd.shard_environment()
    try:
        for direction in (0, 1, 2, 3, 0, 1):
            system.minimizeExpectation()
            system.contractTowards(direction)
            for corner_id in range(4):          # ConstantStateCompressionPolicy(2).apply()
                for d in (0, 1):
                    system.compressCornerStateTowards(corner_id, d, 2)
        energy = system.computeExpectation()
        if not _identical_on_all_ranks(system.state_center_data._t, world) or not np.isfinite(energy):
            failures.append(("sharded sweep diverged across ranks", energy))
        if rank == 0:
            print("sharded sweep of 6 iterations: <H> = %.12f, states identical across ranks" % energy.real, flush=True)
    finally:
        cd.unshard_environment()
    return failures


if __name__ == "__main__":
    main()
