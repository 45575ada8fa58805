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
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok.item() == 1 else "FAIL", failures, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1 else 1)


if __name__ == "__main__":
    main()
