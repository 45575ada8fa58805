#!/usr/bin/env python
"""Headline benchmark: the center-site expectation matvec (BASELINE.json metric "center-site matvec GFLOP/s").

(The term structure is read off the planner itself: see term_table_device / term_table_cpu.)

A step is one application of the expectation multiplier H = sum_t B_t . (A_t . (O_t v)) that
``System.formExpectationMultiplier()`` builds (reference tensors/_2d/sparse.py:100-161, dense.py:115-203) for the
2D transverse-field Ising Hamiltonian after one absorption round: T = 9 sparse terms over 6 + 6 stage-2
environment tensors, state bond D and boundary bond chi, d = 2.
Tensors are synthetic (random complex128) and stay resident in HBM, exactly as they do between the reference's
formExpectationStage2 and the Arnoldi iteration; the environment is far larger than L2 (12 x 16 X D^4 bytes).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--D 8] [--chi 16] [--impl reference]

N > 1 (under torchrun): the joined environment bond X = chi^4 is split into N slabs, one per GPU (strong
scaling: the same matvec, sharded); every rank holds the full vector and the partial results are summed by
one allreduce of the N-vector over NVLink.

FLOP accounting: 8 x the cmac count the reference's CostTracker assigns to the multiplier
(``Multiplier.cost_of_multiply``, data/cost_tracker.py:17-21) = 16 T X D^6 d -- the work the reference performs
for this call, whatever shortcuts the device path takes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "center_site_expectation_matvec_gflops"
UNIT = "GFLOP/s"

_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)


def _index_terms(terms, operator_of):
    """[(tag_0, tag_1, tag_center)] -> (n_half0, n_half1, [(a, b, site operator or None)]) with tensors numbered in
    order of first appearance."""
    a_index, b_index, out = {}, {}, []
    for x, y, z in terms:
        a = a_index.setdefault(x, len(a_index))
        b = b_index.setdefault(y, len(b_index))
        out.append((a, b, operator_of(z)))
    return len(a_index), len(b_index), out


def _operator_lists(model, J):
    if model == "tfim":      # Os=[-Z], OO_LRs=[(X,-J X)], OO_UDs=[(X,-J X)]
        return [-_Z], [(_X, -J * _X)], [(_X, -J * _X)]
    pairs = [(_X, _X), (_Y, _Y), (_Z, _Z)]   # nearest-neighbour Heisenberg
    return [], list(pairs), list(pairs)


def term_table_device(model, J=1.0):
    """The stage-3 term structure as the product's own planner produces it: a trivial system (all bonds 1) of the
    model is absorbed once in every direction and its expectation multiplier's term list is read off.  Only the
    STRUCTURE (which half-0 tensor meets which half-1 tensor under which site operator) is used; the benchmark's
    tensors are synthetic."""
    from carcassonne_b200.data import DeviceData
    from carcassonne_b200.sparse import Identity
    from carcassonne_b200.system import System
    Os, UDs, LRs = _operator_lists(model, J)
    dev = DeviceData.fromArray
    system = System.newTrivialWithSparseOperator([dev(o) for o in Os], [(dev(a), dev(b)) for a, b in UDs],
                                                 [(dev(a), dev(b)) for a, b in LRs])
    for direction in range(4):
        system.contractTowards(direction)
    H, _ = system.formExpectationAndNormalizationMultipliers()
    ops = system.operator_center_tensor
    return _index_terms(H.terms, lambda z: None if z == Identity() else ops[z].toArray())


def term_table_cpu(model, J=1.0):
    """The same structure from the oracle's planner (reference arm / cpu_baseline only)."""
    from oracle import tags
    from oracle.system import System
    Os, UDs, LRs = _operator_lists(model, J)
    system = System.new_trivial(tags.make_sparse_operator(Os, UDs, LRs))
    for direction in range(4):
        system.contract_towards(direction)
    H, _ = system.multipliers()
    return _index_terms(H.terms, lambda z: None if z == tags.I else system.operator_center[z])


def cost_of_multiply(terms, X, D, d):
    """cmac count of the reference's generated contractors (dense.py:115-128 / 146-160)."""
    P = D * D
    cost = 0
    for _, _, o in terms:
        if o is not None:
            cost += d * d * P * P
        cost += X * P * (P * d) * P + P * (P * d) * (P * X)
    return cost


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


# ------------------------------------------------------------------------------------------------------------
def cpu_matvec_sample(D, d, table, X_sample, repeats=1, seed=0):
    """Times the oracle's restatement of the reference matvec (two tensordots per term, NumPy -> BLAS zgemm)
    on an X slab of the workload.  Returns (GFLOP/s, seconds, threads)."""
    from oracle import dense
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    P = D * D

    def crand(*shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    na, nb, terms = table
    A = [crand(X_sample, D, D, D, D) for _ in range(na)]
    B = [crand(X_sample, D, D, D, D) for _ in range(nb)]
    v = crand(D, D, D, D, d)

    def matvec():
        out = np.zeros_like(v)
        for a, b, o in terms:
            out += dense.stage3_multiply_joined(A[a], B[b], v, o)
        return out

    matvec() if X_sample <= 64 else None  # tiny warm-up only when cheap
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        matvec()
        best = min(best, time.perf_counter() - t0)
    flops = 8.0 * cost_of_multiply(terms, X_sample, D, d)
    return flops / best / 1e9, best, threads


def pick_cpu_sample(D, X, nterms=9):
    """X slab that costs about 10-20 s of CPU work at ~10 GFLOP/s."""
    per_x = 8.0 * nterms * 2 * (D ** 6) * 2
    xs = int(max(1, min(X, 1.5e11 / per_x)))
    return xs


def run_reference(args, rank, world):
    if rank != 0:
        return
    D, chi, d = args.D, args.chi, 2
    X = chi ** 4
    table = term_table_cpu(args.model)
    terms = table[2]
    xs = pick_cpu_sample(D, X, len(terms))
    if args.steps * 1.0 > 6:
        xs = max(1, xs * 6 // args.steps)
    for _ in range(args.warmup):
        cpu_matvec_sample(D, d, table, max(1, xs // 8))
    vals, secs = [], []
    threads = 1
    for _ in range(args.steps):
        g, s, threads = cpu_matvec_sample(D, d, table, xs)
        vals.append(g)
        secs.append(s)
    value = float(np.mean(vals))
    sample = "T=%d %s terms, D=%d, X slab of %d of %d (chi=%d)" % (len(terms), args.model, D, xs, X, chi)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3 * (X / xs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128 (f64)",
        "data": "synthetic", "config": workload_config(args, X, len(terms)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, X, nterms):
    name = "TFIM" if args.model == "tfim" else "Heisenberg"
    return {"workload": "%s expectation matvec, T=%d terms, D=%d, chi=%d (X=%d), d=2" % (name, nterms, args.D,
                                                                                          args.chi, X),
            "D": args.D, "chi": args.chi, "terms": nterms, "hamiltonian": args.model,
            "l2": "inputs (stage-2 tensors of 16*X*D^4 B each) larger than L2",
            "sharding": "X slabs over ranks + one-shot all-reduce of the output vector (%s)" % (
                "NVLink peer memory, fused into the stage-3 partial-sum kernel" if getattr(args, "reduce", "peer") == "peer"
                else "NCCL")}


# ------------------------------------------------------------------------------------------------------------
def run_device(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import ctypes as C
    from carcassonne_b200 import _lib
    from carcassonne_b200.data import DeviceData
    from carcassonne_b200.operator import Stage3Operator

    D, chi, d = args.D, args.chi, 2
    X = chi ** 4
    x_lo, x_hi = X * rank // world, X * (rank + 1) // world
    Xl = x_hi - x_lo
    na, nb, terms = term_table_device(args.model)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)

    def rnd(*shape):
        t = torch.empty(shape, dtype=torch.complex128, device="cuda")
        torch.view_as_real(t).normal_(generator=gen)
        return t

    scale = 1.0 / (D * D * np.sqrt(X))
    A = [DeviceData(rnd(Xl, D, D, D, D).mul_(scale)) for _ in range(na)]
    B = [DeviceData(rnd(Xl, D, D, D, D).mul_(scale)) for _ in range(nb)]
    op = Stage3Operator((D, D, D, D, d))
    for a, b, o in terms:
        op.add_term(A[a], B[b], o)
    op.finalize()
    if args.path:
        op.set_path(args.path)
    kernel_name = {1: "stage3_kernel (fused A.v -> O -> B^T, DMMA.8x8x4)",
                   3: "stage3f_kernel (fused A.v -> O -> B^T, spin index folded into the tile columns, DMMA.8x8x4)",
                   2: "zgemm_kernel x 3 per term (unfused DMMA GEMMs)"}.get(op.path, "?")
    A_keep.extend([A, B, op])
    comm = None
    if world > 1 and args.reduce == "peer":
        from carcassonne_b200 import distributed as cd
        comm = cd.PeerComm(D ** 4 * d)
        cd.shard_operator(op, comm)
    use_nccl = world > 1 and comm is None
    n = D ** 4 * d
    v_host = torch.empty((D, D, D, D, d), dtype=torch.complex128).pin_memory()
    torch.view_as_real(v_host).normal_()
    out_host = torch.empty_like(v_host).pin_memory()
    v = v_host.cuda()
    out = torch.empty_like(v)
    if world > 1:
        dist.broadcast(v, 0)

    def step():
        op.apply_raw(v, out)
        if use_nccl:
            dist.all_reduce(out)

    def step_e2e():
        v.copy_(v_host, non_blocking=True)
        op.apply_raw(v, out)
        if use_nccl:
            dist.all_reduce(out)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t0 = time.time()
    total_ms = timed(step, args.steps)
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # duration of the dominant kernel (fused stage-3 + its partial-sum pass; with the peer communicator attached the
    # partial-sum pass is the cross-GPU one, so at N > 1 this includes the exchange)
    kern_ms = timed(lambda: op.apply_raw(v, out), args.steps) / args.steps

    for _ in range(3):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    flops = 8.0 * cost_of_multiply(terms, X, D, d)            # whole job, reference accounting
    flops_local = 8.0 * cost_of_multiply(terms, Xl, D, d)
    ms_per_step = total_ms / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e9

    comm_failed = comm.timed_out() if comm is not None else False
    if comm is not None:
        comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if comm_failed:
        raise SystemExit("peer all-reduce timed out")

    tf = C.c_double()
    _lib.check(_lib.lib.carc_dmma_peak(4000, C.byref(tf), None))
    executed = op.executed_flops                              # DMMA flops the kernel issues (shared-B grouping)
    achieved_tf = executed / (kern_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("D%d_chi%d_n%d" % (D, chi, world))
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s",
                "frac": achieved_tf / tf.value, "traffic": traffic,
                "kernel": kernel_name,
                "executed_flops_per_launch": executed, "reference_flops_per_launch": flops_local,
                "note": "achieved = FP64 flops the kernel issues / its duration.  The kernel decomposes the term list "
                        "into stars (terms sharing a half-1 tensor: first products summed before one second product; "
                        "terms sharing a half-0 tensor: one first product reused), so it issues %.0f%% of the flops "
                        "the reference performs for the same result -- `value` counts the reference's flops"
                        % (100.0 * executed / flops_local),
                "peak_distinct_operands": 31.4,
                "peak_note": "peak = DMMA.8x8x4 issue rate with register-resident operands; with a fresh A/B fragment "
                             "per instruction (what any GEMM inner loop needs) the same microbenchmark tops out at "
                             "31.4 TFLOP/s on this part (scripts/dmma_rate.py, 8 warps/SM)",
                "peak_source": "DMMA.8x8x4 issue-rate microbenchmark run in this process (carc_dmma_peak); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "algorithmic_bytes": (na + nb) * 16 * Xl * D ** 4 + 32 * n,
                "hbm_gbs": ((na + nb) * 16 * Xl * D ** 4 + 32 * n) / (kern_ms * 1e-3) / 1e9}

    cpu = None
    if world == 1 and not args.no_cpu:
        xs = pick_cpu_sample(D, X, len(terms))
        g, s, threads = cpu_matvec_sample(D, d, term_table_cpu(args.model), xs)
        cpu = {"value": g, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "T=%d %s terms, D=%d, X slab of %d of %d, %.1f s, oracle.dense.stage3_multiply_joined "
                         "(NumPy tensordot -> BLAS zgemm)" % (len(terms), args.model, D, xs, X, s)}

    sweep = None
    if world == 1 and not args.no_sweep:
        op.close()
        del A, B, op
        sweep = sweep_section(args.no_cpu)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": workload_config(args, X, len(terms)),
        "clocks": clocks,
        "e2e": {"value": flops / (e2e_ms / args.steps * 1e-3) / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 16 * n,
                "what": "Stage3Operator applied to a pinned host vector: H2D of v, matvec (+allreduce), D2H of H v; "
                        "the environment stays resident as it does behind the reference's Multiplier closure"},
        "gpu_launches": 2 * args.steps,   # stage3_kernel + (s3_reduce_kernel | xgpu_allreduce_kernel) per step
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sweep": sweep,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_section(no_cpu):
    """Second half of the BASELINE metric: seconds per sweep iteration vs bond dimension.  One iteration =
    minimizeExpectation + contractTowards + ConstantStateCompressionPolicy(chi) on a synthetic double-layer TFIM
    environment (scripts/sweep_bench.py); four iterations (one per direction) per size.  The CPU column is the
    oracle's restatement of the same calls on the host cores, at the size where it finishes in tens of seconds."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import sweep_bench
    del A_keep[:]                      # drop the matvec environment (51 GB); torch's allocator reuses the blocks
    rows = []
    sweep_bench.device_iterations(2, 2, False)   # warm-up
    for D, chi in ((3, 6), (4, 8), (6, 8), (8, 8)):
        if D <= 6:
            # first pass at a size builds its per-layout offset tables and grows the allocator pools (3x the steady-state
            # time at D = 3); a sweep runs hundreds of iterations at one size, so time the second pass
            sweep_bench.device_iterations(chi, D, False)
        r = sweep_bench.device_iterations(chi, D, False)
        rows.append({"D": D, "chi": chi, "gpu_s_per_iteration": r["per_iteration"], "minimize_s": r["minimize"] / 4,
                     "contract_s": r["contract"] / 4, "compress_s": r["compress"] / 4,
                     "matvecs_per_minimize": r["mults"]})
    out = {"unit": "s per sweep iteration (minimize + contract + 8 corner compressions)", "sizes": rows}
    if not no_cpu:
        c = sweep_bench.cpu_iterations(6, 3)
        out["cpu"] = {"D": 3, "chi": 6, "cpu_s_per_iteration": c["per_iteration"], "kind": "port (oracle.system.System)",
                      "speedup_at_same_size": c["per_iteration"] / rows[0]["gpu_s_per_iteration"]}
    return out


A_keep = []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--D", type=int, default=8)
    ap.add_argument("--chi", type=int, default=16)
    ap.add_argument("--model", default="tfim", choices=["tfim", "heisenberg"],
                    help="which Hamiltonian's sparse term structure to benchmark (T = 9 / 20 terms)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the seconds-per-sweep-iteration section")
    ap.add_argument("--path", type=int, default=0, choices=[0, 1, 2, 3],
                    help="device path of the matvec: 0 automatic, 1 fused kernel, 3 fused kernel with the folded "
                         "tiling, 2 unfused GEMMs (kernel comparisons; the default is what the library picks)")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"], help="N > 1: how partial outputs are summed")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus %d must be launched with torchrun (one rank per GPU)" % args.gpus)
    run_device(args, rank, world, local_rank)


if __name__ == "__main__":
    # the JSON line must be the only thing on stdout: libraries that write to fd 1 (the NCCL version banner) go to stderr
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, "w")
    main()
    sys.stdout.flush()
