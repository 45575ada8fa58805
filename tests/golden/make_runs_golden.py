#!/usr/bin/env python
"""Golden outputs of the reference's end-to-end simulator runs, produced by importing the UNMODIFIED reference from
/root/reference (build container only; writes tests/golden/runs_r2.json).

Every run is one of the reference's own integration tests with the RNG seeded first -- same constructor, same policies,
same thresholds:

  tests/test_simulator_2d_in_1d.py:14-61    magnetic field, ferromagnetic coupling, transverse Ising, Heisenberg (D -> 14)
  tests/test_policies_2d_in_1d.py:12-61     the four sweep-convergence policies on the transverse-Ising chain
  tests/test_simulator_2d_in_15d.py:11-50   the same models with ConstantStateCompressionPolicy(1), all four directions
  tests/test_simulator_1d.py:14-170         the 1D (MPS/MPO) system: field, transverse Ising, Haldane-Shastry, XY, Heisenberg

Recorded per run: the energy the reference's test asserts on, the sweep / iteration counters and the final center shape.
A run the reference itself cannot finish under the installed SciPy (`assert info == 0` after GMRES)
is recorded as {"reference_error": ...} -- the device test then only checks the reference test's own known answer.
"""
import json
import os
import random
import sys
import time

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("CARCASSONNE_REFERENCE", "/root/reference"))
HERE = os.path.dirname(os.path.abspath(__file__))

np.product = np.prod  # noqa: NPY003  (the reference's tests/__init__ uses the NumPy 1 name)

from carcassonne import policies as rpol  # noqa: E402
from carcassonne.data import NDArrayData as ND  # noqa: E402
from carcassonne.system import System  # noqa: E402
from carcassonne.system._1d import System as System1D  # noqa: E402
from carcassonne.utils import Pauli, buildProductTensor, buildTensor, crand  # noqa: E402


def seed(s):
    np.random.seed(s)
    random.seed(s)


def cplx(z):
    z = complex(z)
    return [z.real, z.imag]


def record(system, energy, t0):
    return {"energy": cplx(energy), "sweeps": int(system.number_of_sweeps),
            "iterations": int(system.number_of_iterations), "shape": [int(x) for x in system.state_center_data.shape],
            "seconds": time.time() - t0}


ONLY = None        # name of the single case this process runs (sub-process mode), or None to list the cases
CASES = []


def guarded(fn):
    """Runs the case only when it is the one this process was started for; always registers its name."""
    return fn


def axis_kw(direction, pair):
    return {"OO_LR" if direction == 0 else "OO_UD": pair}


def run_2d(system, sweep, run, increase, pattern, compression=None, estimated_direction=None):
    t0 = time.time()
    if compression is not None:
        system.setPolicy("state compression", compression)
    system.setPolicy("sweep convergence", sweep)
    system.setPolicy("run convergence", run)
    system.setPolicy("bandwidth increase", increase)
    system.setPolicy("contraction", rpol.RepeatPatternContractionPolicy(pattern))
    system.runUntilConverged()
    if estimated_direction is None:
        return record(system, system.computeOneSiteExpectation(), t0)
    return record(system, system.computeEstimatedOneSiteExpectation(estimated_direction), t0)


One = rpol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy
State = rpol.RelativeStateDifferenceThresholdConvergencePolicy
Estimated = rpol.RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy
Grow = rpol.OneDirectionIncrementBandwidthIncreasePolicy


def simulator_2d_in_1d(out):
    for d in (0, 1):
        seed(100 + d)
        out["2d_in_1d.magnetic_field.dir%d" % d] = guarded(lambda: run_2d(
            System.newTrivialWithSimpleSparseOperator(O=ND.Z), State(1e-5), One(1e-7), Grow(d), [0 + d, 2 + d]))
        seed(110 + d)
        out["2d_in_1d.ferromagnetic.dir%d" % d] = guarded(lambda: run_2d(
            System.newTrivialWithSimpleSparseOperator(**axis_kw(d, [ND.Z, -ND.Z])), One(1e-7), One(1e-7), Grow(d),
            [0 + d, 2 + d]))
        seed(120 + d)
        out["2d_in_1d.transverse_ising.dir%d" % d] = guarded(lambda: run_2d(
            System.newTrivialWithSimpleSparseOperator(O=-ND.Z, **axis_kw(d, [ND.X, -0.01 * ND.X])), State(1e-5),
            One(1e-7), Grow(d, 2), [0 + d, 2 + d]))
        seed(130 + d)
        pairs = [(ND.X, -ND.X), (ND.Y, -ND.Y), (ND.Z, ND.Z)]
        kw = {"OO_LRs" if d == 0 else "OO_UDs": pairs}
        out["2d_in_1d.heisenberg.dir%d" % d] = guarded(lambda: run_2d(
            System.newTrivialWithSparseOperator(**kw), Estimated(1e-5, d), Estimated(1e-4, d), Grow(d, 2),
            [d + 2, d + 0], estimated_direction=d))


def policies_2d_in_1d(out):
    def tfim(d):
        return System.newTrivialWithSimpleSparseOperator(O=-ND.Z, **axis_kw(d, [ND.X, -0.01 * ND.X]))
    for d in (0, 1):
        sweeps = {"one_site_expectation": lambda: One(1e-7), "estimated_one_site_expectation": lambda: Estimated(1e-7, d),
                  "state_difference": lambda: State(1e-5),
                  "periodicity": lambda: rpol.PeriodicyThresholdConvergencePolicy(1e-7, 0, 2)}
        for i, (name, make) in enumerate(sweeps.items()):
            seed(200 + 10 * i + d)
            out["policies.%s.dir%d" % (name, d)] = guarded(lambda: run_2d(tfim(d), make(), One(1e-7), Grow(d, 2),
                                                                           [0 + d, 2 + d]))


def simulator_2d_in_15d(out):
    compress = lambda: rpol.ConstantStateCompressionPolicy(1)  # noqa: E731
    seed(300)
    out["15d.magnetic_field"] = guarded(lambda: run_2d(
        System.newTrivialWithSimpleSparseOperator(O=ND.Z), One(1e-7), One(1e-7),
        rpol.AllDirectionsIncrementBandwidthIncreasePolicy(), range(4), compress()))
    for d in (0, 1):
        seed(310 + d)
        out["15d.ferromagnetic.dir%d" % d] = guarded(lambda: run_2d(
            System.newTrivialWithSimpleSparseOperator(**axis_kw(d, [ND.Z, -ND.Z])), One(1e-7), One(1e-7),
            rpol.AllDirectionsIncrementBandwidthIncreasePolicy(), range(4), compress()))
        seed(320 + d)
        out["15d.transverse_ising.dir%d" % d] = guarded(lambda: run_2d(
            System.newTrivialWithSimpleSparseOperator(O=-ND.Z, **axis_kw(d, [ND.X, -0.01 * ND.X])), One(1e-7), One(1e-7),
            Grow(d), range(4), compress()))


HS_A = [6.18505736e-04, 3.56927507e-01, 7.04055807e-05, 1.77859581e-02, 3.78493975e-03, 6.54917336e-02,
        1.83235170e-01, 4.09930918e-06, 3.72081681e-01]
HS_B = [0.97613415, 0.22719877, 0.99279374, 0.85346561, 0.93626109, 0.70473391, 0.48229086, 0.99858369, 0.0402559]


def run_1d(system, sweep, run, increment, estimated):
    t0 = time.time()
    system.setPolicy("sweep convergence", sweep)
    system.setPolicy("run convergence", run)
    system.setPolicy("bandwidth increase", Grow(0, increment))
    system.setPolicy("contraction", rpol.RepeatPatternContractionPolicy([0, 1]))
    system.runUntilConverged()
    energy = system.computeEstimatedOneSiteExpectation(0) if estimated else system.computeOneSiteExpectation()
    return record(system, energy, t0)


def simulator_1d(out):
    seed(400)
    out["1d.magnetic_field"] = guarded(lambda: run_1d(
        System1D(buildProductTensor([1, 0]), buildProductTensor([0, 1]),
                 buildTensor((2, 2, 2, 2), {(0, 0): Pauli.I, (1, 1): Pauli.I, (0, 1): -Pauli.Z}), np.ones((1, 1, 2))),
        State(1e-5), Estimated(1e-2), 1, True))
    seed(410)
    out["1d.transverse_ising"] = guarded(lambda: run_1d(
        System1D([1, 0, 0], [0, 0, 1],
                 buildTensor((3, 3, 2, 2), {(0, 0): Pauli.I, (0, 2): Pauli.Z, (0, 1): -0.01 * Pauli.X, (1, 2): Pauli.X,
                                            (2, 2): Pauli.I}), np.ones((1, 1, 2))),
        State(1e-5), One(1e-7), 2, False))
    seed(420)
    n = len(HS_A)
    matrix = {}
    l, r = 3 * n, 3 * n + 1
    matrix[l, l] = Pauli.I
    matrix[r, r] = Pauli.I
    for i in range(n):
        for k, P in enumerate((Pauli.X, Pauli.Y, Pauli.Z)):
            matrix[l, k * n + i] = HS_A[i] * P
            matrix[k * n + i, k * n + i] = HS_B[i] * Pauli.I
            matrix[k * n + i, r] = P
    initial = crand(1, 1, 2)
    out["1d.haldane_shastry"] = guarded(lambda: dict(run_1d(
        System1D([0] * (3 * n) + [1, 0], [0] * (3 * n) + [0, 1], buildTensor((3 * n + 2, 3 * n + 2, 2, 2), matrix), initial),
        State(1e-5), One(1e-2), 1, False), initial=[cplx(x) for x in np.asarray(initial).ravel()]))
    seed(430)
    out["1d.xy"] = guarded(lambda: run_1d(
        System1D([1, 0, 0, 0], [0, 0, 0, 1],
                 buildTensor((4, 4, 2, 2), {(0, 0): Pauli.I, (0, 1): Pauli.X, (0, 2): Pauli.Z, (1, 3): -Pauli.X,
                                            (2, 3): -Pauli.Z, (3, 3): Pauli.I}), np.ones((1, 1, 2))),
        State(1e-5), One(1e-2), 2, False))
    seed(440)
    out["1d.heisenberg"] = guarded(lambda: run_1d(
        System1D([1, 0, 0, 0, 0], [0, 0, 0, 0, 1],
                 buildTensor((5, 5, 2, 2), {(0, 0): Pauli.I, (0, 1): Pauli.X, (0, 2): Pauli.Y, (0, 3): Pauli.Z,
                                            (1, 4): -Pauli.X, (2, 4): -Pauli.Y, (3, 4): Pauli.Z, (4, 4): Pauli.I}),
                 np.ones((1, 1, 2))),
        Estimated(1e-5), Estimated(1e-3), 2, True))


class Registry(dict):
    """`out[name] = guarded(lambda: ...)` in the sections: the section code seeds the RNG right before each assignment,
    so a case is executed here, at assignment time, when it is the selected one (and skipped otherwise)."""

    def __setitem__(self, name, fn):
        CASES.append(name)
        if name == ONLY:
            try:
                dict.__setitem__(self, name, fn())
            except Exception as exc:     # the reference's own failure modes under current SciPy
                dict.__setitem__(self, name, {"reference_error": "%s: %s" % (type(exc).__name__, str(exc)[:200])})


SECTIONS = None


def collect(only=None):
    global ONLY
    ONLY = only
    del CASES[:]
    out = Registry()
    for section in (simulator_2d_in_1d, policies_2d_in_1d, simulator_2d_in_15d, simulator_1d):
        section(out)
    return out


def main():
    import subprocess
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        print("RESULT " + json.dumps(dict(collect(sys.argv[2]))), flush=True)
        return
    limit = float(os.environ.get("CASE_TIMEOUT", "240"))
    collect(None)
    names = list(CASES)
    out = {}
    for name in names:
        t0 = time.time()
        try:
            run = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True,
                                 text=True, timeout=limit)
            lines = [l for l in run.stdout.splitlines() if l.startswith("RESULT ")]
            out[name] = json.loads(lines[-1][7:])[name] if lines else {"reference_error": run.stderr[-300:]}
        except subprocess.TimeoutExpired:
            out[name] = {"reference_error": "did not finish within %.0f s" % limit}
        print("%-45s %6.1f s  %s" % (name, time.time() - t0, out[name]), flush=True)
        with open(os.path.join(HERE, "runs_r2.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
