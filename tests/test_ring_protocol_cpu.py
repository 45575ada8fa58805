"""Discrete-event model of the producer / consumer ring protocol of the warp-specialised kernels (csrc/stage3f.cu, WS > 0;
csrc/zgemm.cu follows the same scheme with one ring): mbarrier phases and parities, per-warp A rings with full and empty
barriers, the shared B ring whose empty barrier counts one arrival per consumer warp, (slot, parity) counters that flip
on wrap-around, the producers' "first round is free" parity.  Random interleavings of one producer and NPT consumers over
random op sequences must neither deadlock nor let a consumer read a slot that holds another op's data, nor let the
producer overwrite a slot that a consumer has not released."""
import random

import pytest


class MBarrier:
    """mbarrier with a fixed arrival count: phase `phases` is the current (incomplete) one; a wait on parity p succeeds
    once the phase of that parity has completed, i.e. (phases & 1) != p."""

    def __init__(self, count):
        self.count, self.pending, self.phases = count, count, 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        if self.pending == 0:
            self.pending = self.count
            self.phases += 1

    def done(self, parity):
        return (self.phases & 1) != parity


def producer(ops, npt, nstA, nstB, fullA, emptyA, fullB, emptyB, slotsA, slotsB, in_flight):
    slotA, parA, slotB, parB = 0, 1, 0, 1          # parities of the EMPTY barriers: the first round needs no release
    for op_id, kind in enumerate(ops):
        if kind == "A":
            for cw in range(npt):                  # (in the kernel: lane cw of the producer warp)
                while not emptyA[cw][slotA].done(parA):
                    yield
                assert slotsA[cw][slotA][1], "A slot overwritten before its consumer released it"
                slotsA[cw][slotA] = [None, False]
                in_flight.append(("A", cw, slotA, op_id))   # the copy lands later and completes the full barrier
            slotA += 1
            if slotA == nstA:
                slotA, parA = 0, parA ^ 1
        else:
            while not emptyB[slotB].done(parB):
                yield
            assert slotsB[slotB][1] == npt, "B slot overwritten before every consumer released it"
            slotsB[slotB] = [None, 0]
            in_flight.append(("B", None, slotB, op_id))
            slotB += 1
            if slotB == nstB:
                slotB, parB = 0, parB ^ 1
        yield


def consumer(cw, ops, nstA, nstB, fullA, emptyA, fullB, emptyB, slotsA, slotsB, log):
    sa, pa, sb, pb = 0, 0, 0, 0
    for op_id, kind in enumerate(ops):
        if kind == "A":
            while not fullA[cw][sa].done(pa):
                yield
            assert slotsA[cw][sa][0] == op_id, "consumer %d read A data of op %s for op %d" % (cw, slotsA[cw][sa][0], op_id)
            yield                                   # the DMMAs of the op
            slotsA[cw][sa][1] = True
            emptyA[cw][sa].arrive()
            sa += 1
            if sa == nstA:
                sa, pa = 0, pa ^ 1
        else:
            while not fullB[sb].done(pb):
                yield
            assert slotsB[sb][0] == op_id, "consumer %d read B data of op %s for op %d" % (cw, slotsB[sb][0], op_id)
            yield
            slotsB[sb][1] += 1
            emptyB[sb].arrive()
            sb += 1
            if sb == nstB:
                sb, pb = 0, pb ^ 1
        log.append((cw, op_id))
        yield


@pytest.mark.parametrize("npt, nstA, nstB", [(1, 2, 2), (2, 2, 3), (4, 3, 4), (7, 2, 3), (8, 2, 3), (8, 3, 3)])
@pytest.mark.parametrize("seed", range(6))
def test_ring_protocol(npt, nstA, nstB, seed):
    rng = random.Random(1000 * seed + 100 * npt + 10 * nstA + nstB)
    # op sequences as the planner emits them: stars of one or more A-ops followed by a B-op, or one A-op and several B-ops
    ops = []
    for _ in range(rng.randint(5, 40)):
        if rng.random() < 0.5:
            ops += ["A"] * rng.randint(1, 5) + ["B"]
        else:
            ops += ["A"] + ["B"] * rng.randint(1, 4)
    fullA = [[MBarrier(1) for _ in range(nstA)] for _ in range(npt)]
    emptyA = [[MBarrier(1) for _ in range(nstA)] for _ in range(npt)]
    fullB = [MBarrier(1) for _ in range(nstB)]
    emptyB = [MBarrier(npt) for _ in range(nstB)]
    slotsA = [[[None, True] for _ in range(nstA)] for _ in range(npt)]      # [op whose data the slot holds, released]
    slotsB = [[None, npt] for _ in range(nstB)]                              # [op, number of consumers that released it]
    in_flight, log = [], []
    actors = [producer(ops, npt, nstA, nstB, fullA, emptyA, fullB, emptyB, slotsA, slotsB, in_flight)]
    actors += [consumer(cw, ops, nstA, nstB, fullA, emptyA, fullB, emptyB, slotsA, slotsB, log) for cw in range(npt)]
    alive = list(range(len(actors)))
    steps = 0
    while alive:
        steps += 1
        assert steps < 400000, "deadlock: %d actors never finished" % len(alive)
        if in_flight and rng.random() < 0.5:       # a copy lands (any order: copies of different slots are independent)
            kind, cw, slot, op_id = in_flight.pop(rng.randrange(len(in_flight)))
            if kind == "A":
                slotsA[cw][slot][0] = op_id
                fullA[cw][slot].arrive()
            else:
                slotsB[slot][0] = op_id
                fullB[slot].arrive()
            continue
        i = rng.choice(alive)
        try:
            next(actors[i])
        except StopIteration:
            alive.remove(i)
    assert not in_flight
    assert sorted(log) == [(cw, op) for cw in range(npt) for op in range(len(ops))]
