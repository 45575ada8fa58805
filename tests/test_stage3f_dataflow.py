"""Work-level CPU model of the folded fused kernel, driven by the library's real host plans
(carc_stage3f_describe: S blocks, CTA table, slabs, groups; carc_stage3_describe_stars: the star decomposition):
every (CTA, group) walks its slab exactly as the kernel's cursor does, sums first products per star, applies the
site operators, multiplies with the S block of B_x and writes one partial; the partials are summed in slot order.
The result must equal the oracle's matvec (oracle/dense.py: the reference's two tensordots per term).  This pins the
plan and the star algebra on the CPU; the CUDA code itself is pinned by tests/test_gpu_core.py."""
import numpy as np
import pytest

from carcassonne_b200 import _lib
from oracle import dense
from test_stage3_stars import stars
from test_stage3f_plan import describe


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def model_matvec(terms, A, B, ops, v, P, Q, R, S):
    """terms: [(a, b, op index or None)]; A[a]: [X, P, Q]; B[b]: [X, R, S]; v: [Q, S, 2]."""
    X = A[0].shape[0]
    plan = describe(P, Q, R, S, X, nterms=len(terms))
    groups = stars([(a, b) for a, b, _ in terms], [X] * len(terms))
    G, nsb = plan["G"], plan["NSB"]
    slots = np.zeros((plan["slots"], P, R, 2), dtype=complex)
    for cta in range(plan["ctas"]):
        sb, sl = int(plan["cta_sb"][cta]), int(plan["cta_sl"][cta])
        nsl = int(plan["sb_cta0"][sb + 1] - plan["sb_cta0"][sb])
        s0, s1 = 4 * int(plan["sb_tile0"][sb]), min(4 * int(plan["sb_tile0"][sb + 1]), S)
        vb = v[:, s0:s1, :]                                      # the S block of the state
        for g in range(G):
            acc = slots[cta * G + g]
            for kind, members in groups:                        # cursor: star by star, x in the slab, stride G
                for x in range(X * sl // nsl + g, X * (sl + 1) // nsl, G):
                    if kind == 0:                               # B-star: sum the first products, one second product
                        T = np.zeros((P, s1 - s0, 2), dtype=complex)
                        for t in members:
                            a, b, o = terms[t]
                            U = np.einsum("pq,qSs->pSs", A[a][x], vb)
                            T += U if o is None else np.einsum("ts,pSs->pSt", ops[o], U)
                        b = terms[members[0]][1]
                        acc += np.einsum("rS,pSs->prs", B[b][x][:, s0:s1], T)
                    else:                                       # A-star: one first product, reused per term
                        a = terms[members[0]][0]
                        U = np.einsum("pq,qSs->pSs", A[a][x], vb)
                        for t in members:
                            _, b, o = terms[t]
                            W = U if o is None else np.einsum("ts,pSs->pSt", ops[o], U)
                            acc += np.einsum("rS,pSs->prs", B[b][x][:, s0:s1], W)
    return slots.sum(axis=0)


@pytest.mark.parametrize("dims,X", [((2, 2, 2, 2), 7), ((3, 3, 3, 3), 10), ((5, 5, 5, 5), 4), ((6, 6, 6, 6), 3),
                                     ((7, 7, 7, 7), 2), ((2, 3, 3, 2), 9), ((1, 5, 6, 6), 5), ((8, 8, 8, 8), 2)])
def test_model_reproduces_the_oracle_matvec(dims, X):
    d0, d1, d2, d3 = dims
    P, R = d0 * d1, d2 * d3
    rng = np.random.default_rng(P * 100 + R + X)
    na, nb = 4, 4
    s2_0 = [crand(rng, X, 1, d0, d1, d0, d1) for _ in range(na)]      # [x, y, D0, D1, D0*, D1*] with y = 1
    s2_1 = [crand(rng, 1, X, d2, d3, d2, d3) for _ in range(nb)]
    ops = [np.diag([1.0, -1.0]).astype(complex), np.array([[0, 1], [1, 0]], dtype=complex), crand(rng, 2, 2)]
    # a TFIM-like structure: stars on tensor 0 of either half, cross terms, a repeated pair
    terms = [(1, 0, None), (0, 1, None), (0, 0, 0), (2, 0, 1), (3, 0, 2), (0, 2, 1), (0, 3, 2), (3, 2, None), (2, 3, None),
             (2, 3, 0)]
    v = crand(rng, d0, d1, d2, d3, 2)
    ref = sum(dense.stage3_multiply(s2_0[a], s2_1[b], v, None if o is None else ops[o]) for a, b, o in terms)
    halves = [dense.stage3_prejoin(s2_0[a], s2_1[b]) for a, b in [(i, i) for i in range(4)]]
    A = [np.asarray(h[0]).reshape(X, P, P) for h in halves]
    B = [np.asarray(h[1]).reshape(X, R, R) for h in halves]
    out = model_matvec(terms, A, B, ops, v.reshape(P, R, 2), P, P, R, R)
    err = np.linalg.norm(out.reshape(ref.shape) - ref) / np.linalg.norm(ref)
    assert err < 1e-13, err
