"""Helpers to read the golden fixtures written by tests/golden/make_golden.py."""
import ast
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def sparse(g, prefix):
    """Rebuild a {tuple tag: ndarray} dict in the reference's insertion order."""
    return {ast.literal_eval(str(r)): g[prefix + "|" + str(r)] for r in g[prefix + "|order"]}


def system_parts(g, prefix):
    corners = [sparse(g, "%s.corner%d" % (prefix, i)) for i in range(4)]
    sides = [sparse(g, "%s.side%d" % (prefix, i)) for i in range(4)]
    return corners, sides, g[prefix + ".center"]


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
