#!/usr/bin/env python
"""Golden vectors of the reference's 1D (MPS/MPO) recipes (carcassonne/tensors/_1d.py) and of its MPO builder
(carcassonne/sparse.py: makeMPO), produced by importing the UNMODIFIED reference from /root/reference.  Run in the build
container only; writes tests/golden/recipes_1d.npz."""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("CARCASSONNE_REFERENCE", "/root/reference"))
HERE = os.path.dirname(os.path.abspath(__file__))

from carcassonne.data import NDArrayData as ND  # noqa: E402
from carcassonne.sparse import makeMPO  # noqa: E402
from carcassonne.tensors import _1d as t1  # noqa: E402


def main():
    np.random.seed(21)
    o, s, a, d = 3, 4, 5, 2
    out = {}
    L, R = ND.newRandom(o, s, s), ND.newRandom(o, s, s)
    O = ND.newRandom(o, o, d, d)
    S = ND.newRandom(a, s, d)           # center with left bond s
    Sr = ND.newRandom(s, a, d)          # center with right bond s
    out.update(L=L.toArray(), R=R.toArray(), O=O.toArray(), S=S.toArray(), Sr=Sr.toArray())
    out["oss_left"] = t1.absorbCenterOSSIntoLeftEnvironment(L, O, S, S.conj()).toArray()
    out["oss_right"] = t1.absorbCenterOSSIntoRightEnvironment(R, O, Sr, Sr.conj()).toArray()
    L2, R2 = ND.newRandom(s, s), ND.newRandom(s, s)
    out.update(L2=L2.toArray(), R2=R2.toArray())
    out["ss_left"] = t1.absorbCenterSSIntoLeftEnvironment(L2, S, S.conj()).toArray()
    out["ss_right"] = t1.absorbCenterSSIntoRightEnvironment(R2, Sr, Sr.conj()).toArray()
    Rm, Lm, Sc = ND.newRandom(o, s, s), ND.newRandom(o, a, a), ND.newRandom(s, a, d)
    m = t1.formExpectationMultiplier(Rm, Lm, O)
    out.update(Rm=Rm.toArray(), Lm=Lm.toArray(), Sc=Sc.toArray())
    out["mult_out"] = m(Sc).toArray()
    out["mult_matrix"] = m.formMatrix().toArray()
    out["mult_costs"] = np.array([m.cost_of_multiply, m.cost_of_formMatrix])
    # MPO builder: transverse Ising and a two-term operator
    from carcassonne.utils import Pauli
    X, Z, I2 = Pauli.X, Pauli.Z, Pauli.I
    tensor, right, right_tags, left, left_tags = makeMPO(I2, Os=[-Z], OOs=[(X, -0.7 * X)])
    out.update(mpo1_tensor=np.asarray(tensor.toArray() if hasattr(tensor, "toArray") else tensor),
               mpo1_right=np.asarray(right), mpo1_left=np.asarray(left))
    tensor, right, right_tags, left, left_tags = makeMPO(I2, Os=[], OOs=[(X, X), (Z, -Z)])
    out.update(mpo2_tensor=np.asarray(tensor.toArray() if hasattr(tensor, "toArray") else tensor),
               mpo2_right=np.asarray(right), mpo2_left=np.asarray(left))
    np.savez_compressed(os.path.join(HERE, "recipes_1d.npz"), **out)
    print({k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
