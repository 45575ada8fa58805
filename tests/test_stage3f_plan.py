"""Host-side launch plan of the folded fused kernel (csrc/stage3f.cu: stage3f_configure) checked on the CPU through
carc_stage3f_describe: the S blocks tile S exactly once, every environment index x is visited exactly once per S block
by the (CTA, group) work items -- walked with the kernel's own slab formula -- and the launch fits the SM."""
import ctypes as C

import numpy as np
import pytest

from carcassonne_b200 import _lib

SMEM_LIMIT = 227 * 1024


def describe(P, Q, R, S, X, nterms=9):
    buf = np.full(16 + 17 + 17 + 160 + 160 + 3, -7, dtype=np.int32)
    rc = _lib.lib.carc_stage3f_describe(nterms, P, Q, R, S, 2, X, buf.ctypes.data, len(buf))
    if rc == _lib.ERR_UNSUPPORTED:
        return None
    assert rc == 0
    keys = ("NPT", "NRT", "Q4", "NSB", "G", "nstA", "nstB", "QS", "BSTR", "b_whole", "threads", "ctas", "slots", "smem",
            "slotA", "slotB")
    k = dict(zip(keys, (int(x) for x in buf[:16])))
    k["sb_tile0"] = buf[16:33]
    k["sb_cta0"] = buf[33:50]
    k["cta_sb"] = buf[50:210]
    k["cta_sl"] = buf[210:370]
    k["PB"], k["RB"] = int(buf[370]), int(buf[371])
    k["ws"] = int(buf[372])
    return k


def check_plan(P, Q, R, S, X):
    k = describe(P, Q, R, S, X)
    assert k is not None
    # the launch fits one SM: one CTA per SM, register-file split of 2 / 3 warps per sub-partition, shared memory
    assert 1 <= k["ctas"] <= 148 and k["slots"] == k["ctas"] * k["G"]
    assert k["threads"] == 32 * k["G"] * k["NPT"] and k["threads"] <= (256 if k["NRT"] >= 7 else 384)
    assert k["smem"] <= SMEM_LIMIT
    # warp-specialised launch: 8 consumer-warp slots (232 registers) where a warp accumulates 7 - 8 column tiles, 12 (152
    # registers) otherwise, plus one producer warpgroup that serves at most four groups; the consumers fit their slots
    assert k["ws"] == (8 if k["NRT"] >= 7 else 12) and k["G"] <= 4 and k["G"] * k["NPT"] <= k["ws"]
    assert (k["ws"] + 4) * 32 <= 512 and 32 * (k["ws"] * (232 if k["ws"] == 8 else 152) + 4 * 40) <= 65536
    # the output in PB x RB blocks of at most NPT x NRT tiles (one block when P, R <= 64)
    assert k["NPT"] == -(-(-(-P // 8)) // k["PB"]) and k["NRT"] == -(-(-(-R // 8)) // k["RB"]) and k["Q4"] == -(-Q // 4)
    assert k["NRT"] <= 8 and (k["PB"], k["RB"]) == (1, 1) or P > 64 or R > 64 or k["PB"] >= 1
    # strides that keep the fragment loads conflict-free, and room for the copies
    assert k["QS"] % 8 == 1 and k["QS"] >= 8 * (k["Q4"] // 2)
    nsb = k["NSB"]
    tiles = k["sb_tile0"][:nsb + 1]
    assert tiles[0] == 0 and tiles[-1] == -(-S // 4) and all(b > a for a, b in zip(tiles, tiles[1:]))
    widest = max(b - a for a, b in zip(tiles, tiles[1:]))
    if k["b_whole"]:
        assert nsb == 1 and k["BSTR"] == S
    else:
        assert k["BSTR"] % 8 == 4 and k["BSTR"] >= 4 * widest
    assert k["slotA"] >= (8 * Q + 4) * 16 and k["slotB"] >= k["NRT"] * 8 * k["BSTR"] * 16
    # CTA table: a permutation of (S block, slab) pairs
    ctas0 = k["sb_cta0"][:nsb + 1]
    assert ctas0[0] == 0 and ctas0[-1] == k["ctas"]
    pairs = {(int(k["cta_sb"][c]), int(k["cta_sl"][c])) for c in range(k["ctas"])}
    assert pairs == {(i, s) for i in range(nsb) for s in range(ctas0[i + 1] - ctas0[i])}
    assert all(x == -1 for x in k["cta_sb"][k["ctas"]:])
    # every x exactly once per S block, walking the kernel's slab formula
    for i in range(nsb):
        nsl = int(ctas0[i + 1] - ctas0[i])
        seen = np.zeros(X, dtype=np.int32)
        for sl in range(nsl):
            lo, hi = X * sl // nsl, X * (sl + 1) // nsl
            for g in range(k["G"]):
                seen[lo + g:hi:k["G"]] += 1
        assert (seen == 1).all()
    return k


@pytest.mark.parametrize("D", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("X", [1, 5, 81, 6561, 65536])
def test_uniform_bond_dimensions(D, X):
    check_plan(D * D, D * D, D * D, D * D, X)


def test_ragged_shapes_and_envelope():
    rng = np.random.default_rng(0)
    for _ in range(200):
        P, R = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        X = int(rng.integers(1, 5000))
        check_plan(P, P, R, R, X)
    # beyond 64 rows / columns the output is computed in blocks (round 2); the plan of every block obeys the same rules
    for P, R in ((65, 8), (8, 72), (81, 81), (100, 100), (121, 121), (144, 144), (90, 80), (72, 55)):
        k = check_plan(P, P, R, R, 700)
        assert k["PB"] * 8 * k["NPT"] >= P and k["RB"] * 8 * k["NRT"] >= R
        assert k["RB"] > 1 or R <= 64
        assert k["PB"] > 1 or P <= 96                  # up to 12 warps (96 rows) fit one block when R is narrow
    assert describe(8, 8, 8, 8, 100, nterms=9)["PB"] == 1


def test_block_shares_follow_the_tile_counts():
    k = check_plan(49, 49, 49, 49, 65536)             # D = 7: 13 tiles as 5 + 4 + 4
    tiles = np.diff(k["sb_tile0"][:k["NSB"] + 1])
    ctas = np.diff(k["sb_cta0"][:k["NSB"] + 1])
    assert list(tiles) == [5, 4, 4]
    assert abs(ctas[0] / ctas[1] - 5 / 4) < 0.05 and ctas.sum() >= 145
    # neighbours in blockIdx stream neighbouring stretches of X
    pos = [(k["cta_sl"][c] + 0.5) / ctas[k["cta_sb"][c]] for c in range(k["ctas"])]
    assert all(b >= a for a, b in zip(pos, pos[1:]))


def test_which_tiling_runs():
    """The measured choice: the folded tiling wherever it issues no more DMMA steps than the
    original one (every uniform D since round 2: 25.8 vs 23.1 TFLOP/s at D = 4); unfused GEMMs outside both envelopes."""
    def path(D, d=2, X=65536, force=0, P=None):
        n = P if P is not None else D * D
        return _lib.lib.carc_stage3_path(9, n, n, n, n, d, X, force)

    assert [path(D) for D in (2, 3, 4, 5, 6, 7, 8)] == [3, 3, 3, 3, 3, 3, 3]
    assert path(4, force=3) == 3 and path(8, force=1) == 1 and path(6, force=2) == 2
    assert path(3, d=3) == 2                      # both fused kernels are specific to d = 2
    # row / column blocks of the output (round 2); beyond two column blocks the unfused GEMMs are faster (measured)
    assert [path(D, X=4096) for D in (9, 10, 11, 12)] == [3, 3, 3, 2] and path(12, X=4096, force=3) == 3
    assert path(5, X=1) == 3
