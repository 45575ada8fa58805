"""Lane-level model of the folded stage-3 kernel (csrc/stage3f.cu): checks the fragment index maps on the CPU.

It emulates mma.sync.m8n8k4.f64 per lane (A[row=lane/4][k=lane%4], B[k=lane%4][col=lane/4],
C[row=lane/4][col=2*(lane%4)+{0,1}]) and walks one environment index x through the kernel's data flow: Vt rows are
folded columns f = 2*S + s, a T tile's C fragment holds (S = 4j + c; spin 0 in slot 0, spin 1 in slot 1) and is
reused as the A fragment of the second product against B[R = 8 rt + r, S = 4 j + c].  Design check only; the
parity tests in tests/ are what gate the kernel."""
import numpy as np

LANES = np.arange(32)
R_, C_ = LANES >> 2, LANES & 3


def dmma(c0, c1, a, b):
    A = np.zeros((8, 4), dtype=a.dtype)
    B = np.zeros((4, 8), dtype=a.dtype)
    A[R_, C_] = a
    B[C_, R_] = b
    Cm = A @ B
    return c0 + Cm[R_, 2 * C_], c1 + Cm[R_, 2 * C_ + 1]


def model(P, Q, R, S, rng):
    A = rng.standard_normal((P, Q)) + 1j * rng.standard_normal((P, Q))
    B = rng.standard_normal((R, S)) + 1j * rng.standard_normal((R, S))
    v = rng.standard_normal((Q, S, 2)) + 1j * rng.standard_normal((Q, S, 2))
    O = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    ref = np.einsum("rS,ts,pq,qSs->prt", B, O, A, v)

    NPT, NRT, Q4, NTt = -(-P // 8), -(-R // 8), -(-Q // 4), -(-S // 4)
    npairs, tail = Q4 // 2, Q4 & 1
    QS = Q4 * 4
    Ap = np.zeros((NPT * 8, QS + 8), dtype=complex)
    Ap[:P, :Q] = A
    Vt = np.zeros((NTt * 8, QS + 8), dtype=complex)
    for s in range(2):
        Vt[2 * np.arange(S) + s, :Q] = v[:, :, s].T
    Bs = np.zeros((NRT * 8, NTt * 4), dtype=complex)
    Bs[:R, :S] = B
    out = np.zeros((NPT * 8, NRT * 8, 2), dtype=complex)
    for wg in range(NPT):
        acc = np.zeros((NRT, 2, 2, 32), dtype=complex)   # [rt][spin][slot e][lane]
        for j in range(NTt):
            t0 = np.zeros(32, dtype=complex)
            t1 = np.zeros(32, dtype=complex)
            for kp in range(npairs):
                for e in range(2):
                    a = Ap[wg * 8 + R_, kp * 8 + 2 * C_ + e]
                    b = Vt[j * 8 + R_, kp * 8 + 2 * C_ + e]
                    t0, t1 = dmma(t0, t1, a, b)
            if tail:
                a = Ap[wg * 8 + R_, npairs * 8 + C_]
                b = Vt[j * 8 + R_, npairs * 8 + C_]
                t0, t1 = dmma(t0, t1, a, b)
            # site operator inside the lane: slot 0 is spin 0, slot 1 is spin 1
            w0 = O[0, 0] * t0 + O[0, 1] * t1
            w1 = O[1, 0] * t0 + O[1, 1] * t1
            for rt in range(NRT):
                b = Bs[rt * 8 + R_, 4 * j + C_]
                acc[rt, 0, 0], acc[rt, 0, 1] = dmma(acc[rt, 0, 0], acc[rt, 0, 1], w0, b)
                acc[rt, 1, 0], acc[rt, 1, 1] = dmma(acc[rt, 1, 0], acc[rt, 1, 1], w1, b)
        for rt in range(NRT):
            for s in range(2):
                for e in range(2):
                    out[wg * 8 + R_, rt * 8 + 2 * C_ + e, s] = acc[rt, s, e]
    return np.linalg.norm(out[:P, :R] - ref) / np.linalg.norm(ref)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for shape in [(4, 4, 4, 4), (9, 9, 9, 9), (25, 25, 25, 25), (36, 36, 36, 36), (6, 6, 9, 4), (16, 16, 4, 4), (49, 49, 49, 49)]:
        err = model(*shape, rng)
        print(shape, "%.2e" % err)
        assert err < 1e-13
