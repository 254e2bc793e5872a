"""Host-side emulation of the index algebra of csrc/pw_stream.cu (no GPU): the PTX fragment layouts of
mma.sync.m16n8k16 and ldmatrix.trans are written out lane by lane and the kernel's load / weight / store
permutations are replayed on top of them, then compared with a plain matrix product.  Catches mapping mistakes
before GPU time is spent; it does not execute the CUDA code."""
import numpy as np


def mma16816(a, b, d):
    """a[32][4][2], b[32][2][2], d[32][4]: per-lane fragments (register, element).  Returns updated d."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for e in range(2):
            A[g, 2 * t + e] = a[lane][0][e]; A[g + 8, 2 * t + e] = a[lane][1][e]
            A[g, 2 * t + 8 + e] = a[lane][2][e]; A[g + 8, 2 * t + 8 + e] = a[lane][3][e]
            B[2 * t + e, g] = b[lane][0][e]; B[2 * t + 8 + e, g] = b[lane][1][e]
    C = A @ B
    out = np.array(d, dtype=float)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out[lane][0] += C[g, 2 * t]; out[lane][1] += C[g, 2 * t + 1]
        out[lane][2] += C[g + 8, 2 * t]; out[lane][3] += C[g + 8, 2 * t + 1]
    return out


def rq(NT, q): return 4 if q < NT // 4 else NT % 4
def phys(NT, j, tq, e): return 32 * (j >> 2) + 2 * rq(NT, j >> 2) * tq + 2 * (j & 3) + e


def fwd(K, NO, rng):
    KS, NT = (K + 15) // 16, NO // 8
    X = rng.standard_normal((16, K)); Wt = rng.standard_normal((NO, K))
    out = np.full((16, NO), np.nan)
    d = np.zeros((NT, 32, 4))
    for s in range(KS):
        a = np.zeros((32, 4, 2))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            ch = 16 * s + 4 * t
            lo = X[g, ch:ch + 4] if ch < K else np.zeros(4)
            hi = X[g + 8, ch:ch + 4] if ch < K else np.zeros(4)
            a[lane][0] = lo[0:2]; a[lane][1] = hi[0:2]; a[lane][2] = lo[2:4]; a[lane][3] = hi[2:4]
        for j in range(NT):
            b = np.zeros((32, 2, 2))
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                ch, n = 16 * s + 4 * t, phys(NT, j, g >> 1, g & 1)
                v = Wt[n, ch:ch + 4] if ch < K else np.zeros(4)
                b[lane][0] = v[0:2]; b[lane][1] = v[2:4]
            d[j] = mma16816(a, b, d[j])
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for h in range(2):
            row = g + 8 * h
            for q in range((NT + 3) // 4):
                R = rq(NT, q)
                base = 32 * q + 2 * R * t
                for jj in range(R):
                    assert np.isnan(out[row, base + 2 * jj]) and np.isnan(out[row, base + 2 * jj + 1])
                    out[row, base + 2 * jj] = d[4 * q + jj][lane][2 * h]
                    out[row, base + 2 * jj + 1] = d[4 * q + jj][lane][2 * h + 1]
                    # the statistics / bias index used by the kernel for this element
                    assert phys(NT, 4 * q + jj, t, 0) == base + 2 * jj
    ref = X @ Wt.T
    assert not np.isnan(out).any()
    return np.abs(out - ref).max()


def ldmatrix_x4_trans(smem, addr):
    """smem: 2-D array of bf16 elements [row][col]; addr[lane] = (row, col) of the 8-element row this lane points at.
    Returns r[32][4][2]."""
    r = np.zeros((32, 4, 2))
    for mat in range(4):
        S = np.stack([smem[addr[8 * mat + i][0], addr[8 * mat + i][1]:addr[8 * mat + i][1] + 8] for i in range(8)])
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            r[lane][mat] = [S[2 * t][g], S[2 * t + 1][g]]         # transposed distribution
    return r


def wgrad(Cin, Cout, rng):
    MC, NC = (Cout + 15) // 16, Cin // 8
    DZ = np.zeros((16, 16 * MC)); DZ[:, :Cout] = rng.standard_normal((16, Cout))
    X = np.zeros((16, 8 * NC + 16)); X[:, :Cin] = rng.standard_normal((16, Cin))
    acc = np.zeros((MC, NC, 32, 4))
    bfr = {}
    for n in range(0, NC, 2):
        addr = []
        for lane in range(32):
            lm, lr = lane >> 3, lane & 7
            addr.append((lr + (8 if lm & 1 else 0), (8 if lm & 2 else 0) + 8 * n))
        r4 = ldmatrix_x4_trans(X, addr)
        bfr[n] = r4[:, 0:2]; bfr[n + 1] = r4[:, 2:4]
    for m in range(MC):
        addr = []
        for lane in range(32):
            lm, lr = lane >> 3, lane & 7
            addr.append((lr + (8 if lm & 2 else 0), (8 if lm & 1 else 0) + 16 * m))
        a = ldmatrix_x4_trans(DZ, addr)
        for n in range(NC):
            acc[m][n] = mma16816(a, bfr[n], acc[m][n])
    dw = np.zeros((16 * MC, Cin))
    for m in range(MC):
        for n in range(NC):
            for lane in range(32):
                g, t = lane >> 2, lane & 3
                co, ci = 16 * m + g, 8 * n + 2 * t
                dw[co, ci] += acc[m][n][lane][0]; dw[co, ci + 1] += acc[m][n][lane][1]
                dw[co + 8, ci] += acc[m][n][lane][2]; dw[co + 8, ci + 1] += acc[m][n][lane][3]
    ref = DZ[:, :Cout].T @ X[:, :Cin]
    return np.abs(dw[:Cout] - ref).max()


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for K, NO in [(16, 48), (48, 16), (32, 16), (16, 32), (24, 72), (72, 24), (40, 120), (8, 8)]:
        print("fwd/dgrad", K, NO, fwd(K, NO, rng))
    for Cin, Cout in [(16, 48), (48, 16), (32, 16), (24, 72), (72, 24)]:
        print("wgrad", Cin, Cout, wgrad(Cin, Cout, rng))
