"""Lane-level NumPy emulation of the tile algorithm of k_prepare_tiles (prepare_tiles.cu): validates the fragment-order
shared-memory layout, the mma.m8n8k4 operand / accumulator lane mappings, the C->A register shuffle and the
left-looking Cholesky / row-wise inverse schedules against numpy.linalg before anything runs on a GPU.
Developer tool; not part of the product."""
import numpy as np

LANES = np.arange(32)


# ---- mma.sync.m8n8k4.f64 lane mappings (PTX ISA): A 8x4 row, B 4x8 col, C 8x8 ----
def dmma(c, a, b):
    """c: (32,2) accumulators, a: (32,) A fragment, b: (32,) B fragment."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    for l in range(32):
        A[l >> 2, l & 3] = a[l]
        B[l & 3, l >> 2] = b[l]
        C[l >> 2, 2 * (l & 3)] = c[l, 0]; C[l >> 2, 2 * (l & 3) + 1] = c[l, 1]
    C = C + A @ B
    out = np.zeros((32, 2))
    for l in range(32):
        out[l, 0] = C[l >> 2, 2 * (l & 3)]; out[l, 1] = C[l >> 2, 2 * (l & 3) + 1]
    return out


# ---- tile storage: two 8x4 halves in fragment order: element [r][c] at (c>>2)*32 + r*4 + (c&3) ----
def pos(r, c):
    return (c >> 2) * 32 + r * 4 + (c & 3)


def store_tile(M8):
    t = np.zeros(64)
    for r in range(8):
        for c in range(8):
            t[pos(r, c)] = M8[r, c]
    return t


def load_tile(t):
    M8 = np.zeros((8, 8))
    for r in range(8):
        for c in range(8):
            M8[r, c] = t[pos(r, c)]
    return M8


def frag_direct(t, h):
    """lane l reads word h*32 + l: element [l>>2][(l&3)+4h].  As A operand: tile[:, 4h:4h+4].  As B operand: B[k][n] =
    tile[n][k+4h], i.e. the operand (tile^T)[4h:4h+4, :]."""
    return t[h * 32 + LANES]


def frag_bT(t, h):
    """B operand = tile[4h:4h+4, :] itself (not transposed): lane (kk = l&3, n = l>>2) reads element [kk+4h][n]."""
    return np.array([t[pos((l & 3) + 4 * h, l >> 2)] for l in range(32)])


def c_load(t):
    """accumulator layout from a stored tile: lane (r, q) <- [r][2q], [r][2q+1] (one 16-byte load)."""
    return np.array([[t[pos(l >> 2, 2 * (l & 3))], t[pos(l >> 2, 2 * (l & 3) + 1)]] for l in range(32)])


def c_store(c):
    t = np.zeros(64)
    for l in range(32):
        base = pos(l >> 2, 2 * (l & 3))
        assert pos(l >> 2, 2 * (l & 3) + 1) == base + 1 and base % 2 == 0  # one aligned 16-byte store
        t[base] = c[l, 0]; t[base + 1] = c[l, 1]
    return t


def c_to_a(c, h):
    """A fragment (8x4 half h of the accumulator tile) by register shuffles: src lane (l & ~3) | (2h + ((l&3)>>1)), reg l&1."""
    out = np.zeros(32)
    for l in range(32):
        src = (l & ~3) | (2 * h + ((l & 3) >> 1))
        out[l] = c[src, l & 1]
    return out


def chol8(A8):
    L = np.linalg.cholesky(A8)
    return L, np.linalg.inv(L)


def run(n=50, seed=0):
    rng = np.random.default_rng(seed)
    T = (n + 7) // 8
    npad = 8 * T
    G = rng.normal(size=(n, n)); A = G @ G.T + n * np.eye(n)
    Ap = np.eye(npad); Ap[:n, :n] = A
    e = rng.normal(size=n); y = rng.normal(size=n)
    slot = {}
    for I in range(T):
        for J in range(I + 1):
            slot[(I, J)] = store_tile(Ap[8 * I:8 * I + 8, 8 * J:8 * J + 8])
    for J in range(T):  # border row tile T: rows 0, 1 = e, Y
        Bt = np.zeros((8, 8)); Bt[0, :] = np.pad(e, (0, npad - n))[8 * J:8 * J + 8]; Bt[1, :] = np.pad(y, (0, npad - n))[8 * J:8 * J + 8]
        slot[(T, J)] = store_tile(Bt)
    W = {}
    # ---- left-looking factorisation ----
    for J in range(T):
        held = {}
        for I in range(J, T + 1):
            acc = np.zeros((32, 2))
            for K in range(J):
                for h in (0, 1):
                    acc = dmma(acc, frag_direct(slot[(I, K)], h), frag_direct(slot[(J, K)], h))  # += L[I][K] L[J][K]^T
            held[I] = c_load(slot[(I, J)]) - acc
        Ljj, Wj = chol8(load_tile(c_store(held[J])) * np.tri(8) + np.tril(load_tile(c_store(held[J])), -1).T * 0)
        slot[(J, J)] = store_tile(Ljj)
        W[J] = store_tile(Wj)
        for I in range(J + 1, T + 1):  # TRSM: L[I][J] = C W_J^T
            acc = np.zeros((32, 2))
            for h in (0, 1):
                acc = dmma(acc, c_to_a(held[I], h), frag_direct(W[J], h))
            slot[(I, J)] = c_store(acc)
    Lfull = np.zeros((npad, npad))
    for I in range(T):
        for J in range(I + 1):
            Lfull[8 * I:8 * I + 8, 8 * J:8 * J + 8] = load_tile(slot[(I, J)])
    Lref = np.linalg.cholesky(Ap)
    print("chol err", np.abs(Lfull - Lref).max())
    a = np.concatenate([load_tile(slot[(T, J)])[0] for J in range(T)]); z = np.concatenate([load_tile(slot[(T, J)])[1] for J in range(T)])
    print("a err", np.abs(a[:n] - np.linalg.solve(Lref[:n, :n], e)).max(), "z err", np.abs(z[:n] - np.linalg.solve(Lref[:n, :n], y)).max())
    # ---- inverse, row by row, in place ----
    for I in range(T):
        newN = {}
        for K in range(I):  # N[I][K] = W_I L[I][K]
            acc = np.zeros((32, 2))
            for h in (0, 1):
                acc = dmma(acc, frag_direct(W[I], h), frag_bT(slot[(I, K)], h))
            newN[K] = c_store(acc)
        for K in range(I):
            slot[(I, K)] = newN[K]
        res = {}
        for J in range(I):  # M[I][J] = - sum_{K=J}^{I-1} N[I][K] M[K][J]   (M[J][J] = W_J already in slot (J, J))
            acc = np.zeros((32, 2))
            for K in range(J, I):
                for h in (0, 1):
                    acc = dmma(acc, frag_direct(slot[(I, K)], h), frag_bT(slot[(K, J)], h))
            res[J] = c_store(-acc)
        for J in range(I):
            slot[(I, J)] = res[J]
        slot[(I, I)] = W[I].copy()
    Mfull = np.zeros((npad, npad))
    for I in range(T):
        for J in range(I + 1):
            Mfull[8 * I:8 * I + 8, 8 * J:8 * J + 8] = load_tile(slot[(I, J)])
    print("inv err", np.abs(Mfull - np.linalg.inv(Lref)).max())
    # k_score's M layout: block (R, g) element [rr*4 + kk] = M[8R+rr][4g+kk] == half h = g&1 of slot (R, g>>1)
    for (R, g) in ((T - 1, 0), (T - 1, 2 * T - 1), (3, 5)):
        blk = slot[(R, g >> 1)][(g & 1) * 32:(g & 1) * 32 + 32]
        ref = np.array([Mfull[8 * R + (i >> 2), 4 * g + (i & 3)] for i in range(32)])
        assert np.array_equal(blk, ref)
    print("layout ok")


if __name__ == "__main__":
    run(50)
    run(64, 1)
