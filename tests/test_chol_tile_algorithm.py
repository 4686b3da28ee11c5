"""CPU restatement (numpy, thread by thread) of the register-resident factorisation of the Cholesky's diagonal tile with TWO pivots
per block-wide barrier (sequential-line-search_b200/csrc/chol.cuh: potf2_inverse_regs_pair): the register layout, the shuffles of
the owner warp and the predicates of the two rank-1 updates are the kernel's own, written out for 256 emulated threads. What it
pins down without a GPU is the index logic - every entry of the 64 x 64 array must receive exactly the updates of the unblocked
algorithm, so that the tile comes out as L (lower triangle, after scaling by d^-1/2) and L^-1 (transposed in the strict upper
triangle). The GPU tests then check the kernel itself against the oracle and, bit for bit, against the one-pivot form."""
import numpy as np
import pytest

T = 64


def emulate_pair_pivot_tile(S0):
    c = np.zeros((256, 4, 4))  # c[tid][i][q] = S[tx + 16 i][ty + 16 q], tx = tid & 15, ty = tid >> 4; strict upper triangle = 0
    for tid in range(256):
        tx, ty = tid & 15, tid >> 4
        for i in range(4):
            for q in range(4):
                p, col = tx + 16 * i, ty + 16 * q
                c[tid, i, q] = S0[p, col] if col <= p else 0.0
    dv = np.zeros(T)
    for jq in range(4):
        for m in range(8):
            j = 2 * m + 16 * jq  # pivots j and j + 1: columns of the two half-warps of warp m
            colA, colB = np.zeros(T + 2), np.zeros(T + 2)
            lanes = [m * 32 + lane for lane in range(32)]
            v = np.array([[c[t, i, jq] for i in range(4)] for t in lanes])  # [lane][i]
            rj = 2 * m
            dj, lj1 = v[rj, jq], v[rj + 1, jq]  # __shfl_sync from lanes rj and rj + 1 of the first half-warp
            inv_dj = 1.0 / dj
            for i in range(4):
                a = v[:, i].copy()
                if i == jq:
                    a[rj] = 1.0  # col_j[j] := 1
                for lane in range(16, 32):  # owners of column j + 1 take pivot j (all rows)
                    v[lane, i] = v[lane, i] - (a[lane & 15] * inv_dj) * lj1
            for lane in range(32):
                tx, half = lane & 15, lane >> 4
                if half:
                    for i in range(4):
                        c[lanes[lane], i, jq] = v[lane, i]
                col = colB if half else colA
                for i in range(4):
                    x = v[lane, i]
                    if i == jq and tx == rj + half:
                        col[T] = x
                        dv[j + half] = x
                        x = 1.0
                    col[tx + 16 * i] = x
            inv_a, inv_b = 1.0 / colA[T], 1.0 / colB[T]
            for tid in range(256):  # after the barrier: everybody, both rank-1 updates
                tx, ty = tid & 15, tid >> 4
                ty_gt, tx_ge_ty, tx_le_j, tx_le_j1 = ty > 2 * m + 1, tx >= ty, tx <= 2 * m, tx <= 2 * m + 1
                for i in range(4):
                    for q in range(4):
                        q_gt = q > jq or (q == jq and ty_gt)
                        p_ge_q = i > q or (i == q and tx_ge_ty)
                        p_le_j = i < jq or (i == jq and tx_le_j)
                        p_le_j1 = i < jq or (i == jq and tx_le_j1)
                        if q_gt and (p_ge_q or p_le_j):
                            c[tid, i, q] -= (colA[tx + 16 * i] * inv_a) * colA[ty + 16 * q]
                        if q_gt and (p_ge_q or p_le_j1):
                            c[tid, i, q] -= (colB[tx + 16 * i] * inv_b) * colB[ty + 16 * q]
    rs = 1.0 / np.sqrt(dv)
    L, W = np.zeros((T, T)), np.zeros((T, T))
    for tid in range(256):
        tx, ty = tid & 15, tid >> 4
        for i in range(4):
            for q in range(4):
                p, col = tx + 16 * i, ty + 16 * q
                if col <= p:
                    L[p, col] = c[tid, i, q] * rs[col]
                else:
                    W[col, p] = c[tid, i, q] * rs[col]
                if col == p:
                    W[col, p] = rs[col]
    return L, W


@pytest.mark.parametrize("seed,shift", [(0, 64.0), (1, 1.0), (2, 1e-3)])
def test_pair_pivot_tile_restated_on_the_cpu(seed, shift):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((T, T))
    S0 = A @ A.T + shift * np.eye(T)  # shift = 1e-3: condition number ~1e6
    L, W = emulate_pair_pivot_tile(S0)
    L_ref = np.linalg.cholesky(S0)
    W_ref = np.linalg.inv(L_ref)
    tol = 1e-13 * np.linalg.cond(S0) ** 0.5
    assert np.max(np.abs(L - L_ref)) <= tol * np.max(np.abs(L_ref))
    assert np.max(np.abs(W - W_ref)) <= 10 * tol * np.max(np.abs(W_ref))
    assert np.max(np.abs(L @ L.T - S0)) <= 1e-13 * np.max(np.abs(S0))
    assert not np.triu(L, 1).any() and not np.triu(W, 1).any()
