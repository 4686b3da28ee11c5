"""CPU restatement (numpy, block level) of the launch schedule of the two-level Cholesky (sequential-line-search_b200/csrc/slsgp.cu:
do_factor): panels factored by the step kernel restricted to the panel, one rank-(64 pw) update per panel split into "the next
panel's columns" and "everything beyond", and the single-level sweep for the last block columns. Every launch is replayed as the
block operation it performs, in issue order of the two streams (any order the events allow gives the same numbers, the launches
that may overlap touch disjoint blocks - asserted here). What it pins down without a GPU is the index arithmetic of the schedule
(c0, pe, r1, switch point, ragged last panel): the result must be the Cholesky factor. The kernels themselves are checked on the GPU
(tests/test_gpu_large_parity.py::test_two_level_cholesky_equals_the_single_level_one)."""
import numpy as np
import pytest

B = 8  # emulated tile size (the library's is 64; the schedule only knows block indices)


def blk(A, i, j):
    return A[i * B:(i + 1) * B, j * B:(j + 1) * B]


def step(L, nb, k, pe, do_upd, writes):
    """chol_step_kernel: update the tiles of block columns k+1 .. pe-1 (rows >= column) with block column k, then finish block
    column k+1 (diagonal factor, panel below it)."""
    cols = range(k + 1, min(pe, nb))
    if do_upd:
        for tn in cols:
            for tm in range(tn, nb):
                blk(L, tm, tn)[:] -= blk(L, tm, k) @ blk(L, tn, k).T
                writes.add((tm, tn))
    c = k + 1
    D = np.linalg.cholesky(np.tril(blk(L, c, c)) + np.tril(blk(L, c, c), -1).T)
    blk(L, c, c)[:] = D
    writes.add((c, c))
    Winv = np.linalg.inv(D)
    for tm in range(c + 1, nb):
        blk(L, tm, c)[:] = blk(L, tm, c) @ Winv.T
        writes.add((tm, c))


def syrk(L, nb, c0, pw, r0, ncols, writes):
    """gemm64_dmma_kernel<N, T>, lower tiles: C[r0.., r0 .. r0+ncols) -= P P^T with P = block columns c0 .. c0+pw-1."""
    for tn in range(r0, r0 + ncols):
        for tm in range(tn, nb):
            for u in range(c0, c0 + pw):
                blk(L, tm, tn)[:] -= blk(L, tm, u) @ blk(L, tn, u).T
            writes.add((tm, tn))


def two_level(K, nb, PB, switch_rem):
    L = np.tril(K).copy()
    c0 = 0
    while c0 < nb:
        if nb - c0 <= switch_rem:  # the single-level sweep finishes the matrix
            w = set()
            step(L, nb, c0 - 1, nb, False, w)
            for k in range(c0, nb - 1):
                step(L, nb, k, nb, True, w)
            break
        pw = min(PB, nb - c0)
        pe = c0 + pw
        w_hi = set()
        step(L, nb, c0 - 1, pe, False, w_hi)
        for k in range(c0, pe - 1):
            step(L, nb, k, pe, True, w_hi)
        assert all(c0 <= tn < pe for (_, tn) in w_hi)  # a panel's steps stay inside the panel
        if pe >= nb:
            break
        pw2 = min(PB, nb - pe)
        r1 = pe + pw2
        w_next, w_rest = set(), set()
        syrk(L, nb, c0, pw, pe, pw2, w_next)              # critical stream: the next panel's columns
        if r1 < nb:
            syrk(L, nb, c0, pw, r1, nb - r1, w_rest)      # second stream: everything beyond, under the next panel's steps
        assert not (w_next & w_rest)
        assert all(pe <= tn < r1 for (_, tn) in w_next) and all(tn >= r1 for (_, tn) in w_rest)
        c0 += PB
    return np.tril(L)


@pytest.mark.parametrize("nb,PB,switch_rem", [(11, 2, 0), (11, 4, 0), (11, 3, 0), (18, 4, 6), (18, 3, 7), (18, 2, 5), (13, 8, 3), (9, 16, 0), (12, 4, 12)])
def test_two_level_schedule_restated_on_the_cpu(nb, PB, switch_rem):
    rng = np.random.default_rng(nb * 100 + PB)
    n = nb * B
    A = rng.standard_normal((n, n))
    K = A @ A.T + n * np.eye(n)
    L = two_level(K, nb, PB, switch_rem)
    L_ref = np.linalg.cholesky(K)
    assert np.max(np.abs(L - L_ref)) <= 1e-12 * np.max(np.abs(L_ref))
