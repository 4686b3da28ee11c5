// K2: blocked right-looking Cholesky (FP64) with look-ahead, one launch per 64-wide block column.
//
// Launch k (k = -1 .. nb-2) owns the trailing matrix that starts at block k+1. Every CTA updates one 64 x 64 tile
//     A[tm,tn] -= L[tm,k] L[tn,k]^T                       (skipped for k = -1)
// and the tiles of the FIRST trailing block column then finish that column inside the same launch:
//     CTA 0        (tile (k+1,k+1)): Cholesky of the updated diagonal tile and its inverse W, written to L / W,
//                   then publishes flag[k+1];
//     CTAs 1..rem-1 (tiles (tm,k+1)): keep their updated tile in shared memory, wait for flag[k+1] and write
//                   L[tm,k+1] = tile * W^T.
// So when launch k retires, block column k+1 of L is final and launch k+1 can start: the diagonal factorisation
// (the only serial part) overlaps with the bulk of the trailing update instead of sitting between two launches.
// The waiting CTAs are the lowest-numbered ones of the grid together with the CTA they wait for, so they are
// always co-resident.
//
// Thread (tx, ty) of the 16 x 16 thread grid owns the INTERLEAVED micro-tile rows tx + 16 i, columns ty + 16 j
// (i, j < 4): the shared-memory operand reads of a half-warp are then 16 consecutive doubles (conflict-free) and
// its global accesses 128-byte segments.
//
// Replaces Eigen::LLT (reference call sites src/preference-regressor.cpp:162,290,370); W (the inverted diagonal
// blocks) seeds the recursive-doubling triangular inverse of slsgp.cu:do_trtri.
#pragma once

#include "common.cuh"
#include "dense.cuh" // dmma_8x8x4
#include "gram.cuh"  // lower_tile

namespace slsgp
{
    constexpr int CHOL_LDS        = TILE + 4; // shared-memory leading dimension: rows 16-byte aligned, DMMA fragment reads conflict-free
    constexpr int CHOL_SMEM_BYTES = 2 * TILE * CHOL_LDS * (int) sizeof(double);

    // 1 / d to double precision without the IEEE-division sequence: MUFU seed (about 20 bits) + two Newton steps. The pivot
    // loop below waits on this once per column, 64 columns per launch, 32 launches in a row at N = 2048.
    __device__ __forceinline__ double fast_reciprocal(double d)
    {
        double x;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
        x = fma(x, fma(-d, x, 1.0), x);
        x = fma(x, fma(-d, x, 1.0), x);
        return x;
    }

    // Cholesky + inverse of a 64 x 64 SPD tile held in REGISTERS, c[i][q] = S[tx + 16 i][ty + 16 q], lower triangle =
    // the tile, strict upper triangle = 0.
    // Works on the UNSCALED Schur complement: at step j, with d = S[j][j] and column j broadcast through shared memory,
    //     S[p][q] -= S[p][j] S[q][j] / d        (j < q <= p)        trailing update
    //     R[q][p] -= S[q][j] R[j][p] / d        (p <= j < q)        forward substitution L R = I, R[j][j] = 1
    // where R (strictly lower) lives transposed in the strict upper triangle (R[q][p] at S[p][q], p < q), so both
    // cases read  S[p][q] -= col[p] col[q] / d  with col[j] := 1. One barrier per step (the column buffer is
    // double-buffered), at most 16 predicated FMAs per thread, no shared-memory traffic besides the 64-value broadcast.
    // Afterwards L[p][q] = S[p][q] rs[q] (q <= p) and W[q][p] = S[p][q] rs[q] (p < q), rs[q] = d_q^-1/2 in rs[].
    // info receives 1 + global index of the first non-positive pivot.
    __device__ __forceinline__ void potf2_inverse_regs(double c[4][4], double* colbuf /* [2][TILE + 2] */, double* dv,
                                                       double* rs, int* sbad, int tx, int ty, int diag0, int* info)
    {
        if (threadIdx.x == 0) *sbad = 0x7fffffff;
        int first_bad = 0x7fffffff;
#pragma unroll
        for (int jq = 0; jq < 4; ++jq)
        {
            for (int jt = 0; jt < 16; ++jt)
            {
                // pivot j = jt + 16 jq. With p = tx + 16 i and q = ty + 16 qq the predicate
                //   on = (q > j) && (p >= q || p <= j)
                // reduces to compile-time comparisons of i, qq, jq plus these three thread-level facts
                const int  j     = jt + 16 * jq;
                const bool ty_gt = ty > jt, tx_le = tx <= jt, tx_ge_ty = tx >= ty;
                double*    col   = colbuf + (j & 1) * (TILE + 2);
                if (ty == jt) // the 16 threads that own column j
                {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        double v = c[i][jq];
                        if (i == jq && tx == jt) // the pivot itself
                        {
                            col[TILE] = v, dv[j] = v;
                            if (!(v > 0.0)) first_bad = min(first_bad, j);
                            v = 1.0;
                        }
                        col[tx + 16 * i] = v;
                    }
                }
                __syncthreads();
                const double inv_d = fast_reciprocal(col[TILE]);
                double       rowv[4], colv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rowv[i] = col[tx + 16 * i] * inv_d;
#pragma unroll
                for (int q = 0; q < 4; ++q) colv[q] = col[ty + 16 * q];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        const bool q_gt_j = q > jq || (q == jq && ty_gt);
                        const bool p_ge_q = i > q || (i == q && tx_ge_ty);
                        const bool p_le_j = i < jq || (i == jq && tx_le);
                        if (q_gt_j && (p_ge_q || p_le_j)) c[i][q] = fma(-rowv[i], colv[q], c[i][q]);
                    }
            }
        }
        if (first_bad != 0x7fffffff) atomicMin(sbad, first_bad);
        __syncthreads();
        if (threadIdx.x < TILE) rs[threadIdx.x] = rsqrt(dv[threadIdx.x]);
        if (threadIdx.x == 0 && *sbad != 0x7fffffff) atomicCAS(info, 0, 1 + diag0 + *sbad);
        __syncthreads();
    }

    // Same elimination, TWO pivots per barrier. Columns j = 2 m + 16 jq and j + 1 live in the two half-warps of warp m
    // (ty = 2 m and 2 m + 1, lanes 0-15 and 16-31), so the owners of column j + 1 can apply pivot j to their own column with
    // shuffles (d_j, S[j+1][j] and col_j[p] come from the other half-warp) before BOTH columns are broadcast through shared
    // memory; everybody else then applies the two rank-1 updates back to back. Every entry sees the same operations in the same
    // order as in potf2_inverse_regs (results are bit-identical), with 32 block-wide barriers instead of 64: the pivot chain
    // (broadcast, barrier, reciprocal) is paid once per pair. colbuf is [2][2][TILE + 2] here.
    __device__ __forceinline__ void potf2_inverse_regs_pair(double c[4][4], double* colbuf /* [2][2][TILE + 2] */, double* dv,
                                                            double* rs, int* sbad, int tx, int ty, int diag0, int* info)
    {
        if (threadIdx.x == 0) *sbad = 0x7fffffff;
        int       first_bad = 0x7fffffff;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, half = lane >> 4;
#pragma unroll
        for (int jq = 0; jq < 4; ++jq)
        {
            for (int m = 0; m < 8; ++m)
            {
                const int j    = 2 * m + 16 * jq; // pivots j and j + 1
                double*   colA = colbuf + (m & 1) * 2 * (TILE + 2);
                double*   colB = colA + (TILE + 2);
                if (warp == m) // the 32 threads that own columns j (lanes 0-15) and j + 1 (lanes 16-31)
                {
                    const int    rj     = 2 * m; // row j = rj + 16 jq sits in lane rj of the first half-warp, register [jq][jq]
                    const double dj     = __shfl_sync(0xffffffffu, c[jq][jq], rj);
                    const double lj1    = __shfl_sync(0xffffffffu, c[jq][jq], rj + 1); // S[j+1][j]
                    const double inv_dj = fast_reciprocal(dj);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        double a = c[i][jq];
                        if (i == jq && lane == rj) a = 1.0; // col_j[j] := 1
                        const double colj_p = __shfl_sync(0xffffffffu, a, tx);
                        if (half) c[i][jq] = fma(-(colj_p * inv_dj), lj1, c[i][jq]); // pivot j applied to column j + 1 (all rows)
                    }
                    double* col = half ? colB : colA;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        double v = c[i][jq];
                        if (i == jq && tx == rj + half) // the pivot of this half-warp's column
                        {
                            col[TILE] = v, dv[j + half] = v;
                            if (!(v > 0.0)) first_bad = min(first_bad, j + half);
                            v = 1.0;
                        }
                        col[tx + 16 * i] = v;
                    }
                }
                __syncthreads();
                const double inv_a = fast_reciprocal(colA[TILE]), inv_b = fast_reciprocal(colB[TILE]);
                double       rowa[4], rowb[4], cola[4], colb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rowa[i] = colA[tx + 16 * i] * inv_a, rowb[i] = colB[tx + 16 * i] * inv_b;
#pragma unroll
                for (int q = 0; q < 4; ++q) cola[q] = colA[ty + 16 * q], colb[q] = colB[ty + 16 * q];
                // p = tx + 16 i, q = ty + 16 qq. Pivot j acts on q > j + 1 here (column j + 1 has it already), pivot j + 1 on q > j + 1.
                const bool ty_gt = ty > 2 * m + 1, tx_ge_ty = tx >= ty, tx_le_j = tx <= 2 * m, tx_le_j1 = tx <= 2 * m + 1;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        const bool q_gt    = q > jq || (q == jq && ty_gt);
                        const bool p_ge_q  = i > q || (i == q && tx_ge_ty);
                        const bool p_le_j  = i < jq || (i == jq && tx_le_j);
                        const bool p_le_j1 = i < jq || (i == jq && tx_le_j1);
                        if (q_gt && (p_ge_q || p_le_j)) c[i][q] = fma(-rowa[i], cola[q], c[i][q]);
                        if (q_gt && (p_ge_q || p_le_j1)) c[i][q] = fma(-rowb[i], colb[q], c[i][q]);
                    }
            }
        }
        if (first_bad != 0x7fffffff) atomicMin(sbad, first_bad);
        __syncthreads();
        if (threadIdx.x < TILE) rs[threadIdx.x] = rsqrt(dv[threadIdx.x]);
        if (threadIdx.x == 0 && *sbad != 0x7fffffff) atomicCAS(info, 0, 1 + diag0 + *sbad);
        __syncthreads();
    }

    // acc += As^T Bs over k in [0, 64) on the FP64 tensor pipe: As[k][m], Bs[k][n]; warp w owns the 32 (m) x 16 (n) block at
    // (wm, wn) = ((w & 1) * 32, (w >> 1) * 16) as 4 x 2 DMMA tiles; acc[i][j][h] = element (wm + 8 i + lane / 4,
    // wn + 8 j + 2 (lane % 4) + h).
    // lower: only the 8 x 8 DMMA tiles that touch the lower triangle (row tile >= column tile) are computed, the others stay 0.
    __device__ __forceinline__ void tile_dmma_64(double (*As)[CHOL_LDS], double (*Bs)[CHOL_LDS], int wm, int wn, int lane,
                                                 double acc[4][2][2], int k_end = TILE, bool lower = false)
    {
        const int lr = lane >> 2, lc = lane & 3;
        bool      need[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) need[i][j] = !lower || wm + 8 * i >= wn + 8 * j;
#pragma unroll 4
        for (int ks = 0; ks < k_end; ks += 4)
        {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[ks + lc][wm + i * 8 + lr];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Bs[ks + lc][wn + j * 8 + lr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (need[i][j]) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }

    // dst[kk][m] = src[m + kk * ld] for a 64 x 64 column-major tile: 16 independent 8-byte loads in flight per thread.
    __device__ __forceinline__ void load_tile_64(double (*dst)[CHOL_LDS], const double* src, int ld, int tid, bool cg)
    {
        double v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r)
        {
            const int     e = tid + r * 256, m = e & 63, kk = e >> 6;
            const double* p = src + (size_t) m + (size_t) kk * ld;
            v[r]            = cg ? __ldcg(p) : *p;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r)
        {
            const int e = tid + r * 256;
            dst[e >> 6][e & 63] = v[r];
        }
    }

    // Same tile, same layout, as an asynchronous copy (cp.async.cg, 16 bytes per request, 8 requests per thread): nothing passes
    // through registers, so the copies of both operand tiles and the loads of the C tile are all in flight together and a CTA
    // pays one memory latency per tile instead of three. Complete with chol_cp_async_wait_all() + __syncthreads().
    __device__ __forceinline__ void load_tile_64_async(double (*dst)[CHOL_LDS], const double* src, int ld, int tid)
    {
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
            const int      e = tid + r * 256, kk = e >> 5, m = (e & 31) * 2;
            const uint32_t d = (uint32_t) __cvta_generic_to_shared(&dst[kk][m]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + (size_t) m + (size_t) kk * ld) : "memory");
        }
    }
    __device__ __forceinline__ void chol_cp_async_wait_all()
    {
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }

    // Start of a factorisation in ONE launch (it used to be a device-to-device copy, two memsets and, at the end, a kernel
    // zeroing the upper triangle): L = lower tiles of K, zeros in the tiles above the diagonal (the diagonal tiles get their zeros
    // from the step kernel), flags and info cleared. grid: (nb, nb) tiles, 256 threads.
    __global__ void __launch_bounds__(256)
        chol_prepare_kernel(const double* __restrict__ K, double* __restrict__ L, int ld, int nb, int* __restrict__ flags, int* __restrict__ info)
    {
        const int tm = blockIdx.x, tn = blockIdx.y, tid = threadIdx.x;
        if (tm == 0 && tn == 0)
        {
            for (int i = tid; i < nb; i += 256) flags[i] = 0;
            if (tid == 0) *info = 0;
        }
        const size_t base = (size_t) tm * TILE + (size_t) tn * TILE * ld;
        const bool   copy = tn <= tm;
#pragma unroll
        for (int r = 0; r < 8; ++r)
        {
            const int    e = tid + r * 256, row = (e & 31) * 2, col = e >> 5;
            const size_t o = base + (size_t) row + (size_t) col * ld;
            double2      v = make_double2(0.0, 0.0);
            if (copy) v = *reinterpret_cast<const double2*>(K + o);
            *reinterpret_cast<double2*>(L + o) = v;
        }
    }

    // grid: rem (rem + 1) / 2 CTAs for k >= 0, rem CTAs for k = -1 (rem = nb - k - 1); 256 threads;
    // dynamic shared memory CHOL_SMEM_BYTES.
    // Panel form (two-level factorisation of slsgp.cu:do_factor, N >= 4096): pe < nb restricts the trailing update to the block
    // columns k + 1 .. pe - 1 of the current panel (grid: sum over those columns of nb - column); do_upd == 0 skips the update
    // altogether (first step of a panel: its columns already carry every earlier panel through the SYRK launches; grid: rem).
    template <int PIV> // pivots per barrier in the diagonal tile: 2, or 1 (A/B; same layout, bit-identical results)
    __global__ void __launch_bounds__(256, 2)
        chol_step_kernel(double* L, double* W, int ld, int k, int nb, int pe, int do_upd, int* flags, int* info)
    {
        extern __shared__ __align__(16) double csm[];
        double(*As)[CHOL_LDS] = reinterpret_cast<double(*)[CHOL_LDS]>(csm);
        double(*Bs)[CHOL_LDS] = As + TILE;

        // Programmatic dependent launch: let the next step's CTAs become resident as ours retire (they stop at their own
        // griddepcontrol.wait until this whole grid has completed and its writes are visible), then wait for the previous step.
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");

        const int rem = nb - k - 1, t = blockIdx.x, tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
        int       tm, tn;
        if (t < rem)
            tm = k + 1 + t, tn = k + 1;
        else if (pe >= nb)
        {
            int a, b;
            lower_tile(t - rem, a, b);
            tm = k + 2 + a, tn = k + 2 + b;
        }
        else
        {
            // panel form: only the block columns k + 2 .. pe - 1 of the current panel, column by column
            int r = t - rem, col = k + 2;
            while (r >= nb - col) r -= nb - col, ++col;
            tm = col + r, tn = col;
        }
        double* Ct = L + (size_t) tm * TILE + (size_t) tn * TILE * ld;

        const int warp = tid >> 5, lane = tid & 31, wm = (warp & 1) * 32, wn = (warp >> 1) * 16, lr = lane >> 2, lc = lane & 3;
        double    upd[4][2][2]; // L[tm,k] L[tn,k]^T in DMMA fragment layout
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) upd[i][j][0] = upd[i][j][1] = 0.0;

        if (do_upd) // both operand tiles on their way before anything else is touched (the diagonal tile has one: tm == tn)
        {
            load_tile_64_async(As, L + (size_t) tm * TILE + (size_t) k * TILE * ld, ld, tid);
            if (t != 0) load_tile_64_async(Bs, L + (size_t) tn * TILE + (size_t) k * TILE * ld, ld, tid);
        }
        if (t >= rem) // plain trailing tile: C -= update, read and written in fragment layout
        {
            double cf[4][2][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) cf[i][j][h] = Ct[(size_t) (wm + i * 8 + lr) + (size_t) (wn + j * 8 + lc * 2 + h) * ld];
            chol_cp_async_wait_all();
            __syncthreads();
            tile_dmma_64(As, Bs, wm, wn, lane, upd);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        Ct[(size_t) (wm + i * 8 + lr) + (size_t) (wn + j * 8 + lc * 2 + h) * ld] = cf[i][j][h] - upd[i][j][h];
            return;
        }

        // tiles of the next block column: the register-resident factorisation wants the interleaved layout
        double c[4][4]; // the tile itself first (independent of the operand loads), then tile - update
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i][j] = Ct[(size_t) (tx + 16 * i) + (size_t) (ty + 16 * j) * ld];
        if (do_upd)
        {
            chol_cp_async_wait_all();
            __syncthreads();
            // The update of the DIAGONAL tile is symmetric and only its lower triangle is used: 36 of the 64 DMMA tiles. The eight
            // 32 x 16 warp blocks hold 8, 8, 7, 7, 3, 3, 0, 0 of them; they are dealt so that the two warps of each scheduler
            // (warp, warp + 4) hold 8 + 0 or 7 + 3 (blocks by warp: (32,0) (32,16) (0,0) (32,32) | (0,32) (0,48) (0,16) (32,48)).
            const int bm = t == 0 ? ((0x8B >> warp) & 1) * 32 : wm, bn = t == 0 ? ((0xDE84 >> (2 * warp)) & 3) * 16 : wn;
            if (t == 0)
                tile_dmma_64(As, As, bm, bn, lane, upd, TILE, true);
            else
                tile_dmma_64(As, Bs, bm, bn, lane, upd);
            __syncthreads(); // operands consumed: As becomes the re-layout buffer, As[n][m] = update(m, n)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) As[bn + j * 8 + lc * 2 + h][bm + i * 8 + lr] = upd[i][j][h];
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) c[i][j] -= As[ty + 16 * j][tx + 16 * i];
        }

        __syncthreads(); // everyone is done reading As / Bs
        if (t == 0)
        {
            // next diagonal tile: factorise + invert in registers, publish L, W and the flag
            double* colbuf = csm; // [2][2][TILE + 2]
            double* dv     = csm + 4 * (TILE + 2);
            double* rs     = dv + TILE;
            int*    sbad   = reinterpret_cast<int*>(rs + TILE);
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (ty + 16 * j > tx + 16 * i) c[i][j] = 0.0;
            if (PIV == 2)
                potf2_inverse_regs_pair(c, colbuf, dv, rs, sbad, tx, ty, tn * TILE, info);
            else
                potf2_inverse_regs(c, colbuf, dv, rs, sbad, tx, ty, tn * TILE, info);
            double* Wt = W + (size_t) tn * TILE * ((size_t) ld + 1);
            // W first: it is all the panel CTAs below are waiting for. W[q][p] = R[q][p] rs[q] = S[p][q] rs[q] for p < q:
            // transposed through shared memory so the store coalesces
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const int p = tx + 16 * i, q = ty + 16 * j;
                    Bs[p][q]    = q > p ? c[i][j] * rs[q] : (q == p ? rs[q] : 0.0); // Bs[column of W][row of W]
                }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 16; ++r)
            {
                const int e = tid + r * 256, row = e & 63, cl = e >> 6;
                Wt[(size_t) row + (size_t) cl * ld] = Bs[cl][row];
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicExch(&flags[tn], 1);
            // then L[p][q] = S[p][q] rs[q] for q <= p (the diagonal is d * d^-1/2), 0 above; written straight from registers.
            // Nobody inside the factorisation reads the diagonal tile of L again (the panels use W).
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const int p = tx + 16 * i, q = ty + 16 * j;
                    Ct[(size_t) p + (size_t) q * ld] = q <= p ? c[i][j] * rs[q] : 0.0;
                }
            return;
        }

        // panel tile below the next diagonal: L[tm,tn] = C * W^T once W is published
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) As[ty + 16 * j][tx + 16 * i] = c[i][j]; // As[n][m] = C(m, n)
        if (tid == 0)
        {
            // acquire loads, not atomics: up to nb - 1 CTAs poll this word and read-modify-writes on one address would queue up in
            // front of the publisher's store
            int seen;
            do
            {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flags + tn) : "memory");
                if (!seen) __nanosleep(20);
            } while (!seen);
            __threadfence();
        }
        __syncthreads();
        // Bs[n][cc] = W(cc, n): W is column-major, so row n of Bs is column n of W
        load_tile_64(Bs, W + (size_t) tn * TILE * ((size_t) ld + 1), ld, tid, true);
        __syncthreads();
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        // W is lower-triangular: output columns [wc, wc + 16) need the contraction range [0, wc + 16) only. The four column
        // strips are dealt so that the two warps of each scheduler (warp and warp + 4) hold 16 + 64 or 32 + 48 rows of it.
        const int wsel = warp >> 1, wc = wsel == 2 ? 48 : (wsel == 3 ? 32 : wsel * 16);
        tile_dmma_64(As, Bs, wm, wc, lane, acc, wc + 16);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) Ct[(size_t) (wm + i * 8 + lr) + (size_t) (wc + j * 8 + lc * 2 + h) * ld] = acc[i][j][h];
    }
} // namespace slsgp
