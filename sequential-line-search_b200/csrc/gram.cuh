// K1: pairwise ARD kernel tiles. One 64 x 64 tile per CTA (256 threads, 4 x 4 outputs per thread); the D-dim
// inputs of both sides are staged in shared memory pre-divided by the length scales, so the inner loop is one
// subtract + one FMA per (pair, dim). Used for
//   - the Gram matrix K_y (symmetric: lower tiles computed, mirrored through shared memory so both stores coalesce),
//   - the MAP hyper-gradient weight matrix Wm = (alpha alpha^T - K^-1) o kl(r2)        (map.cuh),
//   - the cross-covariance k* / g* of a block of candidates against the data           (sweep.cuh).
#pragma once

#include "common.cuh"

namespace slsgp
{
    constexpr int DCHUNK = 16;

    // Accumulate r2[i][j] = sum_d (A[d][tx*4+i] - B[d][ty*4+j])^2 over all D, staging DCHUNK dims at a time.
    // PA: D x (>= 64*tile+64) column-major points of side A (ld = ldA), likewise PB. Points >= nA / nB read as 0.
    __device__ __forceinline__ void tile_sq_dist(const double* __restrict__ PA, int ldA, int nA, int baseA,
                                                 const double* __restrict__ PB, int ldB, int nB, int baseB, int D,
                                                 const double* __restrict__ inv_l, double (*sa)[TILE],
                                                 double (*sb)[TILE], double r2[4][4])
    {
        const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) r2[i][j] = 0.0;
        for (int d0 = 0; d0 < D; d0 += DCHUNK)
        {
            // 64 points x DCHUNK dims per side = 1024 values, 4 per thread; consecutive threads read consecutive d
            // of one point (the points are D-contiguous in memory).
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int e = tid + r * 256, dd = e & (DCHUNK - 1), p = e >> 4;
                const int d = d0 + dd;
                double    va = 0.0, vb = 0.0;
                if (d < D)
                {
                    const double s = inv_l[d];
                    if (baseA + p < nA) va = PA[(size_t) d + (size_t) (baseA + p) * ldA] * s;
                    if (baseB + p < nB) vb = PB[(size_t) d + (size_t) (baseB + p) * ldB] * s;
                }
                sa[dd][p] = va;
                sb[dd][p] = vb;
            }
            __syncthreads();
#pragma unroll
            for (int dd = 0; dd < DCHUNK; ++dd)
            {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sa[dd][tx * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = sb[dd][ty * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const double df = a[i] - b[j];
                        r2[i][j]        = fma(df, df, r2[i][j]);
                    }
            }
            __syncthreads();
        }
    }

    // ---- Gram-matrix kernel (K1), second layout ---------------------------------------------------------------------
    // Ownership inside a 64 x 64 tile: thread (tx, ty) = (tid & 15, tid >> 4) holds rows tx + 16 i and columns 4 ty + j
    // (i, j < 4). With the rows interleaved like this
    //   * the 16 threads of a half-warp read 16 CONSECUTIVE doubles of a staged row (one wavefront, no bank conflict) where the
    //     4 x 4 blocked ownership read with a stride of 4 doubles (4-way conflict),
    //   * a store instruction of a warp covers two full 128-byte lines of the output tile,
    //   * the transposing store of the mirrored tile touches every bank pair exactly twice (the minimum for 256 bytes).
    // The staging rows are padded to 65 doubles: the coalesced global read hands consecutive threads consecutive DIMENSIONS of
    // one point, i.e. a shared-memory stride of one row, and 65 is odd (64 put all 16 of them in one bank pair: the 16-way
    // conflict that made 63 % of the wavefronts of the first kernel replays).
    constexpr int GPAD = TILE + 1;

    __device__ __forceinline__ void tile_sq_dist_interleaved(const double* __restrict__ PA, int ldA, int nA, int baseA,
                                                             const double* __restrict__ PB, int ldB, int nB, int baseB, int D,
                                                             const double* __restrict__ inv_l, double* __restrict__ sa,
                                                             double* __restrict__ sb, double r2[4][4])
    {
        const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) r2[i][j] = 0.0;
        for (int d0 = 0; d0 < D; d0 += DCHUNK)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int e = tid + r * 256, dd = e & (DCHUNK - 1), p = e >> 4;
                const int d = d0 + dd;
                double    va = 0.0, vb = 0.0;
                if (d < D)
                {
                    const double s = inv_l[d];
                    if (baseA + p < nA) va = PA[(size_t) d + (size_t) (baseA + p) * ldA] * s;
                    if (baseB + p < nB) vb = PB[(size_t) d + (size_t) (baseB + p) * ldB] * s;
                }
                sa[dd * GPAD + p] = va;
                sb[dd * GPAD + p] = vb;
            }
            __syncthreads();
            const int dmax = min(DCHUNK, D - d0);
#pragma unroll 4
            for (int dd = 0; dd < dmax; ++dd)
            {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sa[dd * GPAD + tx + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = sb[dd * GPAD + ty * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const double df = a[i] - b[j];
                        r2[i][j]        = fma(df, df, r2[i][j]);
                    }
            }
            __syncthreads();
        }
    }

    // kernel value with the non-positive exp above (same formulas as kernel_value in common.cuh)
    template <int KT> __device__ __forceinline__ double gram_kernel_value(double a, double r2)
    {
        if (KT == 0) return a * exp_nonpositive(-0.5 * r2);
        const double s = sqrt(5.0 * r2);
        return a * (1.0 + s + (5.0 / 3.0) * r2) * exp_nonpositive(-s);
    }

    // K_y = k(X, X) + noise I, CalcLargeKY (src/regressor.cpp:61-89): lower tiles computed, mirrored through shared memory.
    // Rows / columns >= N (padding up to ld) get the identity. grid = nt (nt + 1) / 2 tiles, 256 threads, 4 CTAs per SM (the
    // 528 tiles of N = 2048 are then ONE wave on 148 SMs).
    template <int KT>
    __global__ void __launch_bounds__(256, 4)
        gram_sym_kernel(const double* __restrict__ X, int N, int D, int ld, const double* __restrict__ theta,
                        const double* __restrict__ inv_l, double noise, double* __restrict__ out)
    {
        __shared__ double smem[TILE * GPAD];
        double*           sa = smem;
        double*           sb = smem + DCHUNK * GPAD;
        int               tm, tn;
        {
            const int t = blockIdx.x;
            tm          = (int) ((sqrtf(8.0f * (float) t + 1.0f) - 1.0f) * 0.5f);
            while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
            while (tm * (tm + 1) / 2 > t) --tm;
            tn = t - tm * (tm + 1) / 2;
        }
        const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

        double r2[4][4];
        tile_sq_dist_interleaved(X, D, N, tm * TILE, X, D, N, tn * TILE, D, inv_l, sa, sb, r2);

        const double a = theta[0];
        double       v[4][4];
        if (tm != tn && (tm + 1) * TILE <= N) // interior tile: no diagonal, no padding (tn < tm, so its columns are inside too)
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i][j] = gram_kernel_value<KT>(a, r2[i][j]);
        }
        else
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const int    gi = tm * TILE + tx + 16 * i, gj = tn * TILE + ty * 4 + j;
                    const double k  = gram_kernel_value<KT>(a, r2[i][j]);
                    v[i][j]         = (gi < N && gj < N) ? k + (gi == gj ? noise : 0.0) : (gi == gj ? 1.0 : 0.0);
                }
        }
        double* o = out + (size_t) (tm * TILE + tx) + (size_t) (tn * TILE + ty * 4) * ld;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[(size_t) (16 * i) + (size_t) j * ld] = v[i][j];
        if (tm != tn)
        {
            // the staging rows are dead (tile_sq_dist_interleaved ends with a barrier): reuse them for the transpose
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) smem[(tx + 16 * i) * GPAD + ty * 4 + j] = v[i][j];
            __syncthreads();
            double* om = out + (size_t) (tn * TILE) + (size_t) (tm * TILE) * ld;
#pragma unroll 4
            for (int e = tid; e < TILE * TILE; e += 256)
            {
                const int r = e & 63, c = e >> 6; // element (tn*64 + r, tm*64 + c) = tile(c, r)
                om[(size_t) r + (size_t) c * ld] = smem[c * GPAD + r];
            }
        }
    }

    // linear index over lower-triangular tiles -> (tm, tn), tn <= tm
    __device__ __forceinline__ void lower_tile(int t, int& tm, int& tn)
    {
        tm = (int) ((sqrt(8.0 * (double) t + 1.0) - 1.0) * 0.5);
        while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
        while (tm * (tm + 1) / 2 > t) --tm;
        tn = t - tm * (tm + 1) / 2;
    }

    // MODE 0: K_y = k(X, X) + noise I                     CalcLargeKY (src/regressor.cpp:61-89)
    // MODE 1: Wm  = (alpha_i alpha_j - Kinv_ij) * kl(r2)  weights of the theta-gradient trace terms
    //         (CalcObjectiveThetaDerivative src/preference-regressor.cpp:77-115 with the per-pair
    //          dk/dl_t = kl * d_t^2 / l_t^3 of kernel-functions.cpp:22-50, 114-142)
    // Rows / columns >= N (padding up to ld) get the identity (MODE 0) or zero (MODE 1).
    template <int MODE>
    __global__ void __launch_bounds__(256)
        gram_tile_kernel(const double* __restrict__ X, int N, int D, int ld, const double* __restrict__ theta,
                         const double* __restrict__ inv_l, double noise, int kernel_type, double* __restrict__ out,
                         const double* __restrict__ Kinv, const double* __restrict__ alpha)
    {
        // one buffer: the two input stages first, the transposing store of the mirrored tile afterwards
        __shared__ double smem[TILE * (TILE + 1)];
        double(*sa)[TILE]     = reinterpret_cast<double(*)[TILE]>(smem);
        double(*sb)[TILE]     = sa + DCHUNK;
        double(*st)[TILE + 1] = reinterpret_cast<double(*)[TILE + 1]>(smem);
        int tm, tn;
        lower_tile(blockIdx.x, tm, tn);
        const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

        double r2[4][4];
        tile_sq_dist(X, D, N, tm * TILE, X, D, N, tn * TILE, D, inv_l, sa, sb, r2);

        const double a = theta[0];
        double       v[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const int gi = tm * TILE + tx * 4 + i, gj = tn * TILE + ty * 4 + j;
                double    val;
                if (MODE == 0)
                {
                    if (gi < N && gj < N)
                        val = kernel_value(kernel_type, a, r2[i][j]) + (gi == gj ? noise : 0.0);
                    else
                        val = (gi == gj) ? 1.0 : 0.0;
                }
                else
                {
                    if (gi < N && gj < N)
                    {
                        double ka, kl;
                        kernel_theta_weights(kernel_type, a, r2[i][j], ka, kl);
                        val = (alpha[gi] * alpha[gj] - Kinv[(size_t) gi + (size_t) gj * ld]) * kl;
                    }
                    else
                        val = 0.0;
                }
                v[i][j] = val;
            }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                out[(size_t) (tm * TILE + tx * 4 + i) + (size_t) (tn * TILE + ty * 4 + j) * ld] = v[i][j];
        if (tm != tn)
        {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) st[tx * 4 + i][ty * 4 + j] = v[i][j];
            __syncthreads();
            for (int e = tid; e < TILE * TILE; e += 256)
            {
                const int r = e & 63, c = e >> 6; // (tn*64 + r, tm*64 + c) <- tile(c, r)
                out[(size_t) (tn * TILE + r) + (size_t) (tm * TILE + c) * ld] = st[c][r];
            }
        }
    }
} // namespace slsgp
