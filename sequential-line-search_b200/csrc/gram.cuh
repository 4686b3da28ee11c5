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
