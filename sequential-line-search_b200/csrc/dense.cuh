// FP64 dense building blocks: tiled GEMM with triangular k-range pruning (DMMA and DFMA forms) and small matrix
// utilities. Everything works on column-major storage whose leading
// dimension is a multiple of TILE (=64); matrices are padded with an identity block so that no kernel needs
// ragged-edge handling.
#pragma once

#include "common.cuh"

namespace slsgp
{
    struct GemmArgs
    {
        const double* A;
        const double* B;
        double*       C;
        int           m, n, k; // m, n multiples of 64; k multiple of 16
        int           lda, ldb, ldc;
        double        alpha, beta;
        long long     sA, sB, sC; // batch strides (elements), batch index = blockIdx.z
        int           lower_only; // 1: skip tiles strictly above the diagonal (tn > tm)
        int           k_lo_mode;  // 0: 0 | 1: tn*64 (B lower-triangular in (k, n)) | 2: max(tm, tn)*64
        int           k_hi_mode;  // 0: k | 1: (tm+1)*64 (A lower-triangular in (m, k))
        int           row0, row_step, row_limit; // tile exists iff row0 + z*row_step + tm*64 < row_limit
    };

    inline GemmArgs gemm_args(const double* A, const double* B, double* C, int m, int n, int k, int lda, int ldb,
                              int ldc, double alpha, double beta)
    {
        GemmArgs g;
        g.A = A, g.B = B, g.C = C, g.m = m, g.n = n, g.k = k, g.lda = lda, g.ldb = ldb, g.ldc = ldc;
        g.alpha = alpha, g.beta = beta, g.sA = g.sB = g.sC = 0;
        g.lower_only = 0, g.k_lo_mode = 0, g.k_hi_mode = 0, g.row0 = 0, g.row_step = 0, g.row_limit = 1 << 30;
        return g;
    }

    // C = alpha * op(A) * op(B) + beta * C on 64 x 64 output tiles, 256 threads, 4 x 4 register micro-tile.
    // op(A) is m x k: TA == false -> A[m + k*lda], TA == true -> A[k + m*lda]. op(B) is k x n likewise.
    template <bool TA, bool TB> __global__ void __launch_bounds__(256) gemm64_kernel(const GemmArgs g)
    {
        const int tm = blockIdx.x, tn = blockIdx.y, z = blockIdx.z;
        if (g.lower_only && tn > tm) return;
        if (g.row0 + z * g.row_step + tm * TILE >= g.row_limit) return;

        const double* __restrict__ A = g.A + z * g.sA;
        const double* __restrict__ B = g.B + z * g.sB;
        double* C                    = g.C + z * g.sC; // may alias A (in-place panel update)

        int k_lo = 0, k_hi = g.k;
        if (g.k_lo_mode == 1) k_lo = tn * TILE;
        if (g.k_lo_mode == 2) k_lo = max(tm, tn) * TILE;
        if (g.k_hi_mode == 1) k_hi = min(g.k, (tm + 1) * TILE);

        __shared__ double As[16][TILE + 1];
        __shared__ double Bs[16][TILE + 1];

        const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
        double    acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

        for (int k0 = k_lo; k0 < k_hi; k0 += 16)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int e = tid + r * 256;
                if (!TA)
                {
                    const int mm = e & 63, kk = e >> 6;
                    As[kk][mm]   = A[(size_t) (tm * TILE + mm) + (size_t) (k0 + kk) * g.lda];
                }
                else
                {
                    const int kk = e & 15, mm = e >> 4;
                    As[kk][mm]   = A[(size_t) (k0 + kk) + (size_t) (tm * TILE + mm) * g.lda];
                }
                if (!TB)
                {
                    const int kk = e & 15, nn = e >> 4;
                    Bs[kk][nn]   = B[(size_t) (k0 + kk) + (size_t) (tn * TILE + nn) * g.ldb];
                }
                else
                {
                    const int nn = e & 63, kk = e >> 6;
                    Bs[kk][nn]   = B[(size_t) (tn * TILE + nn) + (size_t) (k0 + kk) * g.ldb];
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk)
            {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[kk][tx * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }

#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const size_t idx = (size_t) (tm * TILE + tx * 4 + i) + (size_t) (tn * TILE + ty * 4 + j) * g.ldc;
                double       v   = g.alpha * acc[i][j];
                if (g.beta != 0.0) v += g.beta * C[idx];
                C[idx] = v;
            }
    }

    // D (8x8) += A (8x4, row) * B (4x8, col) in IEEE double on the FP64 tensor pipe (DMMA). Per lane: a = A[lane/4][lane%4],
    // b = B[lane%4][lane/4], c0/c1 = C[lane/4][2*(lane%4) + {0,1}].
    __device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b)
    {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }

    // Same contract as gemm64_kernel (64 x 64 output tile per CTA, the triangular k-range pruning, batching, in-place
    // panel updates), with the inner product on the FP64 tensor pipe: 8 warps, warp = 32 (m) x 16 (n) = 4 x 2 DMMA tiles,
    // 16-deep double-buffered shared-memory stages (one barrier per stage), the next stage's global loads issued into
    // registers before the current stage is multiplied and parked in the other buffer after it.
    constexpr int DMMA_LDS = TILE + 4, DMMA_LDK = 16 + 4;
    template <bool TA, bool TB> __global__ void __launch_bounds__(256) gemm64_dmma_kernel(const GemmArgs g)
    {
        const int tm = blockIdx.x, tn = blockIdx.y, z = blockIdx.z;
        if (g.lower_only && tn > tm) return;
        if (g.row0 + z * g.row_step + tm * TILE >= g.row_limit) return;

        const double* __restrict__ A = g.A + z * g.sA;
        const double* __restrict__ B = g.B + z * g.sB;
        double* C                    = g.C + z * g.sC; // may alias A (in-place panel update)

        int k_lo = 0, k_hi = g.k;
        if (g.k_lo_mode == 1) k_lo = tn * TILE;
        if (g.k_lo_mode == 2) k_lo = max(tm, tn) * TILE;
        if (g.k_hi_mode == 1) k_hi = min(g.k, (tm + 1) * TILE);

        // Two stages per operand. An operand that arrives k-contiguous from global memory (A^T, or B as stored) is kept
        // [row][k] with a row stride of 20 doubles, one that arrives row-contiguous [k][row] with a stride of 68: in both
        // forms the 16 lanes of a half-warp store, and later read their fragments from, 16 distinct 8-byte banks.
        constexpr int ST_KM = 16 * DMMA_LDS, ST_MK = TILE * DMMA_LDK, ST = ST_MK > ST_KM ? ST_MK : ST_KM;
        __shared__ double As[2][ST];
        __shared__ double Bs[2][ST];

        const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
        const int wm = (warp & 1) * 32, wn = (warp >> 1) * 16; // this warp's corner inside the tile
        const int lr = lane >> 2, lc = lane & 3;
        double    acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        double ra[4], rb[4];
        const auto fetch = [&](int k0) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int e = tid + r * 256;
                if (!TA)
                    ra[r] = A[(size_t) (tm * TILE + (e & 63)) + (size_t) (k0 + (e >> 6)) * g.lda];
                else
                    ra[r] = A[(size_t) (k0 + (e & 15)) + (size_t) (tm * TILE + (e >> 4)) * g.lda];
                if (!TB)
                    rb[r] = B[(size_t) (k0 + (e & 15)) + (size_t) (tn * TILE + (e >> 4)) * g.ldb];
                else
                    rb[r] = B[(size_t) (tn * TILE + (e & 63)) + (size_t) (k0 + (e >> 6)) * g.ldb];
            }
        };
        const auto stash = [&](int buf) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
            {
                const int e = tid + r * 256;
                if (!TA)
                    As[buf][(e >> 6) * DMMA_LDS + (e & 63)] = ra[r];
                else
                    As[buf][(e >> 4) * DMMA_LDK + (e & 15)] = ra[r];
                if (!TB)
                    Bs[buf][(e >> 4) * DMMA_LDK + (e & 15)] = rb[r];
                else
                    Bs[buf][(e >> 6) * DMMA_LDS + (e & 63)] = rb[r];
            }
        };

        int buf = 0;
        if (k_lo < k_hi)
        {
            fetch(k_lo);
            stash(0);
        }
        __syncthreads();
        for (int k0 = k_lo; k0 < k_hi; k0 += 16)
        {
            const bool more = k0 + 16 < k_hi;
            if (more) fetch(k0 + 16); // in flight while this stage is multiplied
            const double* as = As[buf];
            const double* bs = Bs[buf];
#pragma unroll
            for (int ks = 0; ks < 16; ks += 4)
            {
                double a[4], b[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = TA ? as[(wm + i * 8 + lr) * DMMA_LDK + ks + lc] : as[(ks + lc) * DMMA_LDS + wm + i * 8 + lr];
#pragma unroll
                for (int j = 0; j < 2; ++j) b[j] = TB ? bs[(ks + lc) * DMMA_LDS + wn + j * 8 + lr] : bs[(wn + j * 8 + lr) * DMMA_LDK + ks + lc];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            if (more) stash(buf ^ 1); // the stage the previous iteration read; every warp is past that iteration's barrier
            __syncthreads();
            buf ^= 1;
        }

#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                {
                    const int    m = tm * TILE + wm + i * 8 + lr, n = tn * TILE + wn + j * 8 + lc * 2 + h;
                    const size_t idx = (size_t) m + (size_t) n * g.ldc;
                    double       v   = g.alpha * acc[i][j][h];
                    if (g.beta != 0.0) v += g.beta * C[idx];
                    C[idx] = v;
                }
    }

    // Mirror the lower triangle into the upper one (64 x 64 tiles through shared memory so both sides coalesce).
    __global__ void __launch_bounds__(256) symmetrize_kernel(double* __restrict__ A, int lda)
    {
        const int tm = blockIdx.x, tn = blockIdx.y;
        if (tn >= tm) return;
        __shared__ double S[TILE][TILE + 1];
        for (int e = threadIdx.x; e < TILE * TILE; e += 256)
        {
            const int r = e & 63, c = e >> 6;
            S[r][c]     = A[(size_t) (tm * TILE + r) + (size_t) (tn * TILE + c) * lda];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < TILE * TILE; e += 256)
        {
            const int r = e & 63, c = e >> 6; // element (tn*64 + r, tm*64 + c) = lower (tm*64 + c, tn*64 + r)
            A[(size_t) (tn * TILE + r) + (size_t) (tm * TILE + c) * lda] = S[c][r];
        }
    }
    // out[0] = 2 * sum_{i<n} log(L_ii)   (mathtoolbox log-determinant.cpp:8-11); single block, fixed order.
    __global__ void __launch_bounds__(256) logdet_kernel(const double* __restrict__ L, int n, int ld,
                                                         double* __restrict__ out)
    {
        __shared__ double part[256];
        double            s = 0.0;
        for (int i = threadIdx.x; i < n; i += 256) s += log(L[(size_t) i + (size_t) i * ld]);
        part[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[0] = 2.0 * part[0];
    }

    // y = op(A) x for an n x n matrix, one warp per output element (deterministic shuffle reduction).
    // TRANS == true:  y_i = sum_j A[j + i*lda] x_j (column dot: coalesced)  -> y = A^T x
    // TRANS == false: y_i = sum_j A[i + j*lda] x_j (row dot: strided; used only for triangular W y)
    // tri: 0 full | 1 A lower-triangular (row i uses j <= i; column i uses j >= i)
    template <bool TRANS>
    __global__ void __launch_bounds__(256) gemv_kernel(const double* __restrict__ A, int n, int lda,
                                                       const double* __restrict__ x, double* __restrict__ y, int tri)
    {
        const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (warp >= n) return;
        const int i = warp;
        double    s = 0.0;
        if (TRANS)
        {
            const int j0 = tri ? i : 0;
            for (int j = j0 + lane; j < n; j += 32) s = fma(A[(size_t) j + (size_t) i * lda], x[j], s);
        }
        else
        {
            const int j1 = tri ? i + 1 : n;
            for (int j = lane; j < j1; j += 32) s = fma(A[(size_t) i + (size_t) j * lda], x[j], s);
        }
        s = warp_sum(s);
        if (lane == 0) y[i] = s;
    }
} // namespace slsgp
