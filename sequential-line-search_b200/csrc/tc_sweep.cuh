// K4 (tensor-core path, SLSGP_SWEEP_TENSOR): the acquisition sweep with the N x N contraction on tcgen05.
//
// Same algebra as sweep.cuh (reference: src/preference-regressor.cpp:293-330, src/regressor.cpp:45-59,91-108), for the
// ARD squared-exponential kernel, where the x-gradient weight is g_i = -c k_i (c = 2 in the reference,
// external/mathtoolbox/src/kernel-functions.cpp:92):
//   u  = K^-1 k                         (N x N x M contraction: tcgen05.mma, fp16 operands, fp32 accumulation in TMEM)
//   q  = sum_i k_i u_i                  sigma^2 = a - q
//   P2_d = -c sum_i X_di k_i u_i        (epilogue of the same kernel, fp32 FMA on the accumulator tile)
//   mu = sum_j k_j alpha_j,  P1_d = -c sum_j X_dj k_j alpha_j
//                                       (extra B columns alpha_j * [1, X_dj], split into fp16 hi + lo, same MMA stream)
// Operands in HBM (fp16, K-major = contiguous along the contraction index j):
//   Ks   [Mpad x ldt]        Ks[m][j]   = sK * k(x_m, X_j)                 written by kstar16_kernel per shard
//   Bmat [(ldt+256) x ldt]   Bmat[i][j] = sA * Kinv[i][j]       (i < N)     rows ldt + 2c / 2c+1: extras hi / lo
// and, for the epilogue, Xt [ldt x XP] fp32 with Xt[i] = (1, X_0i .. X_{D-1}i, 0 ..).
// One CTA owns 128 candidates (TMEM lanes) and walks the ldt/256 column blocks of Kinv plus the extras block; the
// accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of block b overlaps the MMAs of b+1.
#pragma once

#include "common.cuh"
#include "tc.cuh"

namespace slsgp
{
    constexpr int TC_BM          = 128; // candidates per CTA tile (UMMA M)
    constexpr int TC_BN          = 256; // Kinv columns per accumulator (UMMA N)
    constexpr int TC_BK          = 64;  // contraction elements per pipeline stage (128 bytes of fp16 = one swizzle row)
    constexpr int TC_A_BYTES     = TC_BM * TC_BK * 2;
    constexpr int TC_B_BYTES     = TC_BN * TC_BK * 2;
    constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
    constexpr int TC_MAX_STAGES  = 8;
    constexpr int TC_THREADS     = 256; // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..7 epilogue

    struct TcScales
    {
        float sK, sA, sE; // power-of-two multipliers putting k, Kinv and the extras into fp16 range
        float c0;         // log2(a * sK): k16 = exp2(-0.5 log2(e) r2 + c0)
        float inv_u;      // 1 / (sA sK^2): accumulator * k16 -> k_i u_i
        float inv_e;      // 1 / (sK sE)
    };

    __device__ __forceinline__ float pow2_floor_scale(float target, float maxabs)
    {
        if (!(maxabs > 0.f)) return 1.f;
        return exp2f(floorf(log2f(target / maxabs)));
    }

    // One block. |Kinv_ij| <= max_i Kinv_ii for an SPD matrix, so the diagonal bounds the whole operand.
    __global__ void __launch_bounds__(256)
        tc_scales_kernel(const double* __restrict__ Kinv, int ld, int N, const double* __restrict__ alpha,
                         const double* __restrict__ X, int D, const double* __restrict__ theta, TcScales* __restrict__ out)
    {
        __shared__ double sm[3][256];
        double            md = 0.0, ma = 0.0, mx = 1.0;
        for (int i = threadIdx.x; i < N; i += 256)
        {
            md = fmax(md, fabs(Kinv[(size_t) i + (size_t) i * ld]));
            ma = fmax(ma, fabs(alpha[i]));
            for (int d = 0; d < D; ++d) mx = fmax(mx, fabs(X[(size_t) d + (size_t) i * D]));
        }
        sm[0][threadIdx.x] = md, sm[1][threadIdx.x] = ma, sm[2][threadIdx.x] = mx;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o)
                for (int r = 0; r < 3; ++r) sm[r][threadIdx.x] = fmax(sm[r][threadIdx.x], sm[r][threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0)
        {
            const float a = (float) theta[0];
            TcScales    s;
            s.sK    = pow2_floor_scale(32768.f, a);
            s.sA    = pow2_floor_scale(32768.f, (float) sm[0][0]);
            s.sE    = pow2_floor_scale(32768.f, (float) (sm[1][0] * sm[2][0]));
            s.c0    = log2f(a * s.sK);
            s.inv_u = 1.f / (s.sA * s.sK * s.sK);
            s.inv_e = 1.f / (s.sK * s.sE);
            *out    = s;
        }
    }

    // Bmat rows (see header). grid: (ldt / 256, 2 * ldt + 256), 256 threads: one thread per element, j fastest.
    //   [0, ldt)               hi part of sA * Kinv
    //   [ldt, ldt + 256)       extras: row 2c = hi, 2c + 1 = lo of sE * alpha_j * (1, X_0j, ..)[c]
    //   [ldt + 256, 2ldt+256)  lo part of sA * Kinv (the fp16 rounding residual of the hi part)
    __global__ void __launch_bounds__(256)
        tc_pack_b_kernel(const double* __restrict__ Kinv, int ld, int N, int D, int XP, int ldt,
                         const double* __restrict__ alpha, const double* __restrict__ X,
                         const TcScales* __restrict__ sc, __half* __restrict__ Bmat)
    {
        const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
        float     v = 0.f;
        if (i < ldt || i >= ldt + TC_BN)
        {
            const int  r  = i < ldt ? i : i - ldt - TC_BN;
            if (r < N && j < N)
            {
                const float full = (float) (Kinv[(size_t) j + (size_t) r * ld] * (double) sc->sA); // symmetric
                const float hi   = __half2float(__float2half_rn(full));
                v                = i < ldt ? hi : full - hi;
            }
        }
        else if (j < N)
        {
            const int e = i - ldt, c = e >> 1;
            if (c <= D)
            {
                const double xhat = c == 0 ? 1.0 : X[(size_t) (c - 1) + (size_t) j * D] - 0.5; // centred, as Xt
                const float  full = (float) (alpha[j] * xhat * (double) sc->sE);
                const float  hi   = __half2float(__float2half_rn(full));
                v                 = (e & 1) ? full - hi : hi;
            }
        }
        Bmat[(size_t) i * ldt + j] = __float2half_rn(v);
    }

    // Xt[i][c] = (1, X_0i, .., X_{D-1}i, 0..) in fp32 for the epilogue; Xs32[d][j] = -(X_dj - 1/2) / l_d (negated: kstar16
    // forms q - x with a packed add).
    __global__ void __launch_bounds__(256)
        tc_pack_x_kernel(const double* __restrict__ X, int N, int D, int XP, int ldt, const double* __restrict__ inv_l, int n_ones,
                         float* __restrict__ Xt, float* __restrict__ Xs32)
    {
        // Xt columns: n_ones ones (SE: 1, the q column; Matern: 2, q and gb = sum g_i u_i), then X_0i - 1/2 .. X_{D-1}i - 1/2, then
        // zeros. Centring on the box centre halves |X| and with it the fp32 cancellation error of (x_d gb - P2_d) near data points;
        // sweep_finish_kernel applies the same shift to the candidate (x_shift).
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i >= ldt) return;
        for (int c = 0; c < XP; ++c)
        {
            float v = 0.f;
            if (i < N) v = c < n_ones ? 1.f : (c < n_ones + D ? (float) (X[(size_t) (c - n_ones) + (size_t) i * D] - 0.5) : 0.f);
            Xt[(size_t) i * XP + c] = v;
        }
        for (int d = 0; d < D; ++d)
            Xs32[(size_t) d * ldt + i] = i < N ? -(float) ((X[(size_t) d + (size_t) i * D] - 0.5) * inv_l[d]) : 0.f;
    }

    // Ks[m][j] = fp16(sK * a * exp(-r2/2)) and its rounding residual; r2 by direct differences of the centred,
    // length-scaled coordinates in fp32 (no |q|^2 + |x|^2 - 2 q.x cancellation: near a data point sigma^2 = a - k.A.k
    // amplifies any relative error of k by a / sigma^2). The differences and their squares run on the packed FP32
    // pipe (FADD2 + FFMA2: two observations per issue slot); Xs32 holds the NEGATED observation coordinates.
    // Tile: 64 candidates x 128 observations per CTA. grid: (ldt / 128, Mpad / 64); dynamic smem: (64 * DQ + D * 128) floats,
    // DQ = round_up(D, 4).
    // Register budget: 64 per thread (4 CTAs per SM by __launch_bounds__). The contraction kernel below keeps one persistent
    // 256-thread CTA of 188 registers on every SM, which leaves 16384 registers = exactly one 64-register CTA of this kernel
    // (and ~14 KB of shared memory): with that the generator of shard s + 1 runs UNDER the contraction of shard s instead of
    // in front of it (run_sweep puts it on the `pre` stream). At 71 registers it could not co-reside and the overlap was void.
    template <int KT>
    __global__ void __launch_bounds__(256, 4)
        kstar16_kernel(const double* __restrict__ Xq, long long Mc, int D, int N, int ldt,
                       const float* __restrict__ Xs32, const double* __restrict__ inv_l,
                       const TcScales* __restrict__ sc, __half* __restrict__ Ks, __half* __restrict__ Ks_lo,
                       __half* __restrict__ Gs, __half* __restrict__ Gs_lo)
    {
        constexpr int kernel_type = KT;
        extern __shared__ __align__(16) float ksm[];
        const int               DQ = (D + 3) & ~3; // row stride of sq: 16-byte aligned rows
        float*                  sq = ksm;          // [64][DQ]
        float*                  sx = ksm + 64 * DQ; // [D][128], negated
        const int               tid = threadIdx.x, j_base = blockIdx.x * 128;
        const long long         m_base = (long long) blockIdx.y * 64;
        for (int e = tid; e < 64 * DQ; e += 256)
        {
            const int       p = e / DQ, d = e - p * DQ;
            const long long m = m_base + p;
            sq[p * DQ + d]  = (m < Mc && d < D) ? (float) ((Xq[(size_t) d + (size_t) m * D] - 0.5) * inv_l[d]) : 0.f;
        }
        for (int e = tid; e < D * 128; e += 256) sx[e] = Xs32[(size_t) (e >> 7) * ldt + j_base + (e & 127)];
        __syncthreads();

        // thread = 4 candidates (tm) x 8 observations: j = 4 tj + {0..3} and 64 + 4 tj + {0..3}, so the 16-byte operand reads
        // of consecutive threads are consecutive in shared memory (conflict-free) and each row is written as two 8-byte
        // pieces that coalesce into full 128-byte lines across the half-warp
        const int tj = tid & 15, tm = tid >> 4;
        float2    r2[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) r2[i][jj] = make_float2(0.f, 0.f);
        for (int d0 = 0; d0 < DQ; d0 += 4)
        {
            float4 qv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&sq[(tm * 4 + i) * DQ + d0]);
#pragma unroll
            for (int dd = 0; dd < 4; ++dd)
            {
                if (d0 + dd < D)
                {
                    const float4 xa = *reinterpret_cast<const float4*>(&sx[(d0 + dd) * 128 + tj * 4]);
                    const float4 xb = *reinterpret_cast<const float4*>(&sx[(d0 + dd) * 128 + 64 + tj * 4]);
                    const float2 x2[4] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w), make_float2(xb.x, xb.y), make_float2(xb.z, xb.w)};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                    {
                        const float  q  = dd == 0 ? qv[i].x : (dd == 1 ? qv[i].y : (dd == 2 ? qv[i].z : qv[i].w));
                        const float2 q2 = make_float2(q, q);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                        {
                            const float2 df = tc::fadd2(q2, x2[jj]); // q - x
                            r2[i][jj]       = tc::ffma2(df, df, r2[i][jj]);
                        }
                    }
                }
            }
        }
        const float c1 = -0.72134752044448170368f; // -0.5 * log2(e)
        const float c0 = sc->c0;
        const bool  interior = m_base + 64 <= Mc && j_base + 128 <= N; // whole tile in range: no per-value predicates
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const long long m = m_base + tm * 4 + i;
            __half2         h[4], hl[4], gh[4], gl[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
            {
                const int j = j_base + (jj >> 1) * 64 + tj * 4 + (jj & 1) * 2;
                float     v0, v1, g0 = 0.f, g1 = 0.f;
                if (kernel_type == 0)
                    v0 = tc::ex2_approx(fmaf(r2[i][jj].x, c1, c0)), v1 = tc::ex2_approx(fmaf(r2[i][jj].y, c1, c0));
                else
                {
                    // Matern 5/2 (kernel-functions.cpp:95-112, :179-212): s = sqrt(5) r, k = a (1 + s + s^2/3) e^-s and the
                    // x-gradient weight g = -(5/3) a (1 + s) e^-s (so that dk/dx_d = g (x_d - X_d) / l_d^2); both scaled by sK
                    const float s0 = sqrtf(5.f * r2[i][jj].x), s1 = sqrtf(5.f * r2[i][jj].y);
                    const float e0 = tc::ex2_approx(fmaf(s0, -1.44269504088896340736f, c0));
                    const float e1 = tc::ex2_approx(fmaf(s1, -1.44269504088896340736f, c0));
                    v0 = e0 * fmaf(s0, fmaf(s0, 0.33333333333333333f, 1.f), 1.f), v1 = e1 * fmaf(s1, fmaf(s1, 0.33333333333333333f, 1.f), 1.f);
                    g0 = -1.66666666666666667f * e0 * (1.f + s0), g1 = -1.66666666666666667f * e1 * (1.f + s1);
                }
                if (!interior)
                {
                    if (!(m < Mc && j < N)) v0 = 0.f, g0 = 0.f;
                    if (!(m < Mc && j + 1 < N)) v1 = 0.f, g1 = 0.f;
                }
                h[jj]          = __floats2half2_rn(v0, v1);
                const float2 b = __half22float2(h[jj]);
                hl[jj]         = __floats2half2_rn(v0 - b.x, v1 - b.y); // rounding residual (second fp16 term)
                if (kernel_type != 0)
                {
                    gh[jj]         = __floats2half2_rn(g0, g1);
                    const float2 c = __half22float2(gh[jj]);
                    gl[jj]         = __floats2half2_rn(g0 - c.x, g1 - c.y);
                }
            }
            const size_t off = (size_t) m * ldt + j_base + tj * 4;
            *reinterpret_cast<uint2*>(Ks + off)      = *reinterpret_cast<const uint2*>(&h[0]);
            *reinterpret_cast<uint2*>(Ks + off + 64) = *reinterpret_cast<const uint2*>(&h[2]);
            if (Ks_lo)
            {
                *reinterpret_cast<uint2*>(Ks_lo + off)      = *reinterpret_cast<const uint2*>(&hl[0]);
                *reinterpret_cast<uint2*>(Ks_lo + off + 64) = *reinterpret_cast<const uint2*>(&hl[2]);
            }
            if (kernel_type != 0)
            {
                *reinterpret_cast<uint2*>(Gs + off)      = *reinterpret_cast<const uint2*>(&gh[0]);
                *reinterpret_cast<uint2*>(Gs + off + 64) = *reinterpret_cast<const uint2*>(&gh[2]);
                if (Gs_lo)
                {
                    *reinterpret_cast<uint2*>(Gs_lo + off)      = *reinterpret_cast<const uint2*>(&gl[0]);
                    *reinterpret_cast<uint2*>(Gs_lo + off + 64) = *reinterpret_cast<const uint2*>(&gl[2]);
                }
            }
        }
    }

    // ---- k* generator, persistent form ---------------------------------------------------------------------------------
    // Same arithmetic and output as kstar16_kernel. A CTA takes whole strips of 64 candidates x ALL observations and walks them
    // in 64-observation chunks whose length-scaled coordinates (Xs32 rows) arrive through a double-buffered cp.async pipeline,
    // so one resident CTA per SM keeps its eight warps busy: global latency is paid once per strip, not once per 8192 values.
    // That is the form that can live UNDER the contraction kernel of the previous shard: the contraction keeps one persistent
    // 188-register CTA and ~213 KB of shared memory on every SM, which leaves room for exactly one 256-thread CTA of <= 64
    // registers and <= 13 KB; launched with one CTA per SM both kernels are fully resident whatever order the block scheduler
    // places them in. Shared memory: sq [64][DQ] + sx [2][D][64] floats (12 KB at D = 16).
    // thread = 4 candidates (tm = tid >> 4) x 4 consecutive observations (tj = tid & 15) per chunk.
    __device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src)
    {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gmem_src) : "memory");
    }
    __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

    template <int KT>
    __global__ void __launch_bounds__(256, 4)
        kstar16_strip_kernel(const double* __restrict__ Xq, long long Mc, int D, int N, int ldt, const float* __restrict__ Xs32,
                             const double* __restrict__ inv_l, const TcScales* __restrict__ sc, __half* __restrict__ Ks,
                             __half* __restrict__ Ks_lo, __half* __restrict__ Gs, __half* __restrict__ Gs_lo, int n_strips)
    {
        extern __shared__ __align__(16) float ksm[];
        const int DQ = (D + 3) & ~3;
        float*    sq = ksm;           // [64][DQ]
        float*    sx = ksm + 64 * DQ; // [2][D][64], negated observation coordinates
        const int tid = threadIdx.x, tj = tid & 15, tm = tid >> 4, n_chunks = ldt / 64;
        const float c1 = -0.72134752044448170368f; // -0.5 * log2(e)
        const float c0 = sc->c0;

        auto prefetch = [&](int chunk, int buf) {
            // D rows of 64 floats = 16 sixteen-byte pieces per row
            for (int e = tid; e < D * 16; e += 256)
                cp_async_16(sx + buf * D * 64 + (e >> 4) * 64 + (e & 15) * 4, Xs32 + (size_t) (e >> 4) * ldt + chunk * 64 + (e & 15) * 4);
            cp_async_commit();
        };

        for (int strip = blockIdx.x; strip < n_strips; strip += gridDim.x)
        {
            const long long m_base = (long long) strip * 64;
            __syncthreads(); // the previous strip's readers of sq / sx are done
            prefetch(0, 0);
            for (int e = tid; e < 64 * DQ; e += 256)
            {
                const int       p = e / DQ, d = e - p * DQ;
                const long long m = m_base + p;
                sq[p * DQ + d]  = (m < Mc && d < D) ? (float) ((Xq[(size_t) d + (size_t) m * D] - 0.5) * inv_l[d]) : 0.f;
            }
            for (int chunk = 0; chunk < n_chunks; ++chunk)
            {
                const int buf = chunk & 1, j_base = chunk * 64;
                if (chunk + 1 < n_chunks)
                {
                    prefetch(chunk + 1, buf ^ 1);
                    cp_async_wait<1>();
                }
                else
                    cp_async_wait<0>();
                __syncthreads();
                const float* sxb = sx + buf * D * 64;
                float2       r2[4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i) r2[i][0] = r2[i][1] = make_float2(0.f, 0.f);
                for (int d0 = 0; d0 < DQ; d0 += 4)
                {
                    float4 qv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&sq[(tm * 4 + i) * DQ + d0]);
#pragma unroll
                    for (int dd = 0; dd < 4; ++dd)
                    {
                        if (d0 + dd < D)
                        {
                            const float4 xa    = *reinterpret_cast<const float4*>(&sxb[(d0 + dd) * 64 + tj * 4]);
                            const float2 x2[2] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w)};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                            {
                                const float  q  = dd == 0 ? qv[i].x : (dd == 1 ? qv[i].y : (dd == 2 ? qv[i].z : qv[i].w));
                                const float2 q2 = make_float2(q, q);
#pragma unroll
                                for (int jj = 0; jj < 2; ++jj)
                                {
                                    const float2 df = tc::fadd2(q2, x2[jj]); // q - x
                                    r2[i][jj]       = tc::ffma2(df, df, r2[i][jj]);
                                }
                            }
                        }
                    }
                }
                const bool interior = m_base + 64 <= Mc && j_base + 64 <= N;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                {
                    const long long m = m_base + tm * 4 + i;
                    __half2         h[2], hl[2], gh[2], gl[2];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj)
                    {
                        const int j = j_base + tj * 4 + jj * 2;
                        float     v0, v1, g0 = 0.f, g1 = 0.f;
                        if (KT == 0)
                            v0 = tc::ex2_approx(fmaf(r2[i][jj].x, c1, c0)), v1 = tc::ex2_approx(fmaf(r2[i][jj].y, c1, c0));
                        else
                        {
                            const float s0 = sqrtf(5.f * r2[i][jj].x), s1 = sqrtf(5.f * r2[i][jj].y);
                            const float e0 = tc::ex2_approx(fmaf(s0, -1.44269504088896340736f, c0));
                            const float e1 = tc::ex2_approx(fmaf(s1, -1.44269504088896340736f, c0));
                            v0 = e0 * fmaf(s0, fmaf(s0, 0.33333333333333333f, 1.f), 1.f), v1 = e1 * fmaf(s1, fmaf(s1, 0.33333333333333333f, 1.f), 1.f);
                            g0 = -1.66666666666666667f * e0 * (1.f + s0), g1 = -1.66666666666666667f * e1 * (1.f + s1);
                        }
                        if (!interior)
                        {
                            if (!(m < Mc && j < N)) v0 = 0.f, g0 = 0.f;
                            if (!(m < Mc && j + 1 < N)) v1 = 0.f, g1 = 0.f;
                        }
                        h[jj]          = __floats2half2_rn(v0, v1);
                        const float2 b = __half22float2(h[jj]);
                        hl[jj]         = __floats2half2_rn(v0 - b.x, v1 - b.y);
                        if (KT != 0)
                        {
                            gh[jj]         = __floats2half2_rn(g0, g1);
                            const float2 c = __half22float2(gh[jj]);
                            gl[jj]         = __floats2half2_rn(g0 - c.x, g1 - c.y);
                        }
                    }
                    const size_t off = (size_t) m * ldt + j_base + tj * 4;
                    *reinterpret_cast<uint2*>(Ks + off) = *reinterpret_cast<const uint2*>(&h[0]);
                    if (Ks_lo) *reinterpret_cast<uint2*>(Ks_lo + off) = *reinterpret_cast<const uint2*>(&hl[0]);
                    if (KT != 0)
                    {
                        *reinterpret_cast<uint2*>(Gs + off) = *reinterpret_cast<const uint2*>(&gh[0]);
                        if (Gs_lo) *reinterpret_cast<uint2*>(Gs_lo + off) = *reinterpret_cast<const uint2*>(&gl[0]);
                    }
                }
                __syncthreads(); // buffer `buf` is refilled by the prefetch of the iteration after next
            }
        }
    }

    struct TcGemmParams
    {
        int              ldt;           // row length (elements) of Ks and Bmat; multiple of 256
        int              kb;            // pipeline steps along the contraction = round_up(N, 64) / 64
        int              ncb;           // regular column blocks = ldt / 256
        int              D;
        int              stages;        // shared-memory pipeline depth
        int              stage_bytes;   // bytes of one pipeline stage in ONE CTA (see tc_stage_bytes)
        int              n_cand_blocks; // ceil(Mc / 128), rounded up to the CTAs per cluster
        long long        Mc;
        int              passes;    // 1: k16 x A16 | 2: + k16 x A_lo | 3: + k_lo x A16 (split-fp16, fp32-class result)
        int              a_row0;    // first row of this shard buffer inside the Ks tensor map
        int              a_lo_row;  // row offset of the k residuals (relative to a_row0)
        int              b_lo_row;  // row offset of the Kinv residuals inside the Bmat tensor map
        const __half*    Ks_lo;     // k residuals (null when passes == 1)
        const __half*    Ks;
        int              matern;    // Matern 5/2: the gradient weight g is its own operand (SE: g = -c k)
        int              g_row;     // row offset (relative to a_row0) of the g operand inside the Ks tensor map
        const __half*    Gs;        // g values / residuals, same layout as Ks / Ks_lo (Matern only)
        const __half*    Gs_lo;
        const float*     Xt;
        const TcScales*  sc;
        double           se_factor; // c
        double4*         stats;     // per candidate (mu, q, ga, gb), as column_reduce_kernel writes them
        double*          P1;        // ldp x Mc
        double*          P2;
        int              ldp;
        int              split;     // S >= 1: S CTA groups share each candidate block and take 1/S of the column blocks each
        double2*         qx;        // S > 1: (q, gb) partials of the S - 1 groups that do not own the extras block, [S-1][part_stride]
        double*          P2x;       // S > 1: their P2 partials, [S-1][ldp x part_stride]
        long long        part_stride;
        int*             err;
    };

    // One pipeline stage = one 64-wide step of the contraction with EVERY operand tile the split-precision passes need,
    // so each tile is fetched from L2 once per step and reused by up to three UMMA groups:
    //   [k16 128x64] [k_lo 128x64 if passes == 3] [A_hi (256/NCTA)x64] [A_lo (256/NCTA)x64 if passes >= 2]
    // (L2 -> SM bytes per UMMA: 12 KB with one pass per stage, 8 KB here, 5.3 KB with the CTA pair.)
    __host__ __device__ inline int tc_stage_bytes(int passes, int ncta)
    {
        return TC_A_BYTES * (passes == 3 ? 2 : 1) + (TC_B_BYTES / ncta) * (passes >= 2 ? 2 : 1);
    }
    // columns of Xt staged in shared memory at a time (<= 20 KB whatever XP is)
    __host__ __device__ constexpr int tc_xs_cols(int XP) { return XP <= 20 ? 256 : (XP <= 36 ? 128 : 64); }

    // NCTA = 1: one CTA per SM, UMMA 128 x 256. NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) works on 256
    // candidates with UMMA 256 x 256; each CTA stages its own 128 k rows and HALF of every Kinv tile, the leader (cluster
    // rank 0) issues the MMAs for both SMs, and each CTA runs the epilogue of its own 128 TMEM lanes.
    template <int XP, int NCTA, bool MATERN>
    __global__ void __launch_bounds__(TC_THREADS, 1)
        tc_sweep_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const TcGemmParams p)
    {
        constexpr int EC       = (2 * XP + 15) / 16 * 16; // extras columns: (hi, lo) pairs
        constexpr int XS_COLS  = tc_xs_cols(XP);
        constexpr int BH_BYTES = TC_B_BYTES / NCTA;       // this CTA's share of a Kinv tile
        constexpr int BH_ROWS  = TC_BN / NCTA;
        extern __shared__ uint8_t smem_raw[];
        __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], tfull_bar[2], tempty_bar[2];
        __shared__ uint32_t tmem_base_smem;

        const uint32_t raw_addr = tc::smem_u32(smem_raw);
        const uint32_t pad      = ((raw_addr + 1023u) & ~1023u) - raw_addr;
        uint8_t*       smem     = smem_raw + pad; // 1024-byte aligned: required by the 128-byte swizzle
        const uint32_t smem_a0  = raw_addr + pad;
        float*         Xs       = reinterpret_cast<float*>(smem + (size_t) p.stages * p.stage_bytes); // [XS_COLS][XP]

        const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const uint32_t rank = NCTA == 2 ? tc::cluster_ctarank() : 0u;
        const int      cid = blockIdx.x / NCTA, ncl = gridDim.x / NCTA, n_groups = p.n_cand_blocks / NCTA;
        const int      P       = p.passes;
        const uint32_t off_a1  = TC_A_BYTES;                       // k_lo tile (passes == 3)
        const uint32_t off_b0  = TC_A_BYTES * (P == 3 ? 2 : 1);    // Kinv hi (or the extras)
        const uint32_t off_b1  = off_b0 + BH_BYTES;                // Kinv lo (passes >= 2)
        // Work item w = candidate group (NCTA x 128 candidates) x column-block range. With split == S, S consecutive work
        // items share the candidate rows and take 1/split of the Kinv column blocks each (the last one also the extras
        // block): the k* rows in flight at any time shrink by that factor, so they stay L2-resident across their re-reads.
        const int half_cb = (p.ncb + p.split - 1) / p.split;

        if (warp == 0 && lane == 0)
        {
            tc::tma_prefetch_desc(&tmA);
            tc::tma_prefetch_desc(&tmB);
        }
        if (warp == 1 && lane == 0)
        {
            for (int s = 0; s < p.stages; ++s)
            {
                tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
                tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
            }
            for (int s = 0; s < 2; ++s)
            {
                tc::mbar_init(tc::smem_u32(&tfull_bar[s]), 1);
                tc::mbar_init(tc::smem_u32(&tempty_bar[s]), 128 * NCTA);
            }
            tc::fence_mbar_init();
        }
        if (warp == 2)
        {
            if (NCTA == 2)
                tc::tmem_alloc_pair(tc::smem_u32(&tmem_base_smem), 512);
            else
                tc::tmem_alloc(tc::smem_u32(&tmem_base_smem), 512);
        }
        tc::fence_before_sync();
        __syncthreads();
        if (NCTA == 2) tc::cluster_sync_all(); // the peer's barriers are initialised before anything is posted on them
        tc::fence_after_sync();
        const uint32_t tmem_base = tmem_base_smem;

        if (warp == 0)
        {
            // ===== TMA producer: the whole warp walks the loop (warp-uniform control flow keeps the addresses in uniform
            // registers); one elected lane issues. Both CTAs of a pair post their bytes on the leader's barrier. =====
            const bool elected = tc::elect_one();
            uint32_t   it = 0;
            for (int w = cid; w < n_groups * p.split; w += ncl)
            {
                const int g = w / p.split, h = w - g * p.split;
                // the owner of the extras takes block ncb (extras x k) and, for Matern, block ncb + 1 (extras x g)
                const int cb_begin = h * half_cb, cb_end = min(p.ncb, cb_begin + half_cb), last = (h == p.split - 1) ? p.ncb + (MATERN ? 1 : 0) : cb_end - 1;
                const int cbk = g * NCTA + (int) rank, a_row_k = p.a_row0 + cbk * TC_BM;
                for (int cb = cb_begin; cb <= last; ++cb)
                {
                    const bool     extras = cb >= p.ncb;
                    const bool     two_b  = P >= 2 && !extras;
                    const int      a_row  = a_row_k + ((MATERN && cb == p.ncb + 1) ? p.g_row : 0);
                    const int      b_row  = extras ? p.ldt + (int) rank * (EC / NCTA) : cb * TC_BN + (int) rank * BH_ROWS;
                    const uint32_t tx     = (uint32_t) NCTA * (TC_A_BYTES * (P == 3 ? 2 : 1) + BH_BYTES * (two_b ? 2 : 1));
                    for (int k = 0; k < p.kb; ++k, ++it)
                    {
                        const uint32_t s = it % p.stages, n = it / p.stages;
                        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), (n & 1) ^ 1, p.err, 1);
                        const uint32_t sa = smem_a0 + s * p.stage_bytes;
                        if (elected)
                        {
                            if (NCTA == 1)
                            {
                                const uint32_t fb = tc::smem_u32(&full_bar[s]);
                                tc::mbar_arrive_expect_tx(fb, tx);
                                tc::tma_load_2d(sa, &tmA, fb, k * TC_BK, a_row);
                                if (P == 3) tc::tma_load_2d(sa + off_a1, &tmA, fb, k * TC_BK, a_row + p.a_lo_row);
                                tc::tma_load_2d(sa + off_b0, &tmB, fb, k * TC_BK, b_row);
                                if (two_b) tc::tma_load_2d(sa + off_b1, &tmB, fb, k * TC_BK, b_row + p.b_lo_row);
                            }
                            else
                            {
                                const uint32_t fb = tc::map_to_cta(tc::smem_u32(&full_bar[s]), 0);
                                if (rank == 0) tc::mbar_arrive_expect_tx(tc::smem_u32(&full_bar[s]), tx);
                                tc::tma_load_2d_pair(sa, &tmA, fb, k * TC_BK, a_row);
                                if (P == 3) tc::tma_load_2d_pair(sa + off_a1, &tmA, fb, k * TC_BK, a_row + p.a_lo_row);
                                tc::tma_load_2d_pair(sa + off_b0, &tmB, fb, k * TC_BK, b_row);
                                if (two_b) tc::tma_load_2d_pair(sa + off_b1, &tmB, fb, k * TC_BK, b_row + p.b_lo_row);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        else if (warp == 1)
        {
            // ===== MMA issuer (the leader CTA only when paired): warp-uniform loop, one elected lane issues =====
            if (rank == 0)
            {
                const bool     elected    = tc::elect_one();
                const uint32_t idesc_full = tc::instr_desc_f16(TC_BM * NCTA, TC_BN), idesc_extra = tc::instr_desc_f16(TC_BM * NCTA, EC);
                uint32_t       it = 0, t = 0;
                for (int w = cid; w < n_groups * p.split; w += ncl)
                {
                    const int h = w % p.split;
                    const int cb_begin = h * half_cb, cb_end = min(p.ncb, cb_begin + half_cb), last = (h == p.split - 1) ? p.ncb + (MATERN ? 1 : 0) : cb_end - 1;
                    for (int cb = cb_begin; cb <= last; ++cb, ++t)
                    {
                        const bool     extras = cb >= p.ncb;
                        const bool     two_b = P >= 2 && !extras, two_a = P == 3;
                        const uint32_t slot = t & 1, use = t >> 1;
                        tc::mbar_wait(tc::smem_u32(&tempty_bar[slot]), (use & 1) ^ 1, p.err, 2);
                        tc::fence_after_sync();
                        const uint32_t d_tmem = tmem_base + slot * TC_BN;
                        const uint32_t idesc  = extras ? idesc_extra : idesc_full;
                        for (int k = 0; k < p.kb; ++k, ++it)
                        {
                            const uint32_t s = it % p.stages, n = it / p.stages;
                            tc::mbar_wait(tc::smem_u32(&full_bar[s]), n & 1, p.err, 3);
                            tc::fence_after_sync();
                            if (elected)
                            {
                                const uint32_t sa  = smem_a0 + s * p.stage_bytes;
                                const uint64_t da0 = tc::smem_desc_k_sw128(sa), da1 = tc::smem_desc_k_sw128(sa + off_a1);
                                const uint64_t db0 = tc::smem_desc_k_sw128(sa + off_b0), db1 = tc::smem_desc_k_sw128(sa + off_b1);
                                const uint32_t acc0 = k != 0 ? 1u : 0u;
                                // 16 fp16 = 32 bytes = 2 descriptor units per UMMA_K; groups: k16 x A_hi, k16 x A_lo, k_lo x A_hi
#pragma unroll
                                for (int kk = 0; kk < TC_BK / 16; ++kk)
                                    tc::umma_f16_n<NCTA>(d_tmem, da0 + 2 * kk, db0 + 2 * kk, idesc, kk ? 1u : acc0);
                                if (two_b)
                                {
#pragma unroll
                                    for (int kk = 0; kk < TC_BK / 16; ++kk)
                                        tc::umma_f16_n<NCTA>(d_tmem, da0 + 2 * kk, db1 + 2 * kk, idesc, 1u);
                                }
                                if (two_a)
                                {
#pragma unroll
                                    for (int kk = 0; kk < TC_BK / 16; ++kk)
                                        tc::umma_f16_n<NCTA>(d_tmem, da1 + 2 * kk, db0 + 2 * kk, idesc, 1u);
                                }
                                // frees the stage (in both CTAs) once these MMAs have read it
                                tc::umma_commit_n<NCTA>(tc::smem_u32(&empty_bar[s]));
                                if (k == p.kb - 1) tc::umma_commit_n<NCTA>(tc::smem_u32(&tfull_bar[slot])); // accumulator complete
                            }
                            __syncwarp();
                        }
                    }
                }
            }
        }
        else if (warp >= 4)
        {
            // ===== epilogue: thread <-> candidate (TMEM lane) =====
            const int      quad = warp & 3, row = quad * 32 + lane, et = threadIdx.x - 128;
            const float    inv_u = p.sc->inv_u, inv_e = p.sc->inv_e;
            uint32_t       t = 0;
            // the accumulator-free barriers live in the leader CTA
            const uint32_t tempty0 = NCTA == 2 ? tc::map_to_cta(tc::smem_u32(&tempty_bar[0]), 0) : tc::smem_u32(&tempty_bar[0]);
            const uint32_t tempty1 = NCTA == 2 ? tc::map_to_cta(tc::smem_u32(&tempty_bar[1]), 0) : tc::smem_u32(&tempty_bar[1]);
            for (int w = cid; w < n_groups * p.split; w += ncl)
            {
                const int       g = w / p.split, h = w - g * p.split;
                const int       cb_begin = h * half_cb, cb_end = min(p.ncb, cb_begin + half_cb);
                const bool      owns_extras = h == p.split - 1;
                const int       cbk  = g * NCTA + (int) rank;
                const long long m    = (long long) cbk * TC_BM + row;
                const __half*   krow = p.Ks + (size_t) m * p.ldt;
                const __half*   lrow = p.Ks_lo ? p.Ks_lo + (size_t) m * p.ldt : nullptr;
                // passes == 2: u lacks A * dk, so q takes the first-order term 2 dk.u (q = k.A.k is symmetric in k)
                const float     qw = p.passes == 2 ? 2.f : 1.f;
                float2          acc[XP / 2]; // packed pairs: the reduction against Xt runs on FFMA2
#pragma unroll
                for (int c = 0; c < XP / 2; ++c) acc[c] = make_float2(0.f, 0.f);

                const uint4* kbase = reinterpret_cast<const uint4*>(krow);
                const uint4* lbase = reinterpret_cast<const uint4*>(lrow);
                uint4        kn[4], ln[4]; // software-pipelined: the k (and residual) values of the NEXT 32-column chunk
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    kn[v] = __ldg(kbase + cb_begin * (TC_BN / 8) + v), ln[v] = lrow ? __ldg(lbase + cb_begin * (TC_BN / 8) + v) : make_uint4(0, 0, 0, 0);
                // Matern: the gradient weight g_i is not a multiple of k_i and travels as its own operand
                const uint4* gbase  = MATERN ? reinterpret_cast<const uint4*>(p.Gs + (size_t) m * p.ldt) : nullptr;
                const uint4* glbase = (MATERN && p.Gs_lo) ? reinterpret_cast<const uint4*>(p.Gs_lo + (size_t) m * p.ldt) : nullptr;
                uint4        gn[MATERN ? 4 : 1], gln[MATERN ? 4 : 1];
                if constexpr (MATERN)
                {
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        gn[v] = __ldg(gbase + cb_begin * (TC_BN / 8) + v), gln[v] = glbase ? __ldg(glbase + cb_begin * (TC_BN / 8) + v) : make_uint4(0, 0, 0, 0);
                }

                for (int cb = cb_begin; cb < cb_end; ++cb, ++t)
                {
                    const uint32_t slot = t & 1, use = t >> 1;
                    const uint32_t taddr = tmem_base + ((uint32_t) (quad * 32) << 16) + slot * TC_BN;
#pragma unroll 1
                    for (int xs = 0; xs < TC_BN / XS_COLS; ++xs)
                    {
                        // stage the matching rows of Xt (XS_COLS x XP floats)
                        tc::named_bar_sync(1, 128);
                        {
                            const float4* src = reinterpret_cast<const float4*>(p.Xt + ((size_t) cb * TC_BN + (size_t) xs * XS_COLS) * XP);
                            float4*       dst = reinterpret_cast<float4*>(Xs);
                            for (int e = et; e < XS_COLS * XP / 4; e += 128) dst[e] = src[e];
                        }
                        tc::named_bar_sync(1, 128);
                        if (xs == 0)
                        {
                            tc::mbar_wait(tc::smem_u32(&tfull_bar[slot]), use & 1, p.err, 4);
                            tc::fence_after_sync();
                        }
#pragma unroll 1
                        for (int chl = 0; chl < XS_COLS / 32; ++chl)
                        {
                            const int ch = xs * (XS_COLS / 32) + chl;
                            uint32_t  r[32];
                            tc::tmem_ld_x32(taddr + ch * 32, r);
                            // this chunk's k values were fetched one chunk ago; fetch the next chunk's now
                            uint4 kv[4], lv[4], gv[MATERN ? 4 : 1], glv[MATERN ? 4 : 1];
#pragma unroll
                            for (int v = 0; v < 4; ++v) kv[v] = kn[v], lv[v] = ln[v];
                            if constexpr (MATERN)
                            {
#pragma unroll
                                for (int v = 0; v < 4; ++v) gv[v] = gn[v], glv[v] = gln[v];
                            }
                            {
                                const int nx = min(cb * (TC_BN / 32) + ch + 1, p.ncb * (TC_BN / 32) - 1);
#pragma unroll
                                for (int v = 0; v < 4; ++v) kn[v] = __ldg(kbase + nx * 4 + v);
#pragma unroll
                                for (int v = 0; v < 4; ++v) ln[v] = lrow ? __ldg(lbase + nx * 4 + v) : make_uint4(0, 0, 0, 0);
                                if constexpr (MATERN)
                                {
#pragma unroll
                                    for (int v = 0; v < 4; ++v) gn[v] = __ldg(gbase + nx * 4 + v), gln[v] = glbase ? __ldg(glbase + nx * 4 + v) : make_uint4(0, 0, 0, 0);
                                }
                            }
                            tc::tmem_ld_wait();
                            const __half2* kh  = reinterpret_cast<const __half2*>(kv);
                            const __half2* lh  = reinterpret_cast<const __half2*>(lv);
                            const __half2* gh  = reinterpret_cast<const __half2*>(gv);
                            const __half2* glh = reinterpret_cast<const __half2*>(glv);
#pragma unroll
                            for (int c2 = 0; c2 < 16; ++c2)
                            {
                                const float2 kf = __half22float2(kh[c2]), lf = __half22float2(lh[c2]);
                                float2       gf = make_float2(0.f, 0.f);
                                if constexpr (MATERN)
                                {
                                    const float2 g1 = __half22float2(gh[c2]), g2 = __half22float2(glh[c2]);
                                    gf              = make_float2(g1.x + g2.x, g1.y + g2.y);
                                }
#pragma unroll
                                for (int h = 0; h < 2; ++h)
                                {
                                    const int     c  = c2 * 2 + h;
                                    const float   u  = __uint_as_float(r[c]);
                                    const float   k1 = h ? kf.y : kf.x, dk = h ? lf.y : lf.x;
                                    const float   tv = MATERN ? (h ? gf.y : gf.x) * u : (k1 + dk) * u; // gradient sums
                                    const float   tq = fmaf(qw * dk, u, k1 * u); // quadratic form
                                    const float4* xr = reinterpret_cast<const float4*>(Xs + (chl * 32 + c) * XP);
                                    const float2  t2 = make_float2(tv, tv), tq2 = make_float2(tq, tv);
#pragma unroll
                                    for (int q4 = 0; q4 < XP / 4; ++q4)
                                    {
                                        const float4 xv = xr[q4];
                                        acc[q4 * 2 + 0] = tc::ffma2(make_float2(xv.x, xv.y), q4 == 0 ? tq2 : t2, acc[q4 * 2 + 0]);
                                        acc[q4 * 2 + 1] = tc::ffma2(make_float2(xv.z, xv.w), t2, acc[q4 * 2 + 1]);
                                    }
                                }
                            }
                        }
                    }
                    tc::fence_before_sync();
                    if (NCTA == 2)
                        tc::mbar_arrive_cluster(slot ? tempty1 : tempty0);
                    else
                        tc::mbar_arrive(slot ? tempty1 : tempty0);
                }

                // Xt columns: SE [1 | X_0 ..]: acc = (q, P2_0 ..) and gb = -c q, P2 = -c (.);  Matern [1 | 1 | X_0 ..]: acc = (q, gb, P2_0 ..)
                constexpr int XO = MATERN ? 2 : 1;
                const bool   live = m < p.Mc;
                const double c    = MATERN ? -1.0 : p.se_factor; // common factor -c of the SE sums; Matern sums carry g itself
                const double q    = (double) acc[0].x * (double) inv_u;
                const double gb   = MATERN ? (double) acc[0].y * (double) inv_u : -c * q;
                double*      P2o  = owns_extras ? p.P2 : p.P2x + (size_t) h * p.ldp * p.part_stride;
                if (live)
                {
#pragma unroll
                    for (int d = 0; d < XP - XO; ++d)
                        if (d < p.D) P2o[(size_t) d + (size_t) m * p.ldp] = -c * (double) (((XO + d) & 1) ? acc[(XO + d) >> 1].y : acc[(XO + d) >> 1].x) * (double) inv_u;
                    if (!owns_extras) p.qx[(size_t) h * p.part_stride + m] = make_double2(q, gb);
                }
                if (!owns_extras) continue;

                // extras blocks: column 2c = hi, 2c + 1 = lo of sum_j w_mj alpha_j (1, X_0j, ..)[c]; w = k (mu; for SE also ga and
                // P1 up to the factor -c), then for Matern a second block with w = g (ga, P1)
                double mu_k = 0.0;
#pragma unroll 1
                for (int eb = 0; eb < (MATERN ? 2 : 1); ++eb)
                {
                    const uint32_t slot = t & 1, use = t >> 1;
                    tc::mbar_wait(tc::smem_u32(&tfull_bar[slot]), use & 1, p.err, 5);
                    tc::fence_after_sync();
                    const uint32_t taddr = tmem_base + ((uint32_t) (quad * 32) << 16) + slot * TC_BN;
#pragma unroll
                    for (int ch = 0; ch < EC / 16; ++ch)
                    {
                        uint32_t r[16];
                        tc::tmem_ld_x16(taddr + ch * 16, r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                        {
                            const int    cc = ch * 8 + i;
                            const double S  = (double) (__uint_as_float(r[2 * i]) + __uint_as_float(r[2 * i + 1])) * (double) inv_e;
                            if (!MATERN)
                            {
                                if (live)
                                {
                                    if (cc == 0)
                                        p.stats[m] = make_double4(S, q, -c * S, gb);
                                    else if (cc <= p.D)
                                        p.P1[(size_t) (cc - 1) + (size_t) m * p.ldp] = -c * S;
                                }
                            }
                            else if (eb == 0)
                            {
                                if (cc == 0) mu_k = S;
                            }
                            else if (live)
                            {
                                if (cc == 0)
                                    p.stats[m] = make_double4(mu_k, q, S, gb);
                                else if (cc <= p.D)
                                    p.P1[(size_t) (cc - 1) + (size_t) m * p.ldp] = S;
                            }
                        }
                    }
                    tc::fence_before_sync();
                    if (NCTA == 2)
                        tc::mbar_arrive_cluster(slot ? tempty1 : tempty0);
                    else
                        tc::mbar_arrive(slot ? tempty1 : tempty0);
                    ++t;
                }
            }
        }

        tc::fence_before_sync();
        __syncthreads();
        if (NCTA == 2) tc::cluster_sync_all(); // the leader's MMAs read the peer's shared memory until the very end
        if (warp == 2)
        {
            tc::fence_after_sync();
            if (NCTA == 2)
                tc::tmem_dealloc_pair(tmem_base, 512);
            else
                tc::tmem_dealloc(tmem_base, 512);
        }
    }
} // namespace slsgp
