// k* generator of the tensor-core sweep, squared distances on the tensor pipe (sm_100a).
//
// kstar16_strip_kernel forms r2 = sum_d (q_d - x_d)^2 with two packed FP32 instructions per (pair of observations, dimension):
// 38 issue slots per kernel value at D = 16, 19 % of a sweep step. Here
//     r2 = |q|^2 + |x|^2 - 2 q.x
// and the N x M x D contraction q.x runs as a split-fp16 UMMA (q = q_hi + q_lo, x = x_hi + x_lo, fp32 accumulation in TMEM):
//     Qh[m] = [q_hi | q_hi | q_lo | 0],  Xh[j] = [x_hi | x_lo | x_hi | 0]   (K = 3 D padded to 64, 128 or 192: D <= 64)
// so that one 128 x 256 x 64 UMMA group per tile yields q_hi.x_hi + q_hi.x_lo + q_lo.x_hi (the dropped q_lo.x_lo term is 2^-22
// relative): 3 D / N of the contraction work of the sweep itself (2 % at N = 2048). What is left per kernel value is the
// epilogue: one TMEM column, two FMAs, ex2, the fp16 hi / residual split and the stores - about 8 issue slots.
// q and x are the centred, length-scaled coordinates ((x - 1/2) / l), which keeps |q|^2 + |x|^2 small against the fp32
// cancellation in r2 (absolute error ~2^-24 (|q|^2 + |x|^2), i.e. a relative error of k of about 1e-6 at D = 16, l = 0.5: the
// same class as the direct-difference generator, whose exponent argument is rounded at |log2(a sK)| ~ 15). Candidates next to
// data points, where any error of k is amplified by a / sigma^2, are re-evaluated in IEEE double anyway (second tier).
// Same outputs as kstar16_strip_kernel: Ks / Ks_lo (and Gs / Gs_lo for Matern 5/2), zero for padding rows and columns.
//
// One persistent CTA per SM, 384 threads: warp 0 TMA producer, warp 1 UMMA issuer, warp 2 TMEM allocator, warps 4..19 epilogue
// (four warps per TMEM lane quarter, 64 columns each); the accumulator is double-buffered in TMEM (2 x 256 columns).
#pragma once

#include "common.cuh"
#include "tc.cuh"
#include "tc_sweep.cuh" // TcScales, TC_BK

namespace slsgp
{
    constexpr int KT_BM      = 128; // candidates per tile
    constexpr int KT_BN      = 256; // observations per tile
    constexpr int KT_EPI_WARPS = 16;                      // epilogue warps: four per TMEM lane quarter, 64 of the tile's columns each
    constexpr int KT_THREADS = (4 + KT_EPI_WARPS) * 32;
    constexpr int KT_EPI_COLS = KT_BN / (KT_EPI_WARPS / 4); // columns per epilogue warp
    constexpr int KT_A_BYTES = KT_BM * TC_BK * 2; // one 64-wide K slice of the candidate operand
    constexpr int KT_B_BYTES = KT_BN * TC_BK * 2;
    constexpr int KT_NX_SMEM = 4096; // observations whose squared norms are kept in shared memory (16 KB)
    constexpr float KT_FAR   = 1e30f; // |.|^2 of padding rows / columns: every kernel value they touch comes out as 0

    __host__ __device__ inline int kt_kp(int D) { return (3 * D + TC_BK - 1) / TC_BK * TC_BK; } // operand width (fp16 elements)

    // Xh[j] = [x_hi | x_lo | x_hi | 0 ..] and nx[j] = |x_j|^2 with x = (X_j - 1/2) / l; rows j >= N: zeros and KT_FAR.
    __global__ void __launch_bounds__(128)
        tc_pack_xh_kernel(const double* __restrict__ X, int N, int D, int ldt, int KP, const double* __restrict__ inv_l,
                          __half* __restrict__ Xh, float* __restrict__ nx)
    {
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= ldt) return;
        __half* row = Xh + (size_t) j * KP;
        double  n2  = 0.0;
        for (int d = 0; d < D; ++d)
        {
            const double x  = j < N ? (X[(size_t) d + (size_t) j * D] - 0.5) * inv_l[d] : 0.0;
            const __half hi = __double2half(x);
            const __half lo = __double2half(x - (double) __half2float(hi));
            row[d] = hi, row[D + d] = lo, row[2 * D + d] = hi;
            n2 += x * x;
        }
        for (int c = 3 * D; c < KP; ++c) row[c] = __float2half(0.f);
        nx[j] = j < N ? (float) n2 : KT_FAR;
    }

    // Qh[m] = [q_hi | q_hi | q_lo | 0 ..] and nq[m] = |q_m|^2 for the candidates of one shard; rows m >= Mc: zeros and KT_FAR.
    // One warp per candidate, lanes over the dimensions: the D coordinates of a candidate and the three D-wide pieces of its
    // operand row are contiguous (one thread per candidate walked rows 8 D bytes apart: 26 us per shard, a third of the generator).
    __global__ void __launch_bounds__(256)
        tc_pack_qh_kernel(const double* __restrict__ Xq, long long Mc, long long Mpad, int D, int KP, const double* __restrict__ inv_l,
                          __half* __restrict__ Qh, float* __restrict__ nq)
    {
        const long long m    = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int       lane = threadIdx.x & 31;
        if (m >= Mpad) return;
        __half* row = Qh + (size_t) m * KP;
        double  n2  = 0.0;
        for (int d = lane; d < D; d += 32)
        {
            const double q  = m < Mc ? (Xq[(size_t) d + (size_t) m * D] - 0.5) * inv_l[d] : 0.0;
            const __half hi = __double2half(q);
            const __half lo = __double2half(q - (double) __half2float(hi));
            row[d] = hi, row[D + d] = hi, row[2 * D + d] = lo;
            n2 += q * q;
        }
        for (int c = 3 * D + lane; c < KP; c += 32) row[c] = __float2half(0.f);
        n2 = warp_sum(n2);
        if (lane == 0) nq[m] = m < Mc ? (float) n2 : KT_FAR;
    }

    struct KstarTcParams
    {
        int             ldt;      // row length of Ks (multiple of 256)
        int             ncb;      // ldt / 256
        int             n_strips; // Mpad / 128
        int             q_row0;   // first row of this shard buffer inside the Qh tensor map
        int             stages;
        const float*    nq;       // [Mpad]
        const float*    nx;       // [ldt]
        const TcScales* sc;
        __half *        Ks, *Ks_lo, *Gs, *Gs_lo; // as kstar16_strip_kernel writes them (Ks_lo / Gs_lo may be null)
        int*            err;
    };

    template <int KT, int KS>
    __global__ void __launch_bounds__(KT_THREADS, 1)
        kstar_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX, const KstarTcParams p)
    {
        constexpr int STAGE_BYTES = KS * (KT_A_BYTES + KT_B_BYTES);
        constexpr int MAX_STAGES  = 4;
        extern __shared__ uint8_t smem_raw[];
        __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
        __shared__ uint32_t tmem_base_smem;

        const uint32_t raw_addr = tc::smem_u32(smem_raw);
        const uint32_t smem_a0  = (raw_addr + 1023u) & ~1023u; // the 128-byte swizzle wants 1024-byte aligned tiles
        // behind the pipeline stages: one 32 x 32 fp16 transposing buffer per epilogue warp (2 KB each, XOR-swizzled 16-byte pieces),
        // then the squared norms of all observations when they fit (ldt <= KT_NX_SMEM)
        uint4* const   stage_out = reinterpret_cast<uint4*>(smem_raw + (smem_a0 - raw_addr) + (size_t) p.stages * STAGE_BYTES);
        float* const   nx_s      = reinterpret_cast<float*>(stage_out + KT_EPI_WARPS * 32 * 4);
        const bool     nx_in_smem = p.ldt <= KT_NX_SMEM;
        if (nx_in_smem)
            for (int e = threadIdx.x; e < p.ldt; e += KT_THREADS) nx_s[e] = p.nx[e];
        const float* const nx_src = nx_in_smem ? nx_s : p.nx;
        const int      warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int      n_tiles = p.n_strips * p.ncb;

        if (warp == 0 && lane == 0)
        {
            tc::tma_prefetch_desc(&tmQ);
            tc::tma_prefetch_desc(&tmX);
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
                tc::mbar_init(tc::smem_u32(&tempty_bar[s]), KT_EPI_WARPS * 32);
            }
            tc::fence_mbar_init();
        }
        if (warp == 2) tc::tmem_alloc(tc::smem_u32(&tmem_base_smem), 512);
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        const uint32_t tmem_base = tmem_base_smem;

        if (warp == 0)
        {
            // ===== TMA producer: warp-uniform loop, one elected lane issues =====
            const bool elected = tc::elect_one();
            uint32_t   it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
            {
                const int      strip = t / p.ncb, cb = t - strip * p.ncb;
                const uint32_t s = it % p.stages, n = it / p.stages;
                tc::mbar_wait(tc::smem_u32(&empty_bar[s]), (n & 1) ^ 1, p.err, 11);
                if (elected)
                {
                    const uint32_t fb = tc::smem_u32(&full_bar[s]), sa = smem_a0 + s * STAGE_BYTES;
                    tc::mbar_arrive_expect_tx(fb, STAGE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                    {
                        tc::tma_load_2d(sa + ks * KT_A_BYTES, &tmQ, fb, ks * TC_BK, p.q_row0 + strip * KT_BM);
                        tc::tma_load_2d(sa + KS * KT_A_BYTES + ks * KT_B_BYTES, &tmX, fb, ks * TC_BK, cb * KT_BN);
                    }
                }
                __syncwarp();
            }
        }
        else if (warp == 1)
        {
            // ===== UMMA issuer =====
            const bool     elected = tc::elect_one();
            const uint32_t idesc   = tc::instr_desc_f16(KT_BM, KT_BN);
            uint32_t       it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
            {
                const uint32_t s = it % p.stages, n = it / p.stages, slot = it & 1, use = it >> 1;
                tc::mbar_wait(tc::smem_u32(&tempty_bar[slot]), (use & 1) ^ 1, p.err, 12);
                tc::mbar_wait(tc::smem_u32(&full_bar[s]), n & 1, p.err, 13);
                tc::fence_after_sync();
                if (elected)
                {
                    const uint32_t sa = smem_a0 + s * STAGE_BYTES, d_tmem = tmem_base + slot * KT_BN;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                    {
                        const uint64_t da = tc::smem_desc_k_sw128(sa + ks * KT_A_BYTES), db = tc::smem_desc_k_sw128(sa + KS * KT_A_BYTES + ks * KT_B_BYTES);
#pragma unroll
                        for (int kk = 0; kk < TC_BK / 16; ++kk) tc::umma_f16(d_tmem, da + 2 * kk, db + 2 * kk, idesc, (ks | kk) ? 1u : 0u);
                    }
                    tc::umma_commit(tc::smem_u32(&empty_bar[s]));     // the stage is free once the MMAs have read it
                    tc::umma_commit(tc::smem_u32(&tfull_bar[slot])); // and the accumulator complete
                }
                __syncwarp();
            }
        }
        else if (warp >= 4)
        {
            // ===== epilogue: thread <-> candidate row (TMEM lane), 128 of the tile's 256 columns per warp =====
            const int   quad = warp & 3, part = (warp - 4) >> 2;
            const float c1 = -0.72134752044448170368f; // -0.5 * log2(e)
            const float c0 = p.sc->c0;
            uint32_t    it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
            {
                const int       strip = t / p.ncb, cb = t - strip * p.ncb;
                const uint32_t  slot = it & 1, use = it >> 1;
                const long long m    = (long long) strip * KT_BM + quad * 32 + lane;
                const float     nqv  = p.nq[m];
                const float     hq   = fmaf(nqv, c1, c0);
                tc::mbar_wait(tc::smem_u32(&tfull_bar[slot]), use & 1, p.err, 14);
                tc::fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t) (quad * 32) << 16) + slot * KT_BN + part * KT_EPI_COLS;
                const int      j0    = cb * KT_BN + part * KT_EPI_COLS;
#pragma unroll 1
                for (int ch = 0; ch < KT_EPI_COLS / 32; ++ch)
                {
                    uint32_t r[32];
                    tc::tmem_ld_x32(taddr + ch * 32, r);
                    float nxv[32];
                    {
                        const float4* src = reinterpret_cast<const float4*>(nx_src + j0 + ch * 32); // warp-uniform address: a broadcast
#pragma unroll
                        for (int v = 0; v < 8; ++v)
                        {
                            const float4 f = src[v];
                            nxv[4 * v] = f.x, nxv[4 * v + 1] = f.y, nxv[4 * v + 2] = f.z, nxv[4 * v + 3] = f.w;
                        }
                    }
                    tc::tmem_ld_wait();
                    __half2 h[16], hl[16], gh[16], gl[16];
#pragma unroll
                    for (int c2 = 0; c2 < 16; ++c2)
                    {
                        float v[2], g[2] = {0.f, 0.f};
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                        {
                            const float dot = __uint_as_float(r[2 * c2 + e]), nxe = nxv[2 * c2 + e];
                            if (KT == 0)
                                v[e] = tc::ex2_approx(fmaf(dot, -2.f * c1, fmaf(nxe, c1, hq))); // a sK exp(-r2 / 2)
                            else
                            {
                                const float r2 = fmaxf(fmaf(dot, -2.f, nqv + nxe), 0.f);
                                const float s  = sqrtf(5.f * r2);
                                const float ex = tc::ex2_approx(fmaf(s, -1.44269504088896340736f, c0));
                                v[e]           = ex * fmaf(s, fmaf(s, 0.33333333333333333f, 1.f), 1.f);
                                g[e]           = -1.66666666666666667f * ex * (1.f + s);
                            }
                        }
                        h[c2]          = __floats2half2_rn(v[0], v[1]);
                        const float2 b = __half22float2(h[c2]);
                        hl[c2]         = __floats2half2_rn(v[0] - b.x, v[1] - b.y);
                        if (KT != 0)
                        {
                            gh[c2]          = __floats2half2_rn(g[0], g[1]);
                            const float2 cc = __half22float2(gh[c2]);
                            gl[c2]          = __floats2half2_rn(g[0] - cc.x, g[1] - cc.y);
                        }
                    }
                    // A thread holds 64 contiguous bytes of ITS row; stored directly, a warp instruction would touch 32 rows with 16
                    // bytes each (half a sector per row). Through the warp's transposing buffer four lanes write one row's 64 bytes
                    // and an instruction covers 8 rows x 2 whole sectors.
                    uint4* const    tb   = stage_out + (warp - 4) * (32 * 4);
                    const long long row0 = (long long) strip * KT_BM + quad * 32;
                    const size_t    col  = (size_t) j0 + ch * 32 + (lane & 3) * 8;
                    const auto      emit = [&](const __half2* src, __half* dst) {
                        __syncwarp();
#pragma unroll
                        for (int v = 0; v < 4; ++v) tb[lane * 4 + (v ^ ((lane >> 1) & 3))] = reinterpret_cast<const uint4*>(src)[v];
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const int rr = (lane >> 2) + 8 * k;
                            *reinterpret_cast<uint4*>(dst + (size_t) (row0 + rr) * p.ldt + col) = tb[rr * 4 + ((lane & 3) ^ ((rr >> 1) & 3))];
                        }
                    };
                    emit(h, p.Ks);
                    if (p.Ks_lo) emit(hl, p.Ks_lo);
                    if (KT != 0)
                    {
                        emit(gh, p.Gs);
                        if (p.Gs_lo) emit(gl, p.Gs_lo);
                    }
                }
                tc::fence_before_sync();
                tc::mbar_arrive(tc::smem_u32(&tempty_bar[slot]));
            }
        }

        tc::fence_before_sync();
        __syncthreads();
        if (warp == 2)
        {
            tc::fence_after_sync();
            tc::tmem_dealloc(tmem_base, 512);
        }
    }
} // namespace slsgp
