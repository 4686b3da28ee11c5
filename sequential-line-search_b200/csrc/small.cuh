// Small models (N <= 64, D <= 32) in ONE single-CTA launch: Gram matrix, Cholesky factor, its inverse, K^-1, alpha, f_best,
// the scalars of the GP term and the length-scale gradient. This is the regime the optimisers of the reference live in
// (tens of points), where the general path's ~20 launches and three host synchronisations per MAP objective evaluation are
// pure latency. Same arithmetic as the general kernels (gram.cuh rounding sequence for K, plain right-looking Cholesky,
// W = L^-1 by forward substitution, K^-1 = W^T W), same buffers, same padding conventions (identity on rows >= N).
#pragma once

#include "chol.cuh"
#include "common.cuh"
#include "sweep.cuh"

namespace slsgp
{
    constexpr int SMALL_N = 64, SMALL_LD = SMALL_N + 1, SMALL_DMAX = 32;
    constexpr int SMALL_REGS_FROM = 20; // from this N on: the register-resident 64-pivot factor + inverse of chol.cuh (fixed ~15 us)
    constexpr int SMALL_SMEM_BYTES = (3 * SMALL_N * SMALL_LD + SMALL_DMAX * SMALL_N + 5 * SMALL_N + 8 * SMALL_DMAX + 8) * (int) sizeof(double);

    struct SmallModelArgs
    {
        const double *X, *in;                            // in: [theta (D + 1) | 1 / l (D) | y (N)], one host-to-device copy
        int           N, D, kernel_type, want_hyper;
        double        noise;
        double *      theta, *inv_l, *y;                 // the context's own copies of the packed inputs (written here)
        double *      K, *L, *W, *Kinv, *alpha, *Kalpha; // ld = 64
        double *      out;                               // [0] y.alpha [1] alpha.alpha [2] tr K^-1 [3] logdet [4] bad pivot + 1 or 0
                                                         // [5 .. 5 + D) length-scale gradient: one device-to-host copy
        double *      fbest;
        int *         fbest_idx, *info;
        // Bradley-Terry-Luce part of the preference MAP objective, fused (btl_P < 0: not wanted). Tuples in CSR form and its
        // transpose as slsgp_set_preferences stores them; results: out[5 + D] = sum_t log BTL_t, out[6 + D + i] = d/dy_i of the
        // whole objective = gathered BTL terms - alpha_i (btl_grad != 0), so ONE device-to-host copy returns everything.
        int             btl_P, btl_grad;
        double          btl_scale;
        const uint32_t *pref_off, *pref_idx, *slot_off, *slot_list;
        double*         contrib;
    };

    __global__ void __launch_bounds__(256) small_model_kernel(const SmallModelArgs a)
    {
        extern __shared__ __align__(16) double ssm[];
        double* Kb   = ssm;                          // (i, j) at j * SMALL_LD + i
        double* Lb   = Kb + SMALL_N * SMALL_LD;      // factor, later K^-1
        double* Wb   = Lb + SMALL_N * SMALL_LD;      // L^-1
        double* Xs   = Wb + SMALL_N * SMALL_LD;      // Xs[d * 64 + i] = X[d, i] / l_d
        double* ys   = Xs + SMALL_DMAX * SMALL_N;
        double* al   = ys + SMALL_N;
        double* ka   = al + SMALL_N;
        double* dg   = ka + SMALL_N;                 // 1 / L_jj
        double* dq   = dg + SMALL_N;                 // L_jj
        double* part = dq + SMALL_N;                 // [8][SMALL_DMAX]
        int*    bad  = reinterpret_cast<int*>(part + 8 * SMALL_DMAX);

        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        const int N = a.N, D = a.D;
        const double* __restrict__ theta_in = a.in;
        const double* __restrict__ inv_l_in = a.in + D + 1;
        const double* __restrict__ y_in     = a.in + 2 * D + 1;
        const double sig = theta_in[0];

        for (int e = tid; e < D * SMALL_N; e += 256)
        {
            const int i = e & 63, d = e >> 6;
            Xs[e]       = i < N ? a.X[(size_t) d + (size_t) i * D] * inv_l_in[d] : 0.0;
        }
        if (tid < SMALL_N)
        {
            ys[tid]  = tid < N ? y_in[tid] : 0.0;
            a.y[tid] = ys[tid];
        }
        if (tid <= D) a.theta[tid] = theta_in[tid];
        if (tid < D) a.inv_l[tid] = inv_l_in[tid];
        if (tid == 0) *bad = 0;
        __syncthreads();

        // ---- K_y (padding: identity) ---------------------------------------------------------------------------------
        for (int e = tid; e < SMALL_N * SMALL_N; e += 256)
        {
            const int i = e & 63, j = e >> 6;
            double    v;
            if (i < N && j < N)
            {
                double r2 = 0.0;
                for (int d = 0; d < D; ++d)
                {
                    const double df = Xs[d * SMALL_N + i] - Xs[d * SMALL_N + j];
                    r2              = fma(df, df, r2);
                }
                v = kernel_value(a.kernel_type, sig, r2) + (i == j ? a.noise : 0.0);
            }
            else
                v = i == j ? 1.0 : 0.0;
            Kb[j * SMALL_LD + i] = v;
            Lb[j * SMALL_LD + i] = i >= j ? v : 0.0;
            a.K[e]               = v;
        }

        static_assert(TILE == SMALL_N, "potf2_inverse_regs works on 64 x 64 tiles");
        if (N >= SMALL_REGS_FROM)
        {
            // ---- factor and triangular inverse in registers (potf2_inverse_regs_pair, chol.cuh: one barrier per two pivots) -----------
            const int tx = tid & 15, ty = tid >> 4;
            double    c[4][4];
            __syncthreads(); // Kb complete
#pragma unroll
            for (int jq = 0; jq < 4; ++jq)
#pragma unroll
                for (int iq = 0; iq < 4; ++iq)
                {
                    const int p = tx + 16 * iq, q = ty + 16 * jq;
                    c[iq][jq]   = q <= p ? Kb[q * SMALL_LD + p] : 0.0;
                }
            // column buffers + pivots live in Wb, which is only written after the factorisation
            potf2_inverse_regs_pair(c, Wb, Wb + 4 * (TILE + 2), dg, bad, tx, ty, 0, a.info); // dg = d^-1/2
            if (*bad != 0x7fffffff)
            {
                if (tid == 0) a.out[4] = (double) (*bad + 1);
                return;
            }
#pragma unroll
            for (int jq = 0; jq < 4; ++jq)
#pragma unroll
                for (int iq = 0; iq < 4; ++iq)
                {
                    const int    p = tx + 16 * iq, q = ty + 16 * jq;
                    const double v = c[iq][jq] * dg[q];
                    Lb[q * SMALL_LD + p] = q <= p ? v : 0.0;                        // L(p, q)
                    Wb[p * SMALL_LD + q] = q > p ? v : (q == p ? dg[q] : 0.0);      // W(q, p): rows q >= p of column p
                    if (p == q) dq[p] = v;
                }
            __syncthreads();
            for (int e = tid; e < SMALL_N * SMALL_N; e += 256) a.L[e] = Lb[(e >> 6) * SMALL_LD + (e & 63)];
        }
        else
        {
            // ---- right-looking Cholesky of Lb. The diagonal keeps the running pivot until the end (every thread reads it at the
            // top of its step; the square roots are parked in dq), so a step needs two barriers only. ---------------------------
            if (tid < SMALL_N) dq[tid] = 1.0, dg[tid] = 1.0; // rows >= N: the identity
            for (int j = 0; j < N; ++j)
            {
                __syncthreads();
                const double piv = Lb[j * SMALL_LD + j];
                if (!(piv > 0.0) || !isfinite(piv))
                {
                    if (tid == 0) *bad = j + 1;
                    break; // uniform: every thread read the same pivot
                }
                const double d = sqrt(piv), rd = 1.0 / d;
                if (tid < SMALL_N)
                {
                    if (tid > j && tid < N) Lb[j * SMALL_LD + tid] *= rd;
                    if (tid == j) dq[j] = d, dg[j] = rd;
                }
                __syncthreads();
    #pragma unroll 4
                for (int e = tid; e < (N - j - 1) * SMALL_N; e += 256) // columns j + 1 .. N - 1 only
                {
                    const int i = e & 63, c = j + 1 + (e >> 6);
                    if (i >= c && i < N) Lb[c * SMALL_LD + i] = fma(-Lb[j * SMALL_LD + i], Lb[j * SMALL_LD + c], Lb[c * SMALL_LD + i]);
                }
            }
            __syncthreads();
            if (*bad != 0)
            {
                if (tid == 0) a.out[4] = (double) *bad;
                return;
            }
            if (tid < SMALL_N) Lb[tid * SMALL_LD + tid] = dq[tid];
            __syncthreads();
            for (int e = tid; e < SMALL_N * SMALL_N; e += 256) a.L[e] = Lb[(e >> 6) * SMALL_LD + (e & 63)];

            // ---- W = L^-1: column c by forward substitution, four lanes sharing each dot product ----------------------------
            {
                const int c = tid >> 2, q = tid & 3;
                if (q == 0)
                {
                    for (int i = 0; i < SMALL_N; ++i) Wb[c * SMALL_LD + i] = 0.0;
                    Wb[c * SMALL_LD + c] = dg[c];
                }
                __syncwarp();
                for (int i = 1; i < N; ++i) // same trip count for the eight columns of a warp: the shuffles are warp-wide
                {
                    double s = 0.0, s2 = 0.0;
                    if (i > c)
                    {
                        int k = c + q;
                        for (; k + 4 < i; k += 8)
                        {
                            s  = fma(Lb[k * SMALL_LD + i], Wb[c * SMALL_LD + k], s);
                            s2 = fma(Lb[(k + 4) * SMALL_LD + i], Wb[c * SMALL_LD + k + 4], s2);
                        }
                        if (k < i) s = fma(Lb[k * SMALL_LD + i], Wb[c * SMALL_LD + k], s);
                        s += s2;
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (i > c && q == 0) Wb[c * SMALL_LD + i] = -s * dg[i];
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        for (int e = tid; e < SMALL_N * SMALL_N; e += 256) a.W[e] = Wb[(e >> 6) * SMALL_LD + (e & 63)];

        // ---- K^-1 = W^T W (into Lb), alpha = K^-1 y, K alpha -----------------------------------------------------------
        __syncthreads();
        for (int e = tid; e < SMALL_N * SMALL_N; e += 256)
        {
            const int i = e & 63, j = e >> 6;
            double    s = 0.0;
            if (i < N && j < N)
            {
                double s2 = 0.0;
                int    k  = i > j ? i : j;
                for (; k + 1 < N; k += 2)
                {
                    s  = fma(Wb[i * SMALL_LD + k], Wb[j * SMALL_LD + k], s);
                    s2 = fma(Wb[i * SMALL_LD + k + 1], Wb[j * SMALL_LD + k + 1], s2);
                }
                if (k < N) s = fma(Wb[i * SMALL_LD + k], Wb[j * SMALL_LD + k], s);
                s += s2;
            }
            else
                s = i == j ? 1.0 : 0.0;
            a.Kinv[e] = s;
            Lb[j * SMALL_LD + i] = s; // no thread reads L any more
        }
        __syncthreads();
        if (tid < SMALL_N)
        {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s = fma(Lb[j * SMALL_LD + tid], ys[j], s);
            al[tid]      = s;
            a.alpha[tid] = s;
        }
        __syncthreads();
        if (tid < SMALL_N)
        {
            double s = 0.0;
            for (int j = 0; j < N; ++j) s = fma(Kb[j * SMALL_LD + tid], al[j], s);
            ka[tid]       = s;
            a.Kalpha[tid] = s;
        }
        __syncthreads();
        if (tid < SMALL_N) // two warps: per-point terms, shuffle reductions, combined by thread 0 through `part`
        {
            const bool in = tid < N;
            double     ya = in ? ys[tid] * al[tid] : 0.0, aa = in ? al[tid] * al[tid] : 0.0, tr = in ? Lb[tid * SMALL_LD + tid] : 0.0;
            double     ld = in ? log(dq[tid]) : 0.0;
            ArgMax     best;
            best.v = in ? ka[tid] - a.noise * al[tid] : 0.0, best.i = in ? tid : -1;
            ya = warp_sum(ya), aa = warp_sum(aa), tr = warp_sum(tr), ld = warp_sum(ld);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                ArgMax other;
                other.v = __shfl_xor_sync(0xffffffffu, best.v, o), other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
                best    = argmax_combine(best, other);
            }
            if (lane == 0)
            {
                double* p = part + warp * 8;
                p[0] = ya, p[1] = aa, p[2] = tr, p[3] = ld, p[4] = best.v, p[5] = (double) best.i;
            }
        }
        __syncthreads();
        if (tid == 0)
        {
            ArgMax b0, b1;
            b0.v = part[4], b0.i = (long long) part[5], b1.v = part[12], b1.i = (long long) part[13];
            const ArgMax best = argmax_combine(b0, b1);
            a.out[0] = part[0] + part[8], a.out[1] = part[1] + part[9], a.out[2] = part[2] + part[10];
            a.out[3] = 2.0 * (part[3] + part[11]), a.out[4] = 0.0;
            a.fbest[0] = best.v, a.fbest_idx[0] = (int) best.i;
        }
        __syncthreads(); // `part` is reused by the reductions below
        if (a.btl_P >= 0)
        {
            // same arithmetic as btl_tuple_kernel / btl_gather_kernel (map.cuh), on the y and alpha already in shared memory
            double ll = 0.0;
            for (int t = tid; t < a.btl_P; t += 256)
            {
                const uint32_t b = a.pref_off[t], e = a.pref_off[t + 1];
                const double   f0 = ys[a.pref_idx[b]];
                double         sum = 0.0;
                for (uint32_t i = b; i < e; ++i) sum += exp((1.0 / a.btl_scale) * ys[a.pref_idx[i]]);
                const double btl = exp((1.0 / a.btl_scale) * f0) / sum;
                ll += log(btl);
                if (a.btl_grad)
                {
                    const double tmp = -btl * btl / a.btl_scale;
                    double       s2  = 0.0;
                    for (uint32_t i = b + 1; i < e; ++i) s2 += exp((ys[a.pref_idx[i]] - f0) / a.btl_scale);
                    a.contrib[b] = (tmp * (-s2)) / btl;
                    for (uint32_t i = b + 1; i < e; ++i) a.contrib[i] = (tmp * exp((ys[a.pref_idx[i]] - f0) / a.btl_scale)) / btl;
                }
            }
            ll = warp_sum(ll);
            if (lane == 0) part[warp] = ll;
            __syncthreads(); // also orders the contrib[] writes (global) before the gather below
            if (tid == 0)
            {
                double sum = 0.0;
                for (int w = 0; w < 8; ++w) sum += part[w];
                a.out[5 + D] = sum;
            }
            if (a.btl_grad && tid < N)
            {
                double g = 0.0;
                if (a.btl_P > 0)
                    for (uint32_t q = a.slot_off[tid]; q < a.slot_off[tid + 1]; ++q) g += a.contrib[a.slot_list[q]];
                a.out[6 + D + tid] = g + -al[tid];
            }
            __syncthreads();
        }
        if (!a.want_hyper) return;

        // ---- length-scale gradient: G_t = 1 / (2 l_t) sum_ij (alpha_i alpha_j - Kinv_ij) kl(r2_ij) ((x_it - x_jt) / l_t)^2 ----
        double acc[SMALL_DMAX];
#pragma unroll
        for (int t = 0; t < SMALL_DMAX; ++t) acc[t] = 0.0;
        for (int e = tid; e < SMALL_N * SMALL_N; e += 256)
        {
            const int i = e & 63, j = e >> 6;
            if (i >= N || j >= N || i == j) continue;
            double r2 = 0.0;
            for (int d = 0; d < D; ++d)
            {
                const double df = Xs[d * SMALL_N + i] - Xs[d * SMALL_N + j];
                r2              = fma(df, df, r2);
            }
            double kav, kl;
            kernel_theta_weights(a.kernel_type, sig, r2, kav, kl);
            const double w = (al[i] * al[j] - Lb[j * SMALL_LD + i]) * kl;
#pragma unroll
            for (int t = 0; t < SMALL_DMAX; ++t)
                if (t < D)
                {
                    const double df = Xs[t * SMALL_N + i] - Xs[t * SMALL_N + j];
                    acc[t]          = fma(w, df * df, acc[t]);
                }
        }
#pragma unroll
        for (int t = 0; t < SMALL_DMAX; ++t)
        {
            const double v = warp_sum(acc[t]);
            if (lane == 0) part[warp * SMALL_DMAX + t] = v;
        }
        __syncthreads();
        if (tid < D)
        {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += part[w * SMALL_DMAX + tid];
            a.out[5 + tid] = 0.5 * s / theta_in[1 + tid];
        }
    }
} // namespace slsgp
