// K5 / K6: pieces of the MAP objectives that are not plain dense algebra.
#pragma once

#include "common.cuh"

namespace slsgp
{
    // K6. One thread per preference tuple t = idx[off[t] .. off[t+1]) (first member preferred):
    //   loglik[t]        = log BTL(y_tuple)                            calc_log_likelihood, preference-regressor.cpp:118-126
    //   contrib[off+i]   = (d BTL / d f_i) / BTL                       :198-216 with utils.hpp:25-52
    // BTL(f) = exp(f_0/s) / sum_i exp(f_i/s), un-stabilised exactly as utils.hpp:25-29 (overflows for f/s > 709).
    __global__ void btl_tuple_kernel(const double* __restrict__ y, const uint32_t* __restrict__ off,
                                     const uint32_t* __restrict__ idx, int P, double scale,
                                     double* __restrict__ loglik, double* __restrict__ contrib, int want_grad)
    {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= P) return;
        const uint32_t b = off[t], e = off[t + 1];
        const double   f0 = y[idx[b]];
        double         sum = 0.0;
        for (uint32_t i = b; i < e; ++i) sum += exp((1.0 / scale) * y[idx[i]]);
        const double btl = exp((1.0 / scale) * f0) / sum;
        loglik[t]        = log(btl);
        if (!want_grad) return;
        const double tmp = -btl * btl / scale; // CalcBtlDerivative, utils.hpp:31-52
        double       s2  = 0.0;
        for (uint32_t i = b + 1; i < e; ++i) s2 += exp((y[idx[i]] - f0) / scale);
        contrib[b] = (tmp * (-s2)) / btl;
        for (uint32_t i = b + 1; i < e; ++i) contrib[i] = (tmp * exp((y[idx[i]] - f0) / scale)) / btl;
    }

    // grad_y[i] = sum over the tuple slots that reference point i (fixed order: ascending slot) - alpha_i
    // (:198-219). slot_off / slot_list: CSR transpose of idx built on the host by slsgp_set_preferences.
    __global__ void btl_gather_kernel(const double* __restrict__ contrib, const uint32_t* __restrict__ slot_off,
                                      const uint32_t* __restrict__ slot_list, const double* __restrict__ alpha, int N,
                                      double* __restrict__ grad_y)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= N) return;
        double s = 0.0;
        for (uint32_t p = slot_off[i]; p < slot_off[i + 1]; ++p) s += contrib[slot_list[p]];
        grad_y[i] = alpha ? s + -alpha[i] : s; // alpha == null: the likelihood part alone (whitened objective)
    }

    // The whitened MAP objective (slsgp_map_objective_pref_whitened) for N <= 1024 in ONE single-block launch: at these sizes
    // the five kernels of the general path are pure launch latency, and the quasi-Newton driver calls this hundreds of times
    // per fit.  y = L z  ->  BTL terms of every tuple  ->  gather  ->  gz = L^T g - z.   out = [loglik | gz (N) | y (N)].
    // One thread per row for the triangular products' row form, one warp per column for the transposed one.
    __global__ void __launch_bounds__(1024)
        map_whitened_fused_kernel(const double* __restrict__ L, int N, int ld, const double* __restrict__ z_in,
                                  const uint32_t* __restrict__ off, const uint32_t* __restrict__ idx, int P, double scale,
                                  const uint32_t* __restrict__ slot_off, const uint32_t* __restrict__ slot_list,
                                  double* __restrict__ contrib, double* __restrict__ y_dev, int want_grad, double* __restrict__ out)
    {
        extern __shared__ double fsm[];
        double*   z = fsm;          // [N]
        double*   y = fsm + N;      // [N]
        double*   g = fsm + 2 * N;  // [N]
        double*   red = fsm + 3 * N; // [32]
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        for (int i = tid; i < N; i += 1024) z[i] = z_in[i];
        __syncthreads();
        for (int i = tid; i < N; i += 1024) // y_i = sum_{j <= i} L_ij z_j: consecutive threads read consecutive rows of column j
        {
            double s = 0.0;
#pragma unroll 4
            for (int j = 0; j <= i; ++j) s = fma(L[(size_t) i + (size_t) j * ld], z[j], s);
            y[i] = s, y_dev[i] = s, out[1 + N + i] = s;
        }
        __syncthreads();
        double ll = 0.0;
        for (int t = tid; t < P; t += 1024) // btl_tuple_kernel, same arithmetic
        {
            const uint32_t b = off[t], e = off[t + 1];
            const double   f0 = y[idx[b]];
            double         sum = 0.0;
            for (uint32_t i = b; i < e; ++i) sum += exp((1.0 / scale) * y[idx[i]]);
            const double btl = exp((1.0 / scale) * f0) / sum;
            ll += log(btl);
            if (want_grad)
            {
                const double tmp = -btl * btl / scale;
                double       s2  = 0.0;
                for (uint32_t i = b + 1; i < e; ++i) s2 += exp((y[idx[i]] - f0) / scale);
                contrib[b] = (tmp * (-s2)) / btl;
                for (uint32_t i = b + 1; i < e; ++i) contrib[i] = (tmp * exp((y[idx[i]] - f0) / scale)) / btl;
            }
        }
        ll = warp_sum(ll);
        if (lane == 0) red[warp] = ll;
        __syncthreads(); // also publishes contrib[] (global) to the block
        if (tid == 0)
        {
            double s = 0.0;
            for (int w = 0; w < 32; ++w) s += red[w];
            out[0] = s;
        }
        if (!want_grad) return;
        for (int i = tid; i < N; i += 1024)
        {
            double s = 0.0;
            for (uint32_t p = slot_off[i]; p < slot_off[i + 1]; ++p) s += contrib[slot_list[p]];
            g[i] = s;
        }
        __syncthreads();
        for (int j = warp; j < N; j += 32) // gz_j = sum_{i >= j} L_ij g_i - z_j
        {
            double s = 0.0;
            for (int i = j + lane; i < N; i += 32) s = fma(L[(size_t) i + (size_t) j * ld], g[i], s);
            s = warp_sum(s);
            if (lane == 0) out[1 + j] = s - z[j];
        }
    }

    // Deterministic sum of n doubles into out[0] (single block).
    __global__ void __launch_bounds__(256) sum_kernel(const double* __restrict__ v, int n, double* __restrict__ out)
    {
        __shared__ double part[256];
        double            s = 0.0;
        for (int i = threadIdx.x; i < n; i += 256) s += v[i];
        part[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[0] = part[0];
    }

    // Scalars of the GP term, single block:
    //   out[0] = y . alpha     out[1] = alpha . alpha     out[2] = tr(Kinv)
    __global__ void __launch_bounds__(256)
        gp_scalars_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                          const double* __restrict__ Kinv, int N, int ld, double* __restrict__ out)
    {
        __shared__ double p0[256], p1[256], p2[256];
        double            a = 0.0, b = 0.0, c = 0.0;
        for (int i = threadIdx.x; i < N; i += 256)
        {
            a = fma(y[i], alpha[i], a);
            b = fma(alpha[i], alpha[i], b);
            if (Kinv) c += Kinv[(size_t) i + (size_t) i * ld];
        }
        p0[threadIdx.x] = a, p1[threadIdx.x] = b, p2[threadIdx.x] = c;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o)
                p0[threadIdx.x] += p0[threadIdx.x + o], p1[threadIdx.x] += p1[threadIdx.x + o],
                    p2[threadIdx.x] += p2[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[0] = p0[0], out[1] = p1[0], out[2] = p2[0];
    }

    // Length-scale part of the hyper-gradient. With Wm = (alpha alpha^T - Kinv) o kl and Y = Wm * XT1, where
    // XT1 (ld x ldx) holds X^T in columns 0..D-1 and ones in column D:
    //   sum_ij Wm_ij (x_it - x_jt)^2 = 2 [ sum_i x_it^2 s_i - sum_i x_it Y_it ],   s = Y[:, D] = Wm 1
    //   G_t = 1/2 * (1 / l_t^3) * that sum                (term_1 + term_2 of preference-regressor.cpp:93-101)
    // One block per t; deterministic.
    __global__ void __launch_bounds__(256)
        lengthscale_grad_kernel(const double* __restrict__ XT1, const double* __restrict__ Y, int N, int ld, int D,
                                const double* __restrict__ theta, double* __restrict__ g_l, int splits, size_t split_stride)
    {
        // Y arrives as `splits` partial products (slices of the contraction), split_stride elements apart
        __shared__ double part[256];
        const int         t = blockIdx.x;
        double            s = 0.0;
        for (int i = threadIdx.x; i < N; i += 256)
        {
            const double x = XT1[(size_t) i + (size_t) t * ld];
            double       yd = 0.0, yt = 0.0;
            for (int z = 0; z < splits; ++z)
                yd += Y[z * split_stride + (size_t) i + (size_t) D * ld], yt += Y[z * split_stride + (size_t) i + (size_t) t * ld];
            s += x * x * yd - x * yt;
        }
        part[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0)
        {
            const double l = theta[1 + t];
            g_l[t]         = part[0] / (l * l * l);
        }
    }
} // namespace slsgp
