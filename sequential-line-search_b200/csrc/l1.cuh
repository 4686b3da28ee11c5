// The L1 free functions of the reference that the fused kernels never materialise, as stand-alone device kernels for callers
// that want the arrays themselves:
//   CalcSmallK                    (src/regressor.cpp:45-59)    k_i = k(x, X_i)
//   CalcSmallKSmallXDerivative    (src/regressor.cpp:91-108)   column i = d k(x, X_i) / d x
//   CalcLargeKYThetaDerivative    (src/regressor.cpp:110-134)  D + 1 matrices d K / d theta_t
// Same per-pair arithmetic as the Gram kernel (length-scaled coordinates, FMA-accumulated squared distance).
#pragma once

#include "common.cuh"

namespace slsgp
{
    // one thread per data point; x, theta (D + 1), inv_l (D) in device memory
    __global__ void small_k_kernel(const double* __restrict__ X, int N, int D, const double* __restrict__ x,
                                   const double* __restrict__ theta, const double* __restrict__ inv_l, int kernel_type,
                                   double se_xgrad_factor, double* __restrict__ k_out, double* __restrict__ dk_out)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= N) return;
        double r2 = 0.0;
        for (int d = 0; d < D; ++d)
        {
            const double s = inv_l[d], df = x[d] * s - X[(size_t) d + (size_t) i * D] * s;
            r2             = fma(df, df, r2);
        }
        const KernelVal v = kernel_value_and_xgrad_weight(kernel_type, theta[0], r2, se_xgrad_factor);
        if (k_out) k_out[i] = v.k;
        if (dk_out)
            for (int d = 0; d < D; ++d)
            {
                const double s = inv_l[d];
                dk_out[(size_t) d + (size_t) i * D] = v.g * ((x[d] - X[(size_t) d + (size_t) i * D]) * s * s);
            }
    }

    // out: (D + 1) matrices of N x N (column-major, leading dimension N), matrix 0 = d K / d a, matrix 1 + t = d K / d l_t.
    // grid (ceil(N / 16), ceil(N / 16)), block 16 x 16: one pair per thread, all D + 1 planes.
    __global__ void gram_theta_derivative_kernel(const double* __restrict__ X, int N, int D, const double* __restrict__ theta,
                                                 const double* __restrict__ inv_l, int kernel_type, double* __restrict__ out)
    {
        const int i = blockIdx.x * 16 + threadIdx.x, j = blockIdx.y * 16 + threadIdx.y;
        if (i >= N || j >= N) return;
        double r2 = 0.0;
        for (int d = 0; d < D; ++d)
        {
            const double s = inv_l[d], df = X[(size_t) d + (size_t) i * D] * s - X[(size_t) d + (size_t) j * D] * s;
            r2             = fma(df, df, r2);
        }
        double ka, kl;
        kernel_theta_weights(kernel_type, theta[0], r2, ka, kl);
        const size_t plane = (size_t) N * N, at = (size_t) i + (size_t) j * N;
        out[at]            = ka;
        for (int t = 0; t < D; ++t)
        {
            const double s = inv_l[t], df = X[(size_t) t + (size_t) i * D] - X[(size_t) t + (size_t) j * D];
            out[(size_t) (1 + t) * plane + at] = kl * (df * df) * (s * s * s);
        }
    }
} // namespace slsgp
