// Bordered update of a fitted model by one data point (the temporary regressor of FindNextPoints,
// src/acquisition-function.cpp:246-298, which the reference rebuilds and re-inverts for every pending option).
//
//   K' = [K k; k^T kappa],  L' = [L 0; l^T lambda],  l = L^-1 k,  lambda^2 = s = kappa - |l|^2,
//   W' = L'^-1 = [W 0; -u^T / lambda, 1 / lambda],   u = W^T l = K^-1 k,
//   K'^-1 = [K^-1 + u u^T / s, -u / s; -u^T / s, 1 / s],   logdet' = logdet + log s.
//
// O(N^2) instead of O(N^3); l and u come from the explicit triangular inverse W (two triangular GEMVs), never from K^-1, so
// the new pivot carries the conditioning of L rather than that of K.
#pragma once

#include "common.cuh"

namespace slsgp
{
    // vec[j] = k(X_j, x_new) with the rounding sequence of the Gram kernel (gram.cuh: scaled coordinates, one FMA per dim);
    // scalars[11] = kappa = k(x, x) + noise. x_new = column N of X.
    __global__ void __launch_bounds__(128)
        append_kvec_kernel(const double* __restrict__ X, int N, int D, int ld, const double* __restrict__ theta,
                           const double* __restrict__ inv_l, double noise, int kernel_type, double* __restrict__ vec,
                           double* __restrict__ scalars)
    {
        const int j = blockIdx.x * blockDim.x + threadIdx.x;
        if (j >= ld) return;
        double v = 0.0;
        if (j < N)
        {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d)
            {
                const double s  = inv_l[d];
                const double df = X[(size_t) d + (size_t) N * D] * s - X[(size_t) d + (size_t) j * D] * s;
                r2              = fma(df, df, r2);
            }
            v = kernel_value(kernel_type, theta[0], r2);
        }
        vec[j] = v;
        if (j == N) scalars[11] = kernel_value(kernel_type, theta[0], 0.0) + noise;
    }

    // One CTA: s = kappa - |l|^2 (scalars[10]); on success the new rows of L and W;
    // otherwise *info = N + 1 and nothing is written.
    __global__ void __launch_bounds__(1024)
        append_rows_kernel(int N, int ld, const double* __restrict__ l, const double* __restrict__ u, double* __restrict__ L,
                           double* __restrict__ W, double* __restrict__ scalars, int* __restrict__ info)
    {
        __shared__ double part[1024];
        const int tid = threadIdx.x;
        double    acc = 0.0;
        for (int j = tid; j < N; j += 1024) acc = fma(l[j], l[j], acc);
        part[tid] = acc;
        __syncthreads();
        for (int w = 512; w > 0; w >>= 1)
        {
            if (tid < w) part[tid] += part[tid + w];
            __syncthreads();
        }
        const double s = scalars[11] - part[0];
        if (!(s > 0.0) || !isfinite(s))
        {
            if (tid == 0) *info = N + 1, scalars[10] = s;
            return;
        }
        const double lambda = sqrt(s), inv_lambda = 1.0 / lambda;
        for (int j = tid; j < N; j += 1024)
        {
            L[(size_t) N + (size_t) j * ld] = l[j];
            W[(size_t) N + (size_t) j * ld] = -u[j] * inv_lambda;
        }
        if (tid == 0)
        {
            L[(size_t) N + (size_t) N * ld] = lambda;
            W[(size_t) N + (size_t) N * ld] = inv_lambda;
            scalars[10] = s;
        }
    }

    // New row / column of K, rank-one update and border of K^-1 (symmetric by construction: the product u_i u_j commutes).
    __global__ void __launch_bounds__(256)
        append_commit_kernel(int N, int ld, const double* __restrict__ k, const double* __restrict__ u,
                             const double* __restrict__ scalars, double* __restrict__ K, double* __restrict__ Kinv)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
        if (i > N) return;
        const double inv_s = 1.0 / scalars[10];
        const size_t at    = (size_t) i + (size_t) j * ld;
        if (i < N && j < N)
            Kinv[at] = fma(u[i] * u[j], inv_s, Kinv[at]);
        else if (i == N && j == N)
            Kinv[at] = inv_s, K[at] = scalars[11];
        else
        {
            const int m = i < j ? i : j;
            Kinv[at]    = -u[m] * inv_s;
            K[at]       = k[m];
        }
    }
} // namespace slsgp
