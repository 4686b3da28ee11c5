// Shared device/host helpers for libslsgp (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace slsgp
{
    constexpr int TILE = 64; // every N x N matrix is stored with leading dimension round_up(N, TILE)

    __host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
    __host__ __device__ inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

    // Kernel scalar functions. Both kernels are functions of the ARD-scaled squared distance
    //   r2 = sum_d ((xa_d - xb_d) / l_d)^2
    // (external/mathtoolbox/src/kernel-functions.cpp:16-17, 104-105). For each we need
    //   k(r2)  — the value (:7-20, :95-112), and
    //   g(r2)  — the scalar such that d k / d xa_d = g * (xa_d - xb_d) / l_d^2 (:81-93, :179-212).
    // SE:      k = a exp(-r2/2),                    g = -c k   with c = 2 in the reference (:92), 1 analytically
    // Matern:  k = a (1 + s + s^2/3) exp(-s), s = sqrt(5 r2);  g = -(5/3) a (1 + s) exp(-s), 0 when s < 1e-30 (:198)
    struct KernelVal
    {
        double k, g;
    };

    // exp(x) for x <= 0 (every kernel value is a * exp(-something non-negative)). libdevice's exp() costs ~45 issue slots per
    // call inside an unrolled tile (its 64-bit polynomial constants are re-materialised with move instructions at every call
    // site, plus the branches of its slow path); this one is 17 FP64 instructions with the coefficients read straight from the
    // constant bank: n = rint(x log2 e) by the magic-number add, r = x - n ln 2 in two FMAs (Cody-Waite), a degree-13 Taylor
    // polynomial on |r| <= ln 2 / 2 (truncation 4e-18), 2^n through the exponent field. Results below 2^-1021 flush to zero.
    // Accuracy ~1 ulp on [-707, 0]; the Gram-matrix parity tests hold K to 1e-13 of the CPU checker through it.
    __constant__ double c_exp_poly[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                          1.0 / 40320.0,      1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,
                                          1.0 / 6.0,          0.5,               1.0,              1.0};

    __device__ __forceinline__ double exp_nonpositive(double x)
    {
        const double shifter = 6755399441055744.0; // 2^52 + 2^51: adding it leaves rint(t) in the low mantissa bits
        double       t       = fma(x, 1.4426950408889634, shifter);
        const int    n       = __double2loint(t);
        t -= shifter;
        double r = fma(t, -6.93147180369123816490e-01, x);
        r        = fma(t, -1.90821492927058770002e-10, r);
        double p = c_exp_poly[0];
#pragma unroll
        for (int i = 1; i < 14; ++i) p = fma(p, r, c_exp_poly[i]);
        const double scaled = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
        return x < -707.0 ? 0.0 : scaled;
    }

    __device__ __forceinline__ double kernel_value(int kernel_type, double a, double r2)
    {
        if (kernel_type == 0) return a * exp(-0.5 * r2);
        const double s = sqrt(5.0 * r2);
        return a * (1.0 + s + (5.0 / 3.0) * r2) * exp(-s);
    }

    __device__ __forceinline__ KernelVal kernel_value_and_xgrad_weight(int kernel_type, double a, double r2,
                                                                       double se_xgrad_factor)
    {
        KernelVal v;
        if (kernel_type == 0)
        {
            v.k = a * exp(-0.5 * r2);
            v.g = -se_xgrad_factor * v.k;
        }
        else
        {
            const double s = sqrt(5.0 * r2);
            const double e = exp(-s);
            v.k            = a * (1.0 + s + (5.0 / 3.0) * r2) * e;
            v.g            = (s < 1e-30) ? 0.0 : -(5.0 / 3.0) * a * (1.0 + s) * e;
        }
        return v;
    }

    // d k / d theta weights: dk/da = ka(r2); dk/dl_t = kl(r2) * d_t^2 / l_t^3
    // SE (:22-50):      ka = exp(-r2/2),            kl = a exp(-r2/2)
    // Matern (:114-142): ka = (1+s+s^2/3) exp(-s),   kl = (5/3) a exp(-s) (1+s)
    __device__ __forceinline__ void kernel_theta_weights(int kernel_type, double a, double r2, double& ka, double& kl)
    {
        if (kernel_type == 0)
        {
            ka = exp(-0.5 * r2);
            kl = a * ka;
        }
        else
        {
            const double s = sqrt(5.0 * r2);
            const double e = exp(-s);
            ka             = (1.0 + s + (5.0 / 3.0) * r2) * e;
            kl             = (5.0 / 3.0) * a * e * (1.0 + s);
        }
    }

    __device__ __forceinline__ double warp_sum(double v)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }

    // Counter-based candidate generator: coordinate d of candidate i for a given seed, uniform in [0, 1).
    // splitmix64 finaliser over (seed, i, d); identical on host and device so sweeps are reproducible and
    // independent of how [first, first+count) is split across GPUs.
    __host__ __device__ inline double candidate_coord(uint64_t seed, int64_t i, int d)
    {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t) (i + 1) + 0xD1B54A32D192ED03ull * (uint64_t) (d + 1);
        z          = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z          = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z          = z ^ (z >> 31);
        return (double) (z >> 11) * (1.0 / 9007199254740992.0);
    }
} // namespace slsgp
