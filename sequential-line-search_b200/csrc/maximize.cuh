// Device-resident acquisition maximiser: the multi-start local search that follows the global stage of
// FindGlobalSolution (src/acquisition-function.cpp:112-167; the reference runs NLopt L-BFGS from DIRECT's point, or
// from `num_global_search_iters` random starts on threads). Here every start is advanced by ONE batched sweep per
// iteration and nothing but the winner ever leaves the device:
//   slice_argmax_kernel     best candidate of each of K equal slices of the swept range  -> K starting points
//   starts_from_slices      their coordinates (counter-based generator)
//   ascent_update_kernel    projected, normalised-gradient ascent with a per-start trust radius: accept / grow when the
//                           proposed point improved the acquisition value, else shrink and retry from the incumbent
#pragma once

#include "common.cuh"
#include "sweep.cuh"

namespace slsgp
{
    // One block per slice s = blockIdx.x + slice0: folds val[m] for the candidates of this shard that fall into slice s
    // into best[s]. A slice is `slice_len` consecutive candidate indices; (index0 + m) is the global candidate index.
    __global__ void __launch_bounds__(256)
        slice_argmax_kernel(const double* __restrict__ val, long long Mc, long long index0, long long first,
                            long long slice_len, long long slice0, ArgMax* __restrict__ best)
    {
        __shared__ ArgMax sm[256];
        const long long   s  = slice0 + blockIdx.x;
        const long long   lo = max(first + s * slice_len, index0), hi = min(first + (s + 1) * slice_len, index0 + Mc);
        ArgMax            b;
        b.v = 0.0, b.i = -1;
        for (long long g = lo + threadIdx.x; g < hi; g += 256)
        {
            const double v = val[g - index0];
            if (!isnan(v))
            {
                ArgMax c;
                c.v = v, c.i = g;
                b   = argmax_combine(b, c);
            }
        }
        sm[threadIdx.x] = b;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) sm[threadIdx.x] = argmax_combine(sm[threadIdx.x], sm[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) best[s] = argmax_combine(best[s], sm[0]);
    }

    struct AscentState // per start
    {
        double best_val; // incumbent value (-inf before the first evaluation)
        double radius;   // trust radius of the next proposal
    };

    // X (D x K): the points just evaluated; on exit the next proposals. Xbest / Gbest: incumbents and their gradients.
    __global__ void __launch_bounds__(128)
        starts_from_slices_kernel(const ArgMax* __restrict__ best, int K, int D, uint64_t seed, double radius0,
                                  double* __restrict__ X, double* __restrict__ Xbest, AscentState* __restrict__ st)
    {
        const int k = blockIdx.x * blockDim.x + threadIdx.x;
        if (k >= K) return;
        const long long i = best[k].i;
        for (int d = 0; d < D; ++d)
        {
            const double x = i >= 0 ? candidate_coord(seed, i, d) : 0.5; // an empty slice starts at the box centre
            X[(size_t) d + (size_t) k * D] = x, Xbest[(size_t) d + (size_t) k * D] = x;
        }
        st[k].best_val = -INFINITY, st[k].radius = radius0;
    }

    // One WARP per start, lanes over the dimensions: the D coordinates of a start are contiguous, so every access of the warp is a
    // coalesced segment (with one thread per start the three passes over D were strided by D doubles across the warp: 64 us per
    // launch for 1024 starts in 64 dimensions, half of an ascent iteration).
    __global__ void __launch_bounds__(128)
        ascent_update_kernel(int K, int D, const double* __restrict__ val, const double* __restrict__ grad,
                             double* __restrict__ X, double* __restrict__ Xbest, double* __restrict__ Gbest,
                             AscentState* __restrict__ st, double grow, double shrink, double radius_max)
    {
        const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (k >= K) return;
        double*       x  = X + (size_t) k * D;
        double*       xb = Xbest + (size_t) k * D;
        double*       gb = Gbest + (size_t) k * D;
        const double* g  = grad + (size_t) k * D;
        AscentState   s  = st[k];
        const double  v  = val[k];
        bool          bad = isnan(v);
        for (int d = lane; d < D; d += 32) bad = bad || isnan(g[d]) || isinf(g[d]);
        const bool finite = !__any_sync(0xffffffffu, bad);
        if (finite && v > s.best_val)
        {
            const bool first = isinf(s.best_val);
            s.best_val       = v;
            for (int d = lane; d < D; d += 32) xb[d] = x[d], gb[d] = g[d];
            if (!first) s.radius = fmin(s.radius * grow, radius_max);
        }
        else
            s.radius *= shrink;
        // projected gradient at the incumbent: components pushing out of [0, 1]^D are dropped
        // (each lane re-reads only the entries it wrote above, so no fence is needed)
        double n2 = 0.0;
        for (int d = lane; d < D; d += 32)
        {
            const double gd = ((xb[d] <= 0.0 && gb[d] < 0.0) || (xb[d] >= 1.0 && gb[d] > 0.0)) ? 0.0 : gb[d];
            n2 += gd * gd;
        }
        n2 = warp_sum(n2);
        n2 = __shfl_sync(0xffffffffu, n2, 0);
        const double scale = n2 > 0.0 ? s.radius / sqrt(n2) : 0.0;
        for (int d = lane; d < D; d += 32)
        {
            const double gd = ((xb[d] <= 0.0 && gb[d] < 0.0) || (xb[d] >= 1.0 && gb[d] > 0.0)) ? 0.0 : gb[d];
            x[d]            = fmin(fmax(xb[d] + scale * gd, 0.0), 1.0);
        }
        if (lane == 0) st[k] = s;
    }

    // arg-max over the incumbents; single block. out = (value, start index)
    __global__ void __launch_bounds__(256) ascent_winner_kernel(const AscentState* __restrict__ st, int K, ArgMax* __restrict__ out)
    {
        __shared__ ArgMax sm[256];
        ArgMax            b;
        b.v = 0.0, b.i = -1;
        for (int k = threadIdx.x; k < K; k += 256)
        {
            if (!isnan(st[k].best_val) && !isinf(st[k].best_val))
            {
                ArgMax c;
                c.v = st[k].best_val, c.i = k;
                b   = argmax_combine(b, c);
            }
        }
        sm[threadIdx.x] = b;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) sm[threadIdx.x] = argmax_combine(sm[threadIdx.x], sm[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) out[0] = sm[0];
    }
} // namespace slsgp
