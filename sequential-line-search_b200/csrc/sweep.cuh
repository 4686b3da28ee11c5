// K4 (FP64 path): batched posterior mean / variance / gradients and acquisition values for a block of candidates.
//
// For one candidate x the reference evaluates (src/preference-regressor.cpp:293-330, src/regressor.cpp:45-59,91-108)
//   k_i = k(x, X_i),  J_di = d k_i / d x_d = g_i (x_d - X_di) / l_d^2
//   mu = k . alpha                    sigma^2 = a - k . beta,  beta = K^-1 k
//   grad mu = J alpha                 grad sigma = -(1/sigma) J beta
// so with  ga = sum_i g_i alpha_i, gb = sum_i g_i beta_i,  P1_d = sum_i X_di g_i alpha_i,  P2_d = sum_i X_di g_i beta_i:
//   grad mu_d = (x_d ga - P1_d) / l_d^2,        grad sigma_d = -(1/sigma) (x_d gb - P2_d) / l_d^2.
// For a block of Mc candidates this is: k*/g* tiles -> one N x N x Mc GEMM (beta) -> column reductions ->
// two D x N x Mc GEMMs (P1, P2) -> a per-candidate finish that applies the EI / UCB formulas of
// external/mathtoolbox/src/acquisition-functions.cpp:8-78.
#pragma once

#include "common.cuh"
#include "gram.cuh"

namespace slsgp
{
    // Kstar[i + m*ld] = k(X_i, xq_m), Gstar likewise with the x-gradient weight g. Rows i >= N are zero.
    // grid: (ld / 64, ceil(Mc / 64)).
    __global__ void __launch_bounds__(256)
        kstar_tile_kernel(const double* __restrict__ X, int N, int D, int ld, const double* __restrict__ Xq,
                          long long Mc, const double* __restrict__ theta, const double* __restrict__ inv_l,
                          int kernel_type, double se_xgrad_factor, double* __restrict__ Kstar,
                          double* __restrict__ Gstar)
    {
        __shared__ double sa[DCHUNK][TILE];
        __shared__ double sb[DCHUNK][TILE];
        const int         tm = blockIdx.x, tn = blockIdx.y;
        const int         tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
        double            r2[4][4];
        const int         nB = (int) min((long long) 1 << 30, Mc);
        tile_sq_dist(X, D, N, tm * TILE, Xq, D, nB, tn * TILE, D, inv_l, sa, sb, r2);
        const double a = theta[0];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const long long gm = (long long) tn * TILE + ty * 4 + j; // columns Mc .. Mp-1 are zero-filled
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const int gi = tm * TILE + tx * 4 + i;
                KernelVal kv;
                kv.k = 0.0, kv.g = 0.0;
                if (gi < N && gm < Mc) kv = kernel_value_and_xgrad_weight(kernel_type, a, r2[i][j], se_xgrad_factor);
                Kstar[(size_t) gi + (size_t) gm * ld] = kv.k;
                Gstar[(size_t) gi + (size_t) gm * ld] = kv.g;
            }
        }
    }

    // One warp per candidate column m. stats[m] = (mu, q = k.beta, ga, gb); overwrites Gstar with g o alpha and
    // Beta with g o beta (the operands of the P1 / P2 GEMMs).
    __global__ void __launch_bounds__(256)
        column_reduce_kernel(const double* __restrict__ Kstar, double* __restrict__ Gstar, double* __restrict__ Beta,
                             const double* __restrict__ alpha, int ld, long long Mc, double4* __restrict__ stats)
    {
        const long long m    = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        const int       lane = threadIdx.x & 31;
        if (m >= Mc) return;
        const double* kc = Kstar + (size_t) m * ld;
        double*       gc = Gstar + (size_t) m * ld;
        double*       bc = Beta + (size_t) m * ld;
        double        mu = 0.0, q = 0.0, ga = 0.0, gb = 0.0;
        for (int i = lane; i < ld; i += 32)
        {
            const double k = kc[i], g = gc[i], b = bc[i], al = alpha[i];
            mu             = fma(k, al, mu);
            q              = fma(k, b, q);
            const double t1 = g * al, t2 = g * b;
            ga += t1;
            gb += t2;
            gc[i] = t1;
            bc[i] = t2;
        }
        mu = warp_sum(mu), q = warp_sum(q), ga = warp_sum(ga), gb = warp_sum(gb);
        if (lane == 0) stats[m] = make_double4(mu, q, ga, gb);
    }

    // Standard normal pdf / cdf exactly as mathtoolbox probability-distributions.cpp:6-20 writes them.
    __device__ __forceinline__ double std_normal_pdf(double x)
    {
        return (1.0 / sqrt(2.0 * 3.14159265358979323846)) * exp(-0.5 * x * x);
    }
    __device__ __forceinline__ double std_normal_cdf(double x) { return 0.5 * (1.0 + erf(x / sqrt(2.0))); }

    struct SweepOut
    {
        double *mu, *sigma, *dmu, *dsigma, *val, *grad; // any may be null; dmu / dsigma / grad are D x M
    };

    // Two-tier precision of the tensor-core sweep: sweep_finish_kernel lists the candidates whose posterior variance is small
    // against the signal variance (sigma^2 < tau * a: the candidates next to data points, where sigma^2 = a - k.A.k and the
    // difference x_d gb - P2_d behind grad sigma cancel and fp16 / fp32 round-off is amplified by a / sigma^2), and run_sweep
    // re-evaluates exactly those in IEEE double afterwards.
    struct RefineList
    {
        int*       count;     // number of listed candidates (may exceed cap: only the first cap are kept); null = off
        long long* index;     // their positions in the job's candidate sequence
        long long  cap;
        long long  first;     // position of this shard's candidate 0
        double     threshold; // tau
        int        defer;     // arg-max jobs: listed candidates get val = NaN (never wins) until the second tier folds their value in
    };

    // Finish grad mu / grad sigma and apply the acquisition formulas. A block of 256 threads owns FINISH_CPB = 64 candidates:
    //   1. threads 0..63, one per candidate: sigma, Z, Phi, phi (erf / exp once per candidate), the value, the second-tier
    //      listing; meanwhile threads 64.. tabulate 1 / l_d^2;
    //   2. all 256 threads, one per (candidate, dimension): the D gradient entries of the block's candidates are 64 D consecutive
    //      doubles of P1 / P2 / Xq and of the outputs, so every access coalesces, the FP64 divisions of a candidate's D entries
    //      run in parallel, and there are four times as many warps per candidate to hide their latency (the one-thread-per-
    //      candidate form was latency-bound: 21-25 us for 19 or 38 thousand candidates alike).
    // P1, P2: ldp x Mc (rows 0..D-1 used). has_data == 0 reproduces the "regressor has no data" early return
    // (src/acquisition-function.cpp:176-179, 206-209).
    constexpr int FINISH_CPB = 64, FINISH_IL2 = 192;
    __global__ void __launch_bounds__(256)
        sweep_finish_kernel(const double* __restrict__ Xq, int D, long long Mc, const double4* __restrict__ stats,
                            const double* __restrict__ P1, const double* __restrict__ P2, int ldp,
                            const double* __restrict__ theta, const double* __restrict__ f_best_ptr, int acq_type,
                            double ucb_beta, SweepOut o, int n_parts = 0, long long part_stride = 0,
                            const double2* __restrict__ qx = nullptr, const double* __restrict__ P2x = nullptr, double x_shift = 0.0,
                            RefineList refine = RefineList{nullptr, nullptr, 0, 0, 0.0, 0})
    {
        __shared__ double s_ga[FINISH_CPB], s_gb[FINISH_CPB], s_sigma[FINISH_CPB], s_isig[FINISH_CPB], s_Z[FINISH_CPB], s_Phi[FINISH_CPB],
            s_phi[FINISH_CPB], s_diff[FINISH_CPB], s_il2[FINISH_IL2];
        __shared__ int  s_zero[FINISH_CPB];
        const long long m0 = (long long) blockIdx.x * FINISH_CPB;
        const int       t  = threadIdx.x;
        const long long m  = m0 + t;
        const bool      want_grad = o.dmu || o.dsigma || o.grad;
        if (t < FINISH_CPB && m < Mc)
        {
            double4 s = stats[m];
            for (int h = 0; h < n_parts; ++h) // the tensor sweep split the column blocks over several CTA groups: add their partial sums
            {
                const double2 e = qx[(size_t) h * part_stride + m];
                s.y += e.x, s.w += e.y;
            }
            const double  a      = theta[0];
            const double  mu     = s.x;
            const double  sig2   = a - s.y;
            const double  sigma  = sig2 < 0 ? 0.0 : sqrt(sig2); // src/preference-regressor.cpp:311-312
            const double  f_best = *f_best_ptr;
            bool deferred = false;
            if (refine.count && !(sig2 >= refine.threshold * a)) // also catches NaN
            {
                const int slot = atomicAdd(refine.count, 1);
                if (slot < refine.cap) refine.index[slot] = refine.first + m, deferred = refine.defer != 0;
            }
            if (o.mu) o.mu[m] = mu;
            if (o.sigma) o.sigma[m] = sigma;

            // mathtoolbox acquisition-functions.cpp:8-24 / :57-65
            const double diff = mu - f_best;
            const double Z    = diff / sigma;
            const double Phi = std_normal_cdf(Z), phi = std_normal_pdf(Z);
            if (o.val)
            {
                double v;
                if (acq_type == 1)
                    v = mu + ucb_beta * sigma;
                else
                {
                    const double EI = diff * Phi + sigma * phi;
                    v               = (sigma < 1e-16 || isnan(EI)) ? 0.0 : EI;
                }
                o.val[m] = deferred ? nan("") : v;
            }
            s_ga[t] = s.z, s_gb[t] = s.w, s_sigma[t] = sigma, s_isig[t] = 1.0 / sigma, s_Z[t] = Z, s_Phi[t] = Phi, s_phi[t] = phi, s_diff[t] = diff;
            s_zero[t] = (acq_type == 0 && sigma < 1e-16) ? 1 : 0;
        }
        else if (t >= FINISH_CPB && t - FINISH_CPB < min(D, FINISH_IL2))
        {
            const double l = theta[1 + t - FINISH_CPB];
            s_il2[t - FINISH_CPB] = 1.0 / (l * l);
        }
        if (!want_grad) return;
        __syncthreads();

        const int n_el = (int) min((long long) FINISH_CPB, Mc - m0) * D;
        const size_t base = (size_t) m0 * D;
        for (int e = t; e < n_el; e += 256)
        {
            const int       c = e / D, d = e - c * D;
            const long long mc = m0 + c;
            const double    sigma = s_sigma[c];
            double          il2;
            if (d < FINISH_IL2)
                il2 = s_il2[d];
            else
            {
                const double l = theta[1 + d];
                il2            = 1.0 / (l * l);
            }
            const double x   = Xq[base + e] - x_shift; // P1 / P2 were accumulated against X - x_shift
            const double dmu = (x * s_ga[c] - P1[(size_t) d + (size_t) mc * ldp]) * il2;
            double       p2  = P2[(size_t) d + (size_t) mc * ldp];
            for (int h = 0; h < n_parts; ++h) p2 += P2x[(size_t) h * ldp * part_stride + (size_t) d + (size_t) mc * ldp];
            const double dsg = -s_isig[c] * ((x * s_gb[c] - p2) * il2);
            if (o.dmu) o.dmu[base + e] = dmu;
            if (o.dsigma) o.dsigma[base + e] = dsg;
            if (o.grad)
            {
                double gr;
                if (acq_type == 1)
                    gr = dmu + ucb_beta * dsg; // :67-78
                else
                {
                    const double Z = s_Z[c], phi = s_phi[c], dphi = -Z * phi;
                    const double dZ = (dmu - Z * dsg) / sigma; // :26-55
                    gr              = dmu * s_Phi[c] + s_diff[c] * dZ * phi + dsg * phi + sigma * dZ * dphi;
                    if (isnan(gr)) s_zero[c] = 1; // benign race: every writer stores 1
                }
                o.grad[base + e] = gr;
            }
        }
        if (!(o.grad && acq_type == 0)) return;
        __syncthreads();
        // the reference returns the zero vector when sigma < 1e-16 or any entry is NaN
        for (int e = t; e < n_el; e += 256)
            if (s_zero[e / D]) o.grad[base + e] = 0.0;
    }

    // The acquisition formulas alone, for a posterior whose mean and deviation come from two different models
    // (Schonlau's batch criterion: objective_for_multiple_points, src/acquisition-function.cpp:63-110, takes mu from
    // the original regressor and sigma from one that already contains the pending points). Same arithmetic as the
    // tail of sweep_finish_kernel. One thread per candidate; dmu / dsigma / grad are D x M.
    __global__ void __launch_bounds__(256)
        acq_combine_kernel(const double* __restrict__ mu_in, const double* __restrict__ sigma_in,
                           const double* __restrict__ dmu_in, const double* __restrict__ dsigma_in, int D, long long M,
                           double f_best, int acq_type, double ucb_beta, double* __restrict__ val, double* __restrict__ grad)
    {
        const long long m = (long long) blockIdx.x * blockDim.x + threadIdx.x;
        if (m >= M) return;
        const double mu = mu_in[m], sigma = sigma_in[m];
        const double diff = mu - f_best, Z = diff / sigma;
        const double Phi = std_normal_cdf(Z), phi = std_normal_pdf(Z), dphi = -Z * phi;
        if (val)
        {
            const double EI = diff * Phi + sigma * phi;
            val[m]          = acq_type == 1 ? mu + ucb_beta * sigma : ((sigma < 1e-16 || isnan(EI)) ? 0.0 : EI);
        }
        if (!grad) return;
        bool has_nan = false;
        for (int d = 0; d < D; ++d)
        {
            const double dmu = dmu_in[(size_t) d + (size_t) m * D], dsg = dsigma_in[(size_t) d + (size_t) m * D];
            double       gr;
            if (acq_type == 1)
                gr = dmu + ucb_beta * dsg;
            else
            {
                const double dZ = (dmu - Z * dsg) / sigma;
                gr              = dmu * Phi + diff * dZ * phi + dsg * phi + sigma * dZ * dphi;
                has_nan |= isnan(gr);
            }
            grad[(size_t) d + (size_t) m * D] = gr;
        }
        if (acq_type == 0 && (sigma < 1e-16 || has_nan))
            for (int d = 0; d < D; ++d) grad[(size_t) d + (size_t) m * D] = 0.0;
    }

    // Refinement plumbing: candidates listed by RefineList are gathered into a compact D x R block (from a device array or from
    // the counter-based generator), swept in IEEE double, and the compact results are scattered back to their positions.
    __global__ void refine_gather_kernel(const long long* __restrict__ index, int R, int D, const double* __restrict__ Xq_all,
                                         int generate, uint64_t seed, long long first, double* __restrict__ Xr)
    {
        const int e = blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= R * D) return;
        const int       r = e / D, d = e - r * D;
        const long long m = index[r];
        Xr[e]             = generate ? candidate_coord(seed, first + m, d) : Xq_all[(size_t) d + (size_t) m * D];
    }
    // src: compact results (SweepOut over R candidates), dst: the job's outputs (device pointers, any may be null)
    __global__ void refine_scatter_kernel(const long long* __restrict__ index, int R, int D, SweepOut src, SweepOut dst)
    {
        const int e = blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= R * (D + 1)) return;
        const int       r = e / (D + 1), c = e - r * (D + 1);
        const long long m = index[r];
        if (c == D)
        {
            if (dst.mu) dst.mu[m] = src.mu[r];
            if (dst.sigma) dst.sigma[m] = src.sigma[r];
            if (dst.val) dst.val[m] = src.val[r];
        }
        else
        {
            if (dst.dmu) dst.dmu[(size_t) c + (size_t) m * D] = src.dmu[(size_t) c + (size_t) r * D];
            if (dst.dsigma) dst.dsigma[(size_t) c + (size_t) m * D] = src.dsigma[(size_t) c + (size_t) r * D];
            if (dst.grad) dst.grad[(size_t) c + (size_t) m * D] = src.grad[(size_t) c + (size_t) r * D];
        }
    }

    // Counter-based candidates: Xq[d + i*D] = candidate_coord(seed, first + i, d)
    __global__ void candidates_kernel(uint64_t seed, long long first, long long count, int D, double* __restrict__ Xq)
    {
        const long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= count * D) return;
        const long long i = e / D;
        const int       d = (int) (e - i * D);
        Xq[e]             = candidate_coord(seed, first + i, d);
    }

    // Arg-max of val[0..count) with lowest index winning ties; NaN never wins. Two-stage, deterministic.
    struct ArgMax
    {
        double    v;
        long long i;
    };
    __device__ __forceinline__ ArgMax argmax_combine(ArgMax a, ArgMax b)
    {
        if (b.i < 0) return a;
        if (a.i < 0) return b;
        if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
        return a;
    }
    __global__ void __launch_bounds__(256)
        argmax_partial_kernel(const double* __restrict__ val, long long count, long long index0, ArgMax* __restrict__ part)
    {
        __shared__ ArgMax sm[256];
        ArgMax            best;
        best.v = 0.0, best.i = -1;
        for (long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < count;
             e += (long long) gridDim.x * blockDim.x)
        {
            const double v = val[e];
            if (!isnan(v))
            {
                ArgMax c;
                c.v = v, c.i = index0 + e;
                best = argmax_combine(best, c);
            }
        }
        sm[threadIdx.x] = best;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) sm[threadIdx.x] = argmax_combine(sm[threadIdx.x], sm[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) part[blockIdx.x] = sm[0];
    }
    // Second tier of an arg-max job: fold the re-evaluated values of the listed candidates (val[r] belongs to candidate
    // index0 + index[r]) into the running best. Single block.
    __global__ void __launch_bounds__(256)
        argmax_indexed_kernel(const double* __restrict__ val, const long long* __restrict__ index, int R, long long index0, ArgMax* acc)
    {
        __shared__ ArgMax sm[256];
        ArgMax            best;
        best.v = 0.0, best.i = -1;
        for (int e = threadIdx.x; e < R; e += 256)
        {
            const double v = val[e];
            if (!isnan(v))
            {
                ArgMax c;
                c.v = v, c.i = index0 + index[e];
                best = argmax_combine(best, c);
            }
        }
        sm[threadIdx.x] = best;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) sm[threadIdx.x] = argmax_combine(sm[threadIdx.x], sm[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) acc[0] = argmax_combine(acc[0], sm[0]);
    }

    // Folds the partials of this shard into the running best (acc[0]); single block.
    __global__ void __launch_bounds__(256) argmax_final_kernel(const ArgMax* __restrict__ part, int n, ArgMax* acc)
    {
        __shared__ ArgMax sm[256];
        ArgMax            best;
        best.v = 0.0, best.i = -1;
        for (int e = threadIdx.x; e < n; e += 256) best = argmax_combine(best, part[e]);
        sm[threadIdx.x] = best;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) sm[threadIdx.x] = argmax_combine(sm[threadIdx.x], sm[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) acc[0] = argmax_combine(acc[0], sm[0]);
    }

    // f_i = (K_y alpha)_i - noise * alpha_i = mu(X_i); out = (max_i f_i, first arg max)
    // (Regressor::PredictMaximumPointFromData, src/regressor.cpp:29-43). Single block.
    __global__ void __launch_bounds__(256)
        fbest_kernel(const double* __restrict__ Kalpha, const double* __restrict__ alpha, double noise, int N,
                     double* __restrict__ f_best, int* __restrict__ index)
    {
        __shared__ ArgMax sm[256];
        ArgMax            best;
        best.v = 0.0, best.i = -1;
        for (int i = threadIdx.x; i < N; i += 256)
        {
            ArgMax c;
            c.v = Kalpha[i] - noise * alpha[i], c.i = i;
            best = argmax_combine(best, c);
        }
        sm[threadIdx.x] = best;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1)
        {
            if (threadIdx.x < o) sm[threadIdx.x] = argmax_combine(sm[threadIdx.x], sm[threadIdx.x + o]);
            __syncthreads();
        }
        if (threadIdx.x == 0) f_best[0] = sm[0].v, index[0] = (int) sm[0].i;
    }
} // namespace slsgp
