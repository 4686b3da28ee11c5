// libslsgp: C-ABI entry points (include/slsgp.h) and the device context behind them. sm_100a only.
//
// Data layout in HBM (all FP64, column-major):
//   X      D x N              observations, one per column (as Eigen stores the reference's m_X)
//   Xpad   Dp x ld            zero-padded copy, Dp = round_up(D, 64)            (A operand of the P1/P2 GEMMs)
//   XT1    ld x ldx           X^T in columns 0..D-1, ones in column D            (B operand of the MAP gradient GEMM)
//   K, L, W, Kinv, T          ld x ld each, ld = round_up(N, 64); rows/cols >= N hold an identity block so every
//                             kernel works on whole 64 x 64 tiles. L = chol(K) (lower), W = L^-1, Kinv = W^T W.
//   sweep workspace           Kstar, Gstar, Beta: ld x Mcap;  P1, P2: Dp x Mcap;  stats: Mcap x 4
#include "../../include/slsgp.h"

#include "common.cuh"
#include "chol.cuh"
#include "dense.cuh"
#include "gram.cuh"
#include "map.cuh"
#include "sweep.cuh"
#include "append.cuh"
#include "small.cuh"
#include "maximize.cuh"
#include "tc_sweep.cuh"
#include "tc_kstar.cuh"
#include "l1.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

using namespace slsgp;

namespace
{
    constexpr double kPi = 3.14159265358979323846;

    struct DevBuf
    {
        void*  p     = nullptr;
        size_t bytes = 0;
    };

    struct Phase
    {
        cudaEvent_t start = nullptr, stop = nullptr;
        bool        valid = false;
    };
} // namespace

struct slsgp_ctx
{
    int          device     = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // shard pipeline (run_sweep): candidates in + k* on `pre`, contraction + finish on `stream`, results out on `post`
    cudaStream_t pre_stream = nullptr, post_stream = nullptr;
    cudaEvent_t  ev_start = nullptr, ev_in[2] = {nullptr, nullptr}, ev_main[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // two-level Cholesky (do_factor, nb >= chol_two_level_from): critical-path stream, look-ahead stream, their events (lazy)
    cudaStream_t chol_hi = nullptr, chol_side = nullptr;
    cudaEvent_t  ev_chol[3] = {nullptr, nullptr, nullptr};
    int          chol_two_level_from = 96; // block columns (N >= 6144; below, the step chain hides the single-level update); SLSGP_CHOL_TWO_LEVEL_FROM
    int          chol_panel          = 8;  // block columns per panel; SLSGP_CHOL_PANEL
    bool         chol_look_ahead     = true; // SLSGP_CHOL_LOOKAHEAD=0: everything on the critical stream
    int          chol_switch_rem     = 48; // two-level only while more than this many block columns remain; SLSGP_CHOL_SWITCH_REM
    int          chol_pivots         = 2;  // pivots per barrier in the diagonal tile (2, or 1 for A/B); SLSGP_CHOL_PIVOTS
    std::string  err;
    unsigned     compat     = SLSGP_COMPAT_SE_XGRAD_2X;
    int          sweep_mode = SLSGP_SWEEP_FP64;
    uint64_t     launches   = 0;

    int    N = 0, D = 0, ld = 0, Dp = 0, ldx = 0;
    int    kernel_type = 0;
    double noise       = 0.0;
    double logdet_host = 0.0;
    bool   has_data = false, has_gram = false, has_factor = false, has_W = false, has_inverse = false,
         has_alpha = false;
    bool   factor_pending = false; // do_factor(defer_check): has_factor is provisional until the caller has read info

    std::vector<double> theta_host;
    std::vector<double> X_host; // host mirror of X (D x N), for slsgp_set_data_extend's prefix comparison

    DevBuf X, Xpad, XT1, theta, inv_l, K, L, W, Kinv, T, y, alpha, Kalpha, vec, scalars, info, fbest, fbest_idx, chol_flags;
    DevBuf pref_off, pref_idx, slot_off, slot_list, loglik, contrib, grad_y, Ymat, g_l;
    int    P = 0, pref_total = 0;
    DevBuf Xq, Kstar, Gstar, Beta, P1, P2, stats, o_mu, o_sigma, o_dmu, o_dsigma, o_val, o_grad, am_part, am_acc;
    DevBuf rf_count, rf_index, rf_X, rf_out; // two-tier precision of the tensor sweep (RefineList)
    double refine_tau = 0.1;                // candidates with sigma^2 < tau * a are re-evaluated in IEEE double; 0 = off
    uint64_t dense_version = ~0ull;         // model_version for which most candidates needed the second tier: swept in IEEE double directly
    long long Mcap = 0;

    // tensor-core sweep (SLSGP_SWEEP_TENSOR): fp16 operands + their TMA descriptors
    DevBuf      Bmat, Xt, Xs32, tcs, Ks, tc_err, comb, tc_qx, tc_P2x;
    DevBuf      mx_best, mx_X, mx_Xbest, mx_Gbest, mx_state, mx_val, mx_grad; // slsgp_acq_maximize
    CUtensorMap tmA, tmB;
    // k* generator on the tensor pipe (tc_kstar.cuh): split-fp16 coordinates of the observations (per model) and of the candidates
    // (per shard buffer), their squared norms, the TMA descriptors; kt_KP = 0 when D is too large for it (3 D > 128)
    DevBuf      kt_Xh, kt_nx, kt_Qh, kt_nq;
    CUtensorMap tmXh, tmQh;
    int         kt_KP = 0;
    int         ldt = 0, XP = 0;
    bool        tc_ready = false; // Bmat / Xt / Xs32 / scales match the current model
    long long   tc_Mcap = 0;      // rows of Ks (multiple of 128 * tc_ncta)
    int         tc_ncta = 2;      // CTAs per UMMA (1: cta_group::1, 2: CTA pair); SLSGP_TC_PAIR=0 selects 1

    // multi-GPU group (slsgp_ctx_create_multi): contexts on the other devices, owned by this one. The candidate range of
    // slsgp_acq_argmax / slsgp_acq_maximize and the batches of slsgp_posterior_batch / slsgp_acq_batch are split over the group.
    std::vector<slsgp_ctx*> peers;
    uint64_t                model_version = 0;   // bumped whenever (X, hyper-parameters, K^-1, alpha) change
    uint64_t                replica_of    = 0;   // peers: the primary's model_version this replica holds
    bool                    peer_access_tried = false;

    // pinned staging ring of the host-buffer entry points for PAGEABLE caller memory (run_sweep): two shard-sized slots each way
    double *pin_in = nullptr, *pin_out = nullptr;
    size_t  pin_in_bytes = 0, pin_out_bytes = 0;

    double* pinned       = nullptr; // small host staging area
    size_t  pinned_bytes = 0;

    std::map<std::string, Phase> phases;

    // per-kernel profiling (slsgp_profile_enable)
    bool profile = false;
    struct ProfRec
    {
        cudaEvent_t start, stop;
    };
    std::map<std::string, std::vector<ProfRec>> prof;
    std::vector<ProfRec>                        prof_free;
};

namespace
{
    // ---------------------------------------------------------------------------------------------------------
    // error plumbing
    // ---------------------------------------------------------------------------------------------------------
    slsgp_status fail(slsgp_ctx* c, slsgp_status s, const std::string& msg)
    {
        if (c) c->err = msg;
        return s;
    }

#define CUDA_TRY(expr)                                                                                             \
    do                                                                                                             \
    {                                                                                                              \
        cudaError_t e_ = (expr);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            return fail(ctx, SLSGP_ERR_CUDA,                                                                       \
                        std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +                  \
                            std::to_string(__LINE__) + ")");                                                       \
    } while (0)

#define TRY(expr)                                                                                                  \
    do                                                                                                             \
    {                                                                                                              \
        slsgp_status s_ = (expr);                                                                                  \
        if (s_ != SLSGP_OK) return s_;                                                                             \
    } while (0)

#define LAUNCH_CHECK()                                                                                             \
    do                                                                                                             \
    {                                                                                                              \
        ++ctx->launches;                                                                                           \
        CUDA_TRY(cudaGetLastError());                                                                              \
    } while (0)

    slsgp_status ensure(slsgp_ctx* ctx, DevBuf& b, size_t bytes)
    {
        if (b.bytes >= bytes && b.p) return SLSGP_OK;
        if (b.p) CUDA_TRY(cudaFree(b.p));
        b.p = nullptr, b.bytes = 0;
        // Buffers of small and mid-size models get twice what was asked for: an optimiser run grows its model by two or three
        // points per iteration, the leading dimension by 64 at a time, and every growth step used to free and re-allocate some
        // forty buffers in each pooled context (150 - 300 ms per crossing in the D = 64 loop, tools/config5_outliers.py). With
        // the headroom the N x N buffers are re-allocated at 128, 192, 320, 512, ... instead of at every multiple of 64.
        const size_t cap = bytes <= ((size_t) 64 << 20) ? 2 * bytes : bytes;
        cudaError_t  e   = cudaMalloc(&b.p, cap);
        if (e != cudaSuccess && cap != bytes)
        {
            cudaGetLastError();
            e = cudaMalloc(&b.p, bytes);
            if (e == cudaSuccess) b.bytes = bytes;
        }
        else if (e == cudaSuccess)
            b.bytes = cap;
        if (e != cudaSuccess)
        {
            cudaGetLastError();
            b.p = nullptr;
            return fail(ctx, SLSGP_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed");
        }
        return SLSGP_OK;
    }
    template <typename T> T* ptr(const DevBuf& b) { return static_cast<T*>(b.p); }
    double*                  dp(const DevBuf& b) { return static_cast<double*>(b.p); }

    slsgp_status phase_begin(slsgp_ctx* ctx, const char* name)
    {
        Phase& ph = ctx->phases[name];
        if (!ph.start)
        {
            CUDA_TRY(cudaEventCreate(&ph.start));
            CUDA_TRY(cudaEventCreate(&ph.stop));
        }
        ph.valid = false;
        CUDA_TRY(cudaEventRecord(ph.start, ctx->stream));
        return SLSGP_OK;
    }
    slsgp_status phase_end(slsgp_ctx* ctx, const char* name)
    {
        Phase& ph = ctx->phases[name];
        CUDA_TRY(cudaEventRecord(ph.stop, ctx->stream));
        ph.valid = true;
        return SLSGP_OK;
    }

    // Bracket the launches issued inside a scope with a pair of events (only while profiling is enabled).
    struct ProfScope
    {
        slsgp_ctx*         ctx;
        slsgp_ctx::ProfRec rec;
        bool               on;
        cudaStream_t       st;
        ProfScope(slsgp_ctx* c, const char* name, cudaStream_t stream = nullptr) : ctx(c), on(c->profile), st(stream ? stream : c->stream)
        {
            if (!on) return;
            if (!ctx->prof_free.empty())
            {
                rec = ctx->prof_free.back();
                ctx->prof_free.pop_back();
            }
            else if (cudaEventCreate(&rec.start) != cudaSuccess || cudaEventCreate(&rec.stop) != cudaSuccess)
            {
                on = false;
                return;
            }
            cudaEventRecord(rec.start, st);
            ctx->prof[name].push_back(rec);
        }
        ~ProfScope()
        {
            if (on) cudaEventRecord(rec.stop, st);
        }
    };

    // Strided device -> host copy of the leading n x n part of an ld x ld matrix.
    slsgp_status copy_matrix_out(slsgp_ctx* ctx, const DevBuf& src, double* dst, int n)
    {
        if (!dst) return SLSGP_OK;
        CUDA_TRY(cudaMemcpy2DAsync(dst, sizeof(double) * n, src.p, sizeof(double) * ctx->ld, sizeof(double) * n, n,
                                   cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    template <bool TA, bool TB> slsgp_status launch_gemm(slsgp_ctx* ctx, const GemmArgs& g, int batch = 1, cudaStream_t st = nullptr)
    {
        dim3 grid(g.m / TILE, g.n / TILE, batch);
        if (!st) st = ctx->stream;
        static const bool simt = std::getenv("SLSGP_FP64_GEMM") && std::string(std::getenv("SLSGP_FP64_GEMM")) == "simt";
        if (simt)
            gemm64_kernel<TA, TB><<<grid, 256, 0, st>>>(g);      // DFMA reference implementation
        else
            gemm64_dmma_kernel<TA, TB><<<grid, 256, 0, st>>>(g); // FP64 tensor pipe (mma.sync m8n8k4)
        LAUNCH_CHECK();
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // hyper-parameters -> device
    // ---------------------------------------------------------------------------------------------------------
    slsgp_status upload_theta(slsgp_ctx* ctx, const double* theta)
    {
        const int D = ctx->D;
        for (int i = 0; i <= D; ++i)
            if (!std::isfinite(theta[i])) return fail(ctx, SLSGP_ERR_NAN, "non-finite kernel hyper-parameter");
        ctx->theta_host.assign(theta, theta + D + 1);
        double* h = ctx->pinned;
        for (int i = 0; i <= D; ++i) h[i] = theta[i];
        for (int i = 0; i < D; ++i) h[D + 1 + i] = 1.0 / theta[1 + i];
        CUDA_TRY(cudaMemcpyAsync(ctx->theta.p, h, sizeof(double) * (D + 1), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->inv_l.p, h + D + 1, sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
        // the pinned staging area is reused by later calls: make sure the copies have consumed it
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    __global__ void pack_x_kernel(const double* __restrict__ X, int N, int D, int ld, int Dp, int ldx,
                                  double* __restrict__ Xpad, double* __restrict__ XT1)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x; // point (column of X), 0 .. ld-1
        if (i >= ld) return;
        for (int d = 0; d < Dp; ++d) Xpad[(size_t) d + (size_t) i * Dp] = (i < N && d < D) ? X[(size_t) d + (size_t) i * D] : 0.0;
        for (int t = 0; t < ldx; ++t)
        {
            double v = 0.0;
            if (i < N) v = (t < D) ? X[(size_t) t + (size_t) i * D] : (t == D ? 1.0 : 0.0);
            XT1[(size_t) i + (size_t) t * ld] = v;
        }
    }

    // device buffers of a model of ctx->N points in ctx->D dimensions (ctx->ld, Dp, ldx already set)
    slsgp_status ensure_model_buffers(slsgp_ctx* ctx)
    {
        const int    D = ctx->D;
        const size_t ld = ctx->ld, mat = sizeof(double) * ld * ld;
        TRY(ensure(ctx, ctx->X, sizeof(double) * (size_t) ctx->ld * D)); // room for slsgp_append_point up to ld points

        TRY(ensure(ctx, ctx->Xpad, sizeof(double) * (size_t) ctx->Dp * ld));
        TRY(ensure(ctx, ctx->XT1, sizeof(double) * ld * ctx->ldx));
        TRY(ensure(ctx, ctx->theta, sizeof(double) * (D + 1)));
        TRY(ensure(ctx, ctx->inv_l, sizeof(double) * D));
        TRY(ensure(ctx, ctx->K, mat));
        TRY(ensure(ctx, ctx->L, mat));
        TRY(ensure(ctx, ctx->W, mat));
        TRY(ensure(ctx, ctx->Kinv, mat));
        TRY(ensure(ctx, ctx->T, mat));
        TRY(ensure(ctx, ctx->y, sizeof(double) * ld));
        TRY(ensure(ctx, ctx->alpha, sizeof(double) * ld));
        TRY(ensure(ctx, ctx->Kalpha, sizeof(double) * ld));
        TRY(ensure(ctx, ctx->vec, sizeof(double) * ld));
        TRY(ensure(ctx, ctx->grad_y, sizeof(double) * ld));
        TRY(ensure(ctx, ctx->Ymat, sizeof(double) * ld * ctx->ldx * 8)); // up to 8 partial products (map_y_splits)
        TRY(ensure(ctx, ctx->g_l, sizeof(double) * D));
        TRY(ensure(ctx, ctx->scalars, sizeof(double) * 32));
        TRY(ensure(ctx, ctx->info, sizeof(int)));
        TRY(ensure(ctx, ctx->fbest, sizeof(double)));
        TRY(ensure(ctx, ctx->fbest_idx, sizeof(int)));
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // K1 / K2 / K3 on the context's current data
    // ---------------------------------------------------------------------------------------------------------
    slsgp_status do_gram(slsgp_ctx* ctx, int kernel_type, const double* theta, double noise)
    {
        if (!ctx->has_data) return fail(ctx, SLSGP_ERR_STATE, "slsgp_gram before slsgp_set_data");
        if (kernel_type != 0 && kernel_type != 1) return fail(ctx, SLSGP_ERR_INVALID, "unknown kernel_type");
        if (!std::isfinite(noise)) return fail(ctx, SLSGP_ERR_NAN, "non-finite noise level");
        TRY(upload_theta(ctx, theta));
        ctx->kernel_type = kernel_type, ctx->noise = noise;
        ctx->has_gram = ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = false;
        TRY(phase_begin(ctx, "gram"));
        const int nt = ctx->ld / TILE;
        ProfScope prof_scope(ctx, "gram");
        static const bool first_layout = std::getenv("SLSGP_GRAM_V1") && std::atoi(std::getenv("SLSGP_GRAM_V1")) != 0; // A/B switch
        if (first_layout)
            gram_tile_kernel<0><<<nt * (nt + 1) / 2, 256, 0, ctx->stream>>>(dp(ctx->X), ctx->N, ctx->D, ctx->ld, dp(ctx->theta), dp(ctx->inv_l), noise,
                                                                             kernel_type, dp(ctx->K), nullptr, nullptr);
        else if (kernel_type == 0)
            gram_sym_kernel<0><<<nt * (nt + 1) / 2, 256, 0, ctx->stream>>>(dp(ctx->X), ctx->N, ctx->D, ctx->ld, dp(ctx->theta), dp(ctx->inv_l), noise, dp(ctx->K));
        else
            gram_sym_kernel<1><<<nt * (nt + 1) / 2, 256, 0, ctx->stream>>>(dp(ctx->X), ctx->N, ctx->D, ctx->ld, dp(ctx->theta), dp(ctx->inv_l), noise, dp(ctx->K));
        LAUNCH_CHECK();
        TRY(phase_end(ctx, "gram"));
        ctx->has_gram = true;
        return SLSGP_OK;
    }

    // defer_check: enqueue only. The pivot check and the log-determinant stay on the device (info, scalars[8]) and the caller
    // collects them with its own results after ONE synchronisation (factor_pending; slsgp_map_objective_pref).
    slsgp_status do_factor(slsgp_ctx* ctx, double* logdet_out, bool defer_check = false)
    {
        if (!ctx->has_gram) return fail(ctx, SLSGP_ERR_STATE, "slsgp_factor before slsgp_gram");
        const int    ld = ctx->ld, nb = ld / TILE;
        ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = ctx->factor_pending = false;
        TRY(phase_begin(ctx, "factor"));
        // W is written tile by tile (diagonal tiles by the factorisation, the blocks below them by do_trtri); every reader
        // prunes its k-range to the lower blocks, so W needs no clearing
        double* L = dp(ctx->L);
        double* W = dp(ctx->W);
        TRY(ensure(ctx, ctx->chol_flags, sizeof(int) * (size_t) nb));
        // L = lower tiles of K, zero tiles above, flags and info cleared: one launch
        chol_prepare_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(dp(ctx->K), L, ld, nb, ptr<int>(ctx->chol_flags), ptr<int>(ctx->info));
        LAUNCH_CHECK();
        static bool chol_attr[64] = {}; // function attributes are per device
        if (!chol_attr[ctx->device & 63])
        {
            CUDA_TRY(cudaFuncSetAttribute(chol_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_BYTES));
            CUDA_TRY(cudaFuncSetAttribute(chol_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CHOL_SMEM_BYTES));
            chol_attr[ctx->device & 63] = true;
        }
        // one step: launch k finishes block column k+1 (diagonal factor + panel) while updating the trailing block columns
        // k+2 .. pe-1 (pe = nb: the whole trailing matrix)
        const auto step = [&](int k, int pe, int upd, cudaStream_t st) -> slsgp_status {
            const int rem  = nb - k - 1;
            int       grid = rem;
            if (upd)
            {
                if (pe >= nb)
                    grid = rem * (rem + 1) / 2;
                else
                    for (int c = k + 2; c < pe; ++c) grid += nb - c;
            }
            ProfScope           ps(ctx, "chol_step", st);
            cudaLaunchConfig_t  cfg = {};
            cudaLaunchAttribute attr[1];
            cfg.gridDim = dim3(grid), cfg.blockDim = dim3(256), cfg.dynamicSmemBytes = CHOL_SMEM_BYTES, cfg.stream = st;
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; // overlap this launch with the tail of the previous step
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr, cfg.numAttrs = 1;
            if (ctx->chol_pivots >= 2)
                CUDA_TRY(cudaLaunchKernelEx(&cfg, chol_step_kernel<2>, L, W, ld, k, nb, pe, upd, ptr<int>(ctx->chol_flags), ptr<int>(ctx->info)));
            else
                CUDA_TRY(cudaLaunchKernelEx(&cfg, chol_step_kernel<1>, L, W, ld, k, nb, pe, upd, ptr<int>(ctx->chol_flags), ptr<int>(ctx->info)));
            LAUNCH_CHECK();
            return SLSGP_OK;
        };
        if (nb < ctx->chol_two_level_from)
        {
            for (int k = -1; k <= nb - 2; ++k) TRY(step(k, nb, k >= 0, ctx->stream));
        }
        else
        {
            // Two-level form for large N. The right-looking sweep above moves the whole trailing matrix through L2 once per
            // 64 columns (4 flop per byte); here a PANEL of pw block columns is factored with the same step kernel restricted to
            // the panel, and the trailing matrix receives the panel in ONE rank-(64 pw) update C -= P P^T (gemm64_dmma_kernel).
            // Look-ahead: the update of the NEXT panel's columns runs on the critical stream, the rest of the trailing matrix on a
            // second, lower-priority stream under the next panel's factorisation:
            //   hi  : steps(p) . [ev_panel] . wait(ev_side of p-1) . update(next panel's columns)            . steps(p+1) ...
            //   side:              wait(ev_panel) . update(columns beyond the next panel) . [ev_side]
            // (both updates of panel p-1 / p that touch the same columns are ordered by ev_side).
            if (!ctx->chol_hi)
            {
                int lo = 0, hi = 0;
                CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                CUDA_TRY(cudaStreamCreateWithPriority(&ctx->chol_hi, cudaStreamNonBlocking, hi));
                CUDA_TRY(cudaStreamCreateWithPriority(&ctx->chol_side, cudaStreamNonBlocking, lo));
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_chol[0], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_chol[1], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_chol[2], cudaEventDisableTiming));
            }
            cudaStream_t hi = ctx->chol_hi, side = ctx->chol_look_ahead ? ctx->chol_side : ctx->chol_hi;
            cudaEvent_t  ev_panel = ctx->ev_chol[0], ev_side = ctx->ev_chol[1], ev_done = ctx->ev_chol[2];
            CUDA_TRY(cudaEventRecord(ev_done, ctx->stream)); // L = K, cleared flags
            CUDA_TRY(cudaStreamWaitEvent(hi, ev_done, 0));
            if (side != hi) CUDA_TRY(cudaStreamWaitEvent(side, ev_done, 0));
            const int PB = ctx->chol_panel;
            const auto syrk = [&](int c0, int pw, int r0, int ncols, cudaStream_t st) -> slsgp_status {
                // C[r0.., r0 .. r0+ncols) -= P[r0.., :] P[r0 .. r0+ncols, :]^T, P = block columns c0 .. c0+pw-1 of L, lower tiles only
                ProfScope ps(ctx, "chol_syrk", st);
                const double* P = L + (size_t) r0 * TILE + (size_t) c0 * TILE * ld;
                GemmArgs      g = gemm_args(P, P, L + (size_t) r0 * TILE * ((size_t) ld + 1), (nb - r0) * TILE, ncols * TILE, pw * TILE, ld, ld, ld, -1.0, 1.0);
                g.lower_only    = 1;
                return launch_gemm<false, true>(ctx, g, 1, st);
            };
            bool side_pending = false;
            for (int c0 = 0; c0 < nb; c0 += PB)
            {
                if (nb - c0 <= ctx->chol_switch_rem)
                {
                    // the rest is chain-bound: the single-level sweep hides its (small) trailing updates behind the diagonal tiles
                    // and pays no per-panel launches. Its first step factors column c0 as it stands (every panel is in already).
                    if (side_pending)
                    {
                        CUDA_TRY(cudaStreamWaitEvent(hi, ev_side, 0));
                        side_pending = false;
                    }
                    TRY(step(c0 - 1, nb, 0, hi));
                    for (int k = c0; k <= nb - 2; ++k) TRY(step(k, nb, 1, hi));
                    break;
                }
                const int pw = std::min(PB, nb - c0), pe = c0 + pw;
                TRY(step(c0 - 1, pe, 0, hi));
                for (int k = c0; k <= pe - 2; ++k) TRY(step(k, pe, 1, hi));
                if (pe >= nb) break;
                const int pw2 = std::min(PB, nb - pe), r1 = pe + pw2;
                if (side != hi && r1 < nb) CUDA_TRY(cudaEventRecord(ev_panel, hi));
                if (side_pending)
                {
                    CUDA_TRY(cudaStreamWaitEvent(hi, ev_side, 0));
                    side_pending = false;
                }
                TRY(syrk(c0, pw, pe, pw2, hi));
                if (r1 < nb)
                {
                    if (side != hi) CUDA_TRY(cudaStreamWaitEvent(side, ev_panel, 0));
                    TRY(syrk(c0, pw, r1, nb - r1, side));
                    if (side != hi)
                    {
                        CUDA_TRY(cudaEventRecord(ev_side, side));
                        side_pending = true;
                    }
                }
            }
            if (side_pending) CUDA_TRY(cudaStreamWaitEvent(hi, ev_side, 0));
            CUDA_TRY(cudaEventRecord(ev_done, hi));
            CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev_done, 0));
        }
        logdet_kernel<<<1, 256, 0, ctx->stream>>>(L, ctx->N, ld, dp(ctx->scalars) + 8);
        LAUNCH_CHECK();
        TRY(phase_end(ctx, "factor"));
        if (defer_check)
        {
            ctx->has_factor = ctx->factor_pending = true;
            return SLSGP_OK;
        }
        int    info = 0;
        double logdet = 0.0;
        CUDA_TRY(cudaMemcpyAsync(&info, ctx->info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(&logdet, dp(ctx->scalars) + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (info != 0)
            return fail(ctx, SLSGP_ERR_NOT_SPD,
                        "Cholesky: non-positive pivot at index " + std::to_string(info - 1) + " (K_y is not SPD)");
        ctx->logdet_host = logdet;
        if (logdet_out) *logdet_out = logdet;
        ctx->has_factor = true;
        return SLSGP_OK;
    }

    // W = L^-1 by recursive doubling over the already-inverted 64 x 64 diagonal blocks:
    //   [L11 0; L21 L22]^-1 = [W11 0; -W22 L21 W11, W22]
    slsgp_status do_trtri(slsgp_ctx* ctx)
    {
        if (ctx->has_W) return SLSGP_OK;
        if (!ctx->has_factor) return fail(ctx, SLSGP_ERR_STATE, "inverse requested before slsgp_factor");
        const int ld = ctx->ld;
        double *  L = dp(ctx->L), *W = dp(ctx->W), *T = dp(ctx->T);
        for (int s = TILE; s < ld; s *= 2)
        {
            const int       npairs = (ld + 2 * s - 1) / (2 * s);
            const long long stride = 2LL * s * ((long long) ld + 1);
            GemmArgs        a      = gemm_args(L + s, W, T + s, s, s, s, ld, ld, ld, 1.0, 0.0); // T = L21 W11
            a.sA = a.sB = a.sC = stride;
            a.k_lo_mode        = 1;
            a.row0 = s, a.row_step = 2 * s, a.row_limit = ld;
            TRY((launch_gemm<false, false>(ctx, a, npairs)));
            GemmArgs b = gemm_args(W + (size_t) s * ((size_t) ld + 1), T + s, W + s, s, s, s, ld, ld, ld, -1.0, 0.0);
            b.sA = b.sB = b.sC = stride; // W21 = -W22 T
            b.k_hi_mode        = 1;
            b.row0 = s, b.row_step = 2 * s, b.row_limit = ld;
            TRY((launch_gemm<false, false>(ctx, b, npairs)));
        }
        ctx->has_W = true;
        return SLSGP_OK;
    }

    slsgp_status do_inverse(slsgp_ctx* ctx)
    {
        if (ctx->has_inverse) return SLSGP_OK;
        if (!ctx->has_factor) return fail(ctx, SLSGP_ERR_STATE, "slsgp_inverse before slsgp_factor");
        TRY(phase_begin(ctx, "inverse"));
        TRY(do_trtri(ctx));
        const int ld = ctx->ld;
        GemmArgs  g  = gemm_args(dp(ctx->W), dp(ctx->W), dp(ctx->Kinv), ld, ld, ld, ld, ld, ld, 1.0, 0.0);
        g.lower_only = 1, g.k_lo_mode = 2; // Kinv = W^T W, lower tiles, k >= max(tile row, tile col)
        TRY((launch_gemm<true, false>(ctx, g)));
        symmetrize_kernel<<<dim3(ld / TILE, ld / TILE), 256, 0, ctx->stream>>>(dp(ctx->Kinv), ld);
        LAUNCH_CHECK();
        TRY(phase_end(ctx, "inverse"));
        ctx->has_inverse = true;
        return SLSGP_OK;
    }

    // alpha = Kinv y (y already on the device in ctx->y), then f_best.
    slsgp_status do_alpha(slsgp_ctx* ctx)
    {
        TRY(do_inverse(ctx));
        ctx->tc_ready = false;
        TRY(phase_begin(ctx, "alpha"));
        const int ld = ctx->ld, blocks = (ld * 32 + 255) / 256;
        gemv_kernel<true><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->Kinv), ld, ld, dp(ctx->y), dp(ctx->alpha), 0);
        LAUNCH_CHECK();
        gemv_kernel<true><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->K), ld, ld, dp(ctx->alpha), dp(ctx->Kalpha), 0);
        LAUNCH_CHECK();
        fbest_kernel<<<1, 256, 0, ctx->stream>>>(dp(ctx->Kalpha), dp(ctx->alpha), ctx->noise, ctx->N, dp(ctx->fbest),
                                                 ptr<int>(ctx->fbest_idx));
        LAUNCH_CHECK();
        TRY(phase_end(ctx, "alpha"));
        ctx->has_alpha = true;
        ++ctx->model_version;
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // K4: one shard of Mc <= Mcap candidates already in d_Xq (D x Mc, device). Outputs are device pointers.
    // ---------------------------------------------------------------------------------------------------------
    long long tc_shard_cap();
    bool      is_tensor_mode(int mode);
    slsgp_status prepare_tensor(slsgp_ctx* ctx);
    slsgp_status tensor_map_2d(slsgp_ctx* ctx, CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

    // Per-shard scratch. FP64 mode: Kstar / Gstar / Beta (ld x cap each) dominate, so cap <= 16384 candidates;
    // tensor mode: the fp16 Ks operand (cap x ldt), cap = SLSGP_TC_SHARD (default two waves of 148 x 128).
    slsgp_status ensure_sweep_workspace(slsgp_ctx* ctx, long long want)
    {
        const bool tensor = is_tensor_mode(ctx->sweep_mode);
        if (tensor) TRY(prepare_tensor(ctx));
        long long cap = std::min<long long>(std::max<long long>(want, 64), tensor ? tc_shard_cap() : 16384);
        cap           = round_up64(cap, tensor ? TC_BM * ctx->tc_ncta : TILE);
        if (tensor && ctx->tc_Mcap < cap)
        {
            // two shard buffers, each: rows [0, cap) k16, [cap, 2 cap) its fp16 rounding residual (read only by the split-precision
            // passes), [2 cap, 3 cap) / [3 cap, 4 cap) the Matern gradient weight g16 and its residual
            TRY(ensure(ctx, ctx->Ks, sizeof(__half) * 8 * (size_t) cap * ctx->ldt));
            TRY(tensor_map_2d(ctx, &ctx->tmA, ctx->Ks.p, 8 * (uint64_t) cap, (uint64_t) ctx->ldt, TC_BM));
            if (ctx->kt_KP)
            {
                TRY(ensure(ctx, ctx->kt_Qh, sizeof(__half) * 2 * (size_t) cap * ctx->kt_KP));
                TRY(ensure(ctx, ctx->kt_nq, sizeof(float) * 2 * (size_t) cap));
                TRY(tensor_map_2d(ctx, &ctx->tmQh, ctx->kt_Qh.p, 2 * (uint64_t) cap, (uint64_t) ctx->kt_KP, KT_BM));
            }
            ctx->tc_Mcap = cap;
        }
        if (ctx->Mcap >= cap) return SLSGP_OK;
        const size_t col = sizeof(double) * (size_t) ctx->ld;
        if (!tensor)
        {
            TRY(ensure(ctx, ctx->Kstar, col * cap));
            TRY(ensure(ctx, ctx->Gstar, col * cap));
            TRY(ensure(ctx, ctx->Beta, col * cap));
        }
        TRY(ensure(ctx, ctx->P1, sizeof(double) * (size_t) ctx->Dp * cap));
        TRY(ensure(ctx, ctx->P2, sizeof(double) * (size_t) ctx->Dp * cap));
        TRY(ensure(ctx, ctx->stats, sizeof(double4) * (size_t) cap));
        if (tensor)
        {
            TRY(ensure(ctx, ctx->tc_qx, 3 * sizeof(double2) * (size_t) cap)); // partial sums of up to 3 extra column-block groups
            TRY(ensure(ctx, ctx->tc_P2x, 3 * sizeof(double) * (size_t) ctx->Dp * cap));
        }
        // candidate and result staging: two shard buffers each (run_sweep)
        TRY(ensure(ctx, ctx->Xq, 2 * sizeof(double) * (size_t) ctx->D * cap));
        TRY(ensure(ctx, ctx->o_mu, 2 * sizeof(double) * cap));
        TRY(ensure(ctx, ctx->o_sigma, 2 * sizeof(double) * cap));
        TRY(ensure(ctx, ctx->o_val, 2 * sizeof(double) * cap));
        TRY(ensure(ctx, ctx->o_dmu, 2 * sizeof(double) * (size_t) ctx->D * cap));
        TRY(ensure(ctx, ctx->o_dsigma, 2 * sizeof(double) * (size_t) ctx->D * cap));
        TRY(ensure(ctx, ctx->o_grad, 2 * sizeof(double) * (size_t) ctx->D * cap));
        TRY(ensure(ctx, ctx->am_part, sizeof(ArgMax) * 1024));
        TRY(ensure(ctx, ctx->am_acc, sizeof(ArgMax)));
        ctx->Mcap = cap;
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // K4, tensor-core path (tc_sweep.cuh)
    // ---------------------------------------------------------------------------------------------------------
    bool is_tensor_mode(int mode) { return mode == SLSGP_SWEEP_TENSOR || mode == SLSGP_SWEEP_TENSOR_X2 || mode == SLSGP_SWEEP_TENSOR_X1; }
    int  tensor_passes(int mode) { return mode == SLSGP_SWEEP_TENSOR ? 3 : (mode == SLSGP_SWEEP_TENSOR_X2 ? 2 : 1); }

    int tc_xp(int D) // epilogue register-tile width: smallest instantiated XP >= D + 1
    {
        const int want = D + 1;
        for (int xp : {8, 12, 20, 36, 68})
            if (want <= xp) return xp;
        return 0;
    }

    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

    // fp16 row-major [rows x cols] matrix, boxes of box_rows x 64 elements (128 bytes), 128-byte swizzle.
    slsgp_status tensor_map_2d(slsgp_ctx* ctx, CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows)
    {
        static EncodeTiledFn encode = nullptr;
        if (!encode)
        {
            void*                           fn = nullptr;
            cudaDriverEntryPointQueryResult qr;
            CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
            if (qr != cudaDriverEntryPointSuccess || !fn)
                return fail(ctx, SLSGP_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
            encode = reinterpret_cast<EncodeTiledFn>(fn);
        }
        const cuuint64_t dims[2]    = {cols, rows};
        const cuuint64_t strides[1] = {cols * sizeof(__half)};
        const cuuint32_t box[2]     = {(cuuint32_t) TC_BK, box_rows};
        const cuuint32_t estr[2]    = {1, 1};
        const CUresult   r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(ctx, SLSGP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int) r));
        return SLSGP_OK;
    }

    // fp16 operands derived from the fitted model (Kinv, alpha, X, theta): once per model.
    slsgp_status prepare_tensor(slsgp_ctx* ctx)
    {
        if (ctx->tc_ready) return SLSGP_OK;
        const bool matern = ctx->kernel_type == SLSGP_KERNEL_ARD_MATERN52;
        if (matern && ctx->tc_ncta != 2) return fail(ctx, SLSGP_ERR_INVALID, "the Matern tensor sweep needs the CTA-pair kernel (SLSGP_TC_PAIR=1)");
        const int XP = tc_xp(ctx->D + (matern ? 1 : 0)); // Matern carries one more reduction column (gb = sum g_i u_i)
        if (!XP) return fail(ctx, SLSGP_ERR_INVALID, "SLSGP_SWEEP_TENSOR supports D <= 67 (66 for the Matern kernel)");
        const int ldt = round_up(ctx->N, TC_BN);
        ctx->XP = XP, ctx->ldt = ldt;
        const size_t brows = 2 * (size_t) ldt + TC_BN;
        TRY(ensure(ctx, ctx->Bmat, sizeof(__half) * brows * ldt));
        TRY(ensure(ctx, ctx->Xt, sizeof(float) * (size_t) ldt * XP));
        TRY(ensure(ctx, ctx->Xs32, sizeof(float) * (size_t) ldt * ctx->D));
        TRY(ensure(ctx, ctx->tcs, sizeof(TcScales)));
        if (!ctx->tc_err.p)
        {
            TRY(ensure(ctx, ctx->tc_err, sizeof(int)));
            CUDA_TRY(cudaMemsetAsync(ctx->tc_err.p, 0, sizeof(int), ctx->stream));
        }
        tc_scales_kernel<<<1, 256, 0, ctx->stream>>>(dp(ctx->Kinv), ctx->ld, ctx->N, dp(ctx->alpha), dp(ctx->X), ctx->D,
                                                     dp(ctx->theta), ptr<TcScales>(ctx->tcs));
        LAUNCH_CHECK();
        tc_pack_b_kernel<<<dim3(ldt / 256, (unsigned) brows), 256, 0, ctx->stream>>>(
            dp(ctx->Kinv), ctx->ld, ctx->N, ctx->D, XP, ldt, dp(ctx->alpha), dp(ctx->X), ptr<TcScales>(ctx->tcs),
            ptr<__half>(ctx->Bmat));
        LAUNCH_CHECK();
        tc_pack_x_kernel<<<(ldt + 255) / 256, 256, 0, ctx->stream>>>(dp(ctx->X), ctx->N, ctx->D, XP, ldt, dp(ctx->inv_l), matern ? 2 : 1,
                                                                    ptr<float>(ctx->Xt), ptr<float>(ctx->Xs32));
        LAUNCH_CHECK();
        TRY(tensor_map_2d(ctx, &ctx->tmB, ctx->Bmat.p, brows, (uint64_t) ldt, TC_BN / ctx->tc_ncta));
        // operands of the tensor-pipe k* generator (tc_kstar.cuh): up to three 64-wide K slices, i.e. D <= 64
        const int KP = kt_kp(ctx->D);
        if (KP != ctx->kt_KP) ctx->tc_Mcap = 0; // the per-shard candidate operand has another width
        ctx->kt_KP = KP <= 192 ? KP : 0;
        if (ctx->kt_KP)
        {
            TRY(ensure(ctx, ctx->kt_Xh, sizeof(__half) * (size_t) ldt * KP));
            TRY(ensure(ctx, ctx->kt_nx, sizeof(float) * (size_t) ldt));
            tc_pack_xh_kernel<<<(ldt + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->X), ctx->N, ctx->D, ldt, KP, dp(ctx->inv_l), ptr<__half>(ctx->kt_Xh),
                                                                         ptr<float>(ctx->kt_nx));
            LAUNCH_CHECK();
            TRY(tensor_map_2d(ctx, &ctx->tmXh, ctx->kt_Xh.p, (uint64_t) ldt, (uint64_t) KP, KT_BN));
        }
        ctx->tc_ready = true;
        return SLSGP_OK;
    }

    long long tc_shard_cap()
    {
        static long long cap = 0;
        if (!cap)
        {
            const char* e = std::getenv("SLSGP_TC_SHARD");
            cap           = e ? std::atoll(e) : 148LL * TC_BM * 2;
            cap           = std::max<long long>(TC_BM, round_up64(cap, TC_BM));
        }
        return cap;
    }

    template <int XP, int NCTA, bool MATERN> slsgp_status launch_tc_gemm_n(slsgp_ctx* ctx, const TcGemmParams& prm_in, int n_sm)
    {
        TcGemmParams prm   = prm_in;
        const size_t fixed = (size_t) tc_xs_cols(XP) * XP * sizeof(float) + 1024;
        prm.stage_bytes    = tc_stage_bytes(prm.passes, NCTA);
        prm.stages         = (int) std::min<size_t>(TC_MAX_STAGES, (232448 - 512 - fixed) / prm.stage_bytes);
        if (prm.stages < 2) return fail(ctx, SLSGP_ERR_INVALID, "tensor sweep: pipeline does not fit in shared memory");
        const size_t smem  = (size_t) prm.stages * prm.stage_bytes + fixed;
        static size_t attr_smem_dev[64] = {}; // function attributes are per device
        size_t&       attr_smem = attr_smem_dev[ctx->device & 63];
        if (attr_smem < smem)
        {
            CUDA_TRY(cudaFuncSetAttribute(tc_sweep_gemm_kernel<XP, NCTA, MATERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            attr_smem = smem;
        }
        const int groups = prm.n_cand_blocks / NCTA * prm.split; // work items: candidate groups x column-block halves
        const int grid   = std::min(groups, n_sm / NCTA) * NCTA;
        if (NCTA == 1)
            tc_sweep_gemm_kernel<XP, NCTA, MATERN><<<grid, TC_THREADS, smem, ctx->stream>>>(ctx->tmA, ctx->tmB, prm);
        else
        {
            cudaLaunchConfig_t  cfg = {};
            cudaLaunchAttribute attr[1];
            cfg.gridDim = dim3(grid), cfg.blockDim = dim3(TC_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = ctx->stream;
            attr[0].id               = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = NCTA, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr, cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_sweep_gemm_kernel<XP, NCTA, MATERN>, ctx->tmA, ctx->tmB, prm));
        }
        LAUNCH_CHECK();
        return SLSGP_OK;
    }

    template <int XP> slsgp_status launch_tc_gemm(slsgp_ctx* ctx, const TcGemmParams& prm, int n_sm)
    {
        if (prm.matern) return launch_tc_gemm_n<XP, 2, true>(ctx, prm, n_sm); // CTA-pair kernel only (prepare_tensor checks)
        return ctx->tc_ncta == 2 ? launch_tc_gemm_n<XP, 2, false>(ctx, prm, n_sm) : launch_tc_gemm_n<XP, 1, false>(ctx, prm, n_sm);
    }

    slsgp_status sweep_finish(slsgp_ctx* ctx, int acq_type, double ucb_beta, const double* d_Xq, long long Mc, SweepOut out,
                              int n_parts = 0, double x_shift = 0.0, int ldp = 0, RefineList refine = RefineList{nullptr, nullptr, 0, 0, 0.0, 0})
    {
        ProfScope ps(ctx, "sweep_finish");
        sweep_finish_kernel<<<(unsigned) ((Mc + FINISH_CPB - 1) / FINISH_CPB), 256, 0, ctx->stream>>>(
            d_Xq, ctx->D, Mc, ptr<double4>(ctx->stats), dp(ctx->P1), dp(ctx->P2), ldp ? ldp : ctx->Dp, dp(ctx->theta), dp(ctx->fbest),
            acq_type, ucb_beta, out, n_parts, ctx->Mcap, ptr<double2>(ctx->tc_qx), dp(ctx->tc_P2x), x_shift, refine);
        LAUNCH_CHECK();
        return SLSGP_OK;
    }

    // k* operand of one shard into shard buffer `buf`, on `st`. under_gemm: the launch shares the SMs with the persistent
    // contraction kernel of the previous shard (one CTA per SM, see kstar16_strip_kernel); else it has the GPU to itself.
    slsgp_status tensor_kstar(slsgp_ctx* ctx, int buf, const double* d_Xq, long long Mc, cudaStream_t st, bool under_gemm)
    {
        const int       D = ctx->D, ldt = ctx->ldt, passes = tensor_passes(ctx->sweep_mode);
        const long long Mpad = round_up64(Mc, TC_BM * ctx->tc_ncta);
        __half*         Ks    = ptr<__half>(ctx->Ks) + (size_t) buf * 4 * ctx->tc_Mcap * ldt;
        __half*         Ks_lo = passes > 1 ? Ks + (size_t) ctx->tc_Mcap * ldt : nullptr;
        __half*         Gs    = Ks + (size_t) 2 * ctx->tc_Mcap * ldt; // Matern only
        __half*         Gs_lo = passes > 1 ? Ks + (size_t) 3 * ctx->tc_Mcap * ldt : nullptr;
        ProfScope       ps(ctx, "tc_kstar", st);
        // The strip form needs >= 4 strips of 64 candidates per SM to fill the machine (one strip per CTA at a time); below that
        // (shards under ~38 thousand candidates: the last shard of a job, the starts of slsgp_acq_maximize, small batches) the
        // tiled form has 16 times as many CTAs and is the faster one (9472 candidates, N = 2048: 39 vs 56 us; 18944: 71 vs 76 us).
        // SLSGP_KSTAR_V1=1 / 0 forces the tiled / strip form (A/B).
        static const int  tiled_env = std::getenv("SLSGP_KSTAR_V1") ? std::atoi(std::getenv("SLSGP_KSTAR_V1")) : -1;
        int               n_sm_k    = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm_k, cudaDevAttrMultiProcessorCount, ctx->device));
        const bool        tiled = tiled_env >= 0 ? tiled_env != 0 : (!under_gemm && Mpad / 64 < 4LL * n_sm_k);
        static bool       kstar_attr_dev[64] = {};
        bool&             kstar_attr = kstar_attr_dev[ctx->device & 63];
        if (!kstar_attr) // D > 62 needs more than the 48 KB a kernel gets without opting in
        {
            const int big = (int) (sizeof(float) * (64 * 68 + 67 * 128));
            CUDA_TRY(cudaFuncSetAttribute(kstar16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            CUDA_TRY(cudaFuncSetAttribute(kstar16_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            CUDA_TRY(cudaFuncSetAttribute(kstar16_strip_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            CUDA_TRY(cudaFuncSetAttribute(kstar16_strip_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            kstar_attr = true;
        }
        // Squared distances on the tensor pipe (tc_kstar.cuh) for D <= 64 when the generator has the GPU to itself (it allocates all of
        // an SM's TMEM, as the contraction kernel does); SLSGP_KSTAR_TC=0 selects the FP32-pipe generators below (A/B).
        static const int tc_env = std::getenv("SLSGP_KSTAR_TC") ? std::atoi(std::getenv("SLSGP_KSTAR_TC")) : 1;
        // (from ~1000 observations on: below that a shard is a few tiles per SM and the FP32-pipe generators, whose cost shrinks
        // with N, are the faster ones - N = 115, D = 64: 0.066 vs 0.101 ms per shard; SLSGP_KSTAR_TC=2 forces the tensor-pipe form)
        if (tc_env != 0 && tiled_env < 0 && ctx->kt_KP && !under_gemm && (ldt >= 1024 || tc_env == 2))
        {
            const int       KP = ctx->kt_KP, KS = KP / TC_BK;
            const long long Mp128 = Mpad; // every row the contraction reads (a multiple of 128 or, for the CTA pair, 256)
            __half*         Qh = ptr<__half>(ctx->kt_Qh) + (size_t) buf * ctx->tc_Mcap * KP;
            float*          nq = ptr<float>(ctx->kt_nq) + (size_t) buf * ctx->tc_Mcap;
            tc_pack_qh_kernel<<<(unsigned) ((Mp128 * 32 + 255) / 256), 256, 0, st>>>(d_Xq, Mc, Mp128, D, KP, dp(ctx->inv_l), Qh, nq);
            LAUNCH_CHECK();
            KstarTcParams prm;
            prm.ldt = ldt, prm.ncb = ldt / KT_BN, prm.n_strips = (int) (Mp128 / KT_BM), prm.q_row0 = (int) ((long long) buf * ctx->tc_Mcap);
            prm.stages = KS == 1 ? 3 : 1; // 144 KB of operand tiles either way (KS = 2: 96 KB)
            prm.nq = nq, prm.nx = ptr<float>(ctx->kt_nx), prm.sc = ptr<TcScales>(ctx->tcs);
            prm.Ks = Ks, prm.Ks_lo = Ks_lo, prm.Gs = Gs, prm.Gs_lo = Gs_lo, prm.err = ptr<int>(ctx->tc_err);
            const size_t smem = (size_t) prm.stages * KS * (KT_A_BYTES + KT_B_BYTES) + 1024 + KT_EPI_WARPS * 32 * 64 + KT_NX_SMEM * sizeof(float); // stages, alignment, transposing buffers, |x|^2
            const int    grid = std::min(prm.n_strips * prm.ncb, n_sm_k);
            static bool  kt_attr_dev[64] = {};
            bool&        kt_attr = kt_attr_dev[ctx->device & 63];
            if (!kt_attr)
            {
                const int big = 3 * (KT_A_BYTES + KT_B_BYTES) + 1024 + KT_EPI_WARPS * 32 * 64 + KT_NX_SMEM * (int) sizeof(float);
                CUDA_TRY(cudaFuncSetAttribute(kstar_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                CUDA_TRY(cudaFuncSetAttribute(kstar_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                CUDA_TRY(cudaFuncSetAttribute(kstar_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                CUDA_TRY(cudaFuncSetAttribute(kstar_tc_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                CUDA_TRY(cudaFuncSetAttribute(kstar_tc_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                CUDA_TRY(cudaFuncSetAttribute(kstar_tc_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
                kt_attr = true;
            }
            if (ctx->kernel_type == 0)
            {
                if (KS == 1)
                    kstar_tc_kernel<0, 1><<<grid, KT_THREADS, smem, st>>>(ctx->tmQh, ctx->tmXh, prm);
                else if (KS == 2)
                    kstar_tc_kernel<0, 2><<<grid, KT_THREADS, smem, st>>>(ctx->tmQh, ctx->tmXh, prm);
                else
                    kstar_tc_kernel<0, 3><<<grid, KT_THREADS, smem, st>>>(ctx->tmQh, ctx->tmXh, prm);
            }
            else if (KS == 1)
                kstar_tc_kernel<1, 1><<<grid, KT_THREADS, smem, st>>>(ctx->tmQh, ctx->tmXh, prm);
            else if (KS == 2)
                kstar_tc_kernel<1, 2><<<grid, KT_THREADS, smem, st>>>(ctx->tmQh, ctx->tmXh, prm);
            else
                kstar_tc_kernel<1, 3><<<grid, KT_THREADS, smem, st>>>(ctx->tmQh, ctx->tmXh, prm);
            LAUNCH_CHECK();
            return SLSGP_OK;
        }
        if (tiled)
        {
            const size_t smem = sizeof(float) * (size_t) (64 * ((D + 3) & ~3) + D * 128);
            const dim3   grid(ldt / 128, (unsigned) (Mpad / 64));
            if (ctx->kernel_type == 0)
                kstar16_kernel<0><<<grid, 256, smem, st>>>(d_Xq, Mc, D, ctx->N, ldt, ptr<float>(ctx->Xs32), dp(ctx->inv_l), ptr<TcScales>(ctx->tcs), Ks, Ks_lo, Gs, Gs_lo);
            else
                kstar16_kernel<1><<<grid, 256, smem, st>>>(d_Xq, Mc, D, ctx->N, ldt, ptr<float>(ctx->Xs32), dp(ctx->inv_l), ptr<TcScales>(ctx->tcs), Ks, Ks_lo, Gs, Gs_lo);
        }
        else
        {
            const int    n_sm     = n_sm_k;
            const size_t smem     = sizeof(float) * (size_t) (64 * ((D + 3) & ~3) + 2 * D * 64);
            const int    n_strips = (int) (Mpad / 64);
            const int    grid     = std::min(n_strips, (under_gemm ? 1 : 4) * n_sm);
            if (ctx->kernel_type == 0)
                kstar16_strip_kernel<0><<<grid, 256, smem, st>>>(d_Xq, Mc, D, ctx->N, ldt, ptr<float>(ctx->Xs32), dp(ctx->inv_l), ptr<TcScales>(ctx->tcs), Ks,
                                                                  Ks_lo, Gs, Gs_lo, n_strips);
            else
                kstar16_strip_kernel<1><<<grid, 256, smem, st>>>(d_Xq, Mc, D, ctx->N, ldt, ptr<float>(ctx->Xs32), dp(ctx->inv_l), ptr<TcScales>(ctx->tcs), Ks,
                                                                  Ks_lo, Gs, Gs_lo, n_strips);
        }
        LAUNCH_CHECK();
        return SLSGP_OK;
    }

    // contraction + fused epilogue + acquisition formulas of one shard whose k* operand sits in shard buffer `buf`
    slsgp_status tensor_main(slsgp_ctx* ctx, int buf, int acq_type, double ucb_beta, const double* d_Xq, long long Mc, SweepOut out,
                             RefineList refine = RefineList{nullptr, nullptr, 0, 0, 0.0, 0})
    {
        const int       D = ctx->D, ldt = ctx->ldt, passes = tensor_passes(ctx->sweep_mode);
        const long long Mpad = round_up64(Mc, TC_BM * ctx->tc_ncta);
        const long long row0 = (long long) buf * 4 * ctx->tc_Mcap;
        __half*         Ks    = ptr<__half>(ctx->Ks) + (size_t) row0 * ldt;
        __half*         Ks_lo = passes > 1 ? Ks + (size_t) ctx->tc_Mcap * ldt : nullptr;
        int n_sm = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
        int split = 1;
        {
            ProfScope    ps(ctx, "tc_gemm");
            TcGemmParams prm;
            prm.ldt = ldt, prm.kb = round_up(ctx->N, TC_BK) / TC_BK, prm.ncb = ldt / TC_BN, prm.D = D, prm.stages = 0;
            prm.stage_bytes = 0;
            prm.n_cand_blocks = (int) (Mpad / TC_BM), prm.Mc = Mc, prm.a_row0 = (int) row0;
            prm.passes = passes, prm.a_lo_row = (int) ctx->tc_Mcap, prm.b_lo_row = ldt + TC_BN, prm.Ks_lo = Ks_lo;
            prm.matern = ctx->kernel_type == SLSGP_KERNEL_ARD_MATERN52, prm.g_row = (int) (2 * ctx->tc_Mcap);
            prm.Gs = Ks + (size_t) 2 * ctx->tc_Mcap * ldt, prm.Gs_lo = passes > 1 ? Ks + (size_t) 3 * ctx->tc_Mcap * ldt : nullptr;
            prm.Ks = Ks, prm.Xt = ptr<float>(ctx->Xt), prm.sc = ptr<TcScales>(ctx->tcs);
            prm.se_factor = (ctx->compat & SLSGP_COMPAT_SE_XGRAD_2X) ? 2.0 : 1.0;
            prm.stats = ptr<double4>(ctx->stats), prm.P1 = dp(ctx->P1), prm.P2 = dp(ctx->P2);
            prm.ldp = D; // packed: one candidate's D sums are contiguous (the FP64 path's GEMM output needs the 64-row padding)
            prm.err = ptr<int>(ctx->tc_err);
            static const int split_env = std::getenv("SLSGP_TC_SPLIT") ? std::atoi(std::getenv("SLSGP_TC_SPLIT")) : 2;
            prm.split = std::max(1, std::min(std::min(split_env, 4), prm.ncb));
            prm.qx = ptr<double2>(ctx->tc_qx), prm.P2x = dp(ctx->tc_P2x), prm.part_stride = ctx->Mcap;
            split = prm.split;
            switch (ctx->XP)
            {
                case 8: TRY(launch_tc_gemm<8>(ctx, prm, n_sm)); break;
                case 12: TRY(launch_tc_gemm<12>(ctx, prm, n_sm)); break;
                case 20: TRY(launch_tc_gemm<20>(ctx, prm, n_sm)); break;
                case 36: TRY(launch_tc_gemm<36>(ctx, prm, n_sm)); break;
                case 68: TRY(launch_tc_gemm<68>(ctx, prm, n_sm)); break;
                default: return fail(ctx, SLSGP_ERR_INVALID, "tensor sweep: unsupported D");
            }
        }
        return sweep_finish(ctx, acq_type, ucb_beta, d_Xq, Mc, out, split - 1, 0.5, D, refine);
    }

    slsgp_status sweep_shard(slsgp_ctx* ctx, int acq_type, double ucb_beta, const double* d_Xq, long long Mc,
                             SweepOut out)
    {
        const int       ld = ctx->ld, D = ctx->D, Dp = ctx->Dp;
        const long long Mp = round_up64(Mc, TILE);
        const bool      want_grad = out.dmu || out.dsigma || out.grad;
        const double    se_factor = (ctx->compat & SLSGP_COMPAT_SE_XGRAD_2X) ? 2.0 : 1.0;

        {
            ProfScope ps(ctx, "sweep_kstar");
            kstar_tile_kernel<<<dim3(ld / TILE, (unsigned) (Mp / TILE)), 256, 0, ctx->stream>>>(
                dp(ctx->X), ctx->N, D, ld, d_Xq, Mc, dp(ctx->theta), dp(ctx->inv_l), ctx->kernel_type, se_factor,
                dp(ctx->Kstar), dp(ctx->Gstar));
            LAUNCH_CHECK();
        }
        {
            // Beta = Kinv * Kstar
            ProfScope ps(ctx, "sweep_gemm");
            GemmArgs  g = gemm_args(dp(ctx->Kinv), dp(ctx->Kstar), dp(ctx->Beta), ld, (int) Mp, ld, ld, ld, ld, 1.0, 0.0);
            TRY((launch_gemm<false, false>(ctx, g)));
        }
        {
            ProfScope ps(ctx, "sweep_reduce");
            column_reduce_kernel<<<(unsigned) ((Mc * 32 + 255) / 256), 256, 0, ctx->stream>>>(
                dp(ctx->Kstar), dp(ctx->Gstar), dp(ctx->Beta), dp(ctx->alpha), ld, Mc, ptr<double4>(ctx->stats));
            LAUNCH_CHECK();
        }
        if (want_grad)
        {
            ProfScope ps(ctx, "sweep_grad_gemm");
            GemmArgs  p1 = gemm_args(dp(ctx->Xpad), dp(ctx->Gstar), dp(ctx->P1), Dp, (int) Mp, ld, Dp, ld, Dp, 1.0, 0.0);
            TRY((launch_gemm<false, false>(ctx, p1)));
            GemmArgs p2 = gemm_args(dp(ctx->Xpad), dp(ctx->Beta), dp(ctx->P2), Dp, (int) Mp, ld, Dp, ld, Dp, 1.0, 0.0);
            TRY((launch_gemm<false, false>(ctx, p2)));
        }
        return sweep_finish(ctx, acq_type, ucb_beta, d_Xq, Mc, out);
    }

    slsgp_status require_model(slsgp_ctx* ctx)
    {
        if (!ctx->has_alpha)
            return fail(ctx, SLSGP_ERR_STATE,
                        "sweep needs a fitted model: slsgp_set_data, slsgp_gram, slsgp_factor, slsgp_solve_alpha");
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // The shard pipeline shared by every K4 entry point. Shard s goes through
    //   in   (pre stream)   candidates into shard buffer s % 2 (H2D copy / counter-based generator / caller's device
    //                       pointer) and, in tensor mode, their k* operand
    //   main (ctx->stream)  contraction, epilogue, acquisition formulas (+ running arg-max)
    //   out  (post stream)  D2H of the results (host-buffer entry points only)
    // so the copies and the k* generator of neighbouring shards hide behind the contraction. Everything is ordered
    // behind ctx->stream: work enqueued there before the call is complete before `in` starts, and ctx->stream
    // waits for the last `out` before the call returns.
    // ---------------------------------------------------------------------------------------------------------
    struct SweepJob
    {
        int           acq_type = 0;
        double        ucb_beta = 0.0;
        long long     M        = 0;
        const double* d_Xq     = nullptr; // candidates already on the device (D x M), or
        const double* h_Xq     = nullptr; // on the host, or
        bool          generate = false;   // counter-based: candidate i = candidate_coord(seed, first + i, .)
        uint64_t      seed     = 0;
        long long     first    = 0;
        bool          host_out = false;   // outputs below are host pointers (else device pointers); any may be null
        double *      mu = nullptr, *sigma = nullptr, *dmu = nullptr, *dsigma = nullptr, *val = nullptr, *grad = nullptr;
        bool          argmax = false;     // fold every shard's values into ctx->am_acc (index = first + i)
        long long     slice_len = 0;      // > 0: also keep the best candidate of every `slice_len` consecutive indices
        ArgMax*       slice_best = nullptr;
    };

    bool is_pageable(const void* p)
    {
        if (!p) return false;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
        {
            cudaGetLastError();
            return true;
        }
        return at.type == cudaMemoryTypeUnregistered;
    }
    slsgp_status ensure_pinned(slsgp_ctx* ctx, double** buf, size_t* have, size_t bytes)
    {
        if (*have >= bytes) return SLSGP_OK;
        if (*buf) cudaFreeHost(*buf);
        *buf = nullptr, *have = 0;
        if (cudaMallocHost(buf, bytes) != cudaSuccess)
        {
            cudaGetLastError();
            return fail(ctx, SLSGP_ERR_NOMEM, "cudaMallocHost of " + std::to_string(bytes) + " bytes failed");
        }
        *have = bytes;
        return SLSGP_OK;
    }

    // Second tier of the tensor sweep: the candidates sweep_finish_kernel listed (sigma^2 < tau * a) go through the IEEE-double
    // sweep in chunks and their results replace the tensor-path ones. One host synchronisation (the count); nothing to do in the
    // common case of candidates away from the data.
    slsgp_status refine_listed(slsgp_ctx* ctx, const SweepJob& job, long long cap)
    {
        cudaStream_t main = ctx->stream;
        const int    D = ctx->D;
        int          count = 0;
        CUDA_TRY(cudaMemcpyAsync(&count, ctx->rf_count.p, sizeof(int), cudaMemcpyDeviceToHost, main));
        CUDA_TRY(cudaStreamSynchronize(main)); // main has already waited for the last result copies of a host-buffer job
        // Dense data (many observations in few dimensions): nearly every candidate sits close to a data point and is re-evaluated, so
        // the tensor pass is wasted work. Remember it for this model: run_sweep then sweeps in IEEE double from the start.
        if (job.M >= 8192 && 2LL * count > job.M) ctx->dense_version = ctx->model_version;
        const int R = (int) std::min<long long>(count, cap);
        if (R <= 0) return SLSGP_OK;
        const int Rc = (int) std::min<long long>(2048, ctx->Mcap); // P1 / P2 / stats hold Mcap candidates
        const size_t col = sizeof(double) * (size_t) ctx->ld;
        TRY(ensure(ctx, ctx->Kstar, col * Rc));
        TRY(ensure(ctx, ctx->Gstar, col * Rc));
        TRY(ensure(ctx, ctx->Beta, col * Rc));
        TRY(ensure(ctx, ctx->rf_X, sizeof(double) * (size_t) D * Rc));
        TRY(ensure(ctx, ctx->rf_out, sizeof(double) * (size_t) (3 + 3 * D) * Rc));
        std::vector<long long> index;
        std::vector<double>    hx, hout;
        if (job.h_Xq || job.host_out)
        {
            index.resize((size_t) R);
            CUDA_TRY(cudaMemcpyAsync(index.data(), ctx->rf_index.p, sizeof(long long) * (size_t) R, cudaMemcpyDeviceToHost, main));
            CUDA_TRY(cudaStreamSynchronize(main));
        }
        double*  base = dp(ctx->rf_out);
        // the chunk results travel to the host as one block of Rc-sized arrays: entries past the chunk's count (and arrays the job
        // does not ask for) are never written by the sweep, so give them a defined value
        CUDA_TRY(cudaMemsetAsync(base, 0, sizeof(double) * (size_t) (3 + 3 * D) * Rc, main));
        SweepOut ro;
        ro.mu = base, ro.sigma = base + Rc, ro.val = base + 2 * (size_t) Rc;
        ro.dmu = base + 3 * (size_t) Rc, ro.dsigma = ro.dmu + (size_t) D * Rc, ro.grad = ro.dsigma + (size_t) D * Rc;
        const bool want_grad = job.dmu || job.dsigma || job.grad;
        if (!want_grad) ro.dmu = ro.dsigma = ro.grad = nullptr;
        for (int r0 = 0; r0 < R; r0 += Rc)
        {
            const int        Rn  = std::min(Rc, R - r0);
            const long long* idx = ptr<long long>(ctx->rf_index) + r0;
            if (job.h_Xq)
            {
                hx.resize((size_t) Rn * D);
                for (int r = 0; r < Rn; ++r) std::memcpy(&hx[(size_t) r * D], job.h_Xq + (size_t) index[(size_t) r0 + r] * D, sizeof(double) * D);
                CUDA_TRY(cudaMemcpyAsync(ctx->rf_X.p, hx.data(), sizeof(double) * hx.size(), cudaMemcpyHostToDevice, main));
                CUDA_TRY(cudaStreamSynchronize(main)); // hx is pageable and reused by the next chunk
            }
            else
            {
                refine_gather_kernel<<<(Rn * D + 255) / 256, 256, 0, main>>>(idx, Rn, D, job.d_Xq, job.generate ? 1 : 0, job.seed, job.first, dp(ctx->rf_X));
                LAUNCH_CHECK();
            }
            TRY(sweep_shard(ctx, job.acq_type, job.ucb_beta, dp(ctx->rf_X), Rn, ro));
            if (job.argmax)
            {
                argmax_indexed_kernel<<<1, 256, 0, main>>>(ro.val, idx, Rn, job.first, ptr<ArgMax>(ctx->am_acc));
                LAUNCH_CHECK();
            }
            if (job.host_out)
            {
                hout.resize((size_t) (3 + 3 * D) * Rc);
                CUDA_TRY(cudaMemcpyAsync(hout.data(), base, sizeof(double) * hout.size(), cudaMemcpyDeviceToHost, main));
                CUDA_TRY(cudaStreamSynchronize(main));
                const double *h_mu = hout.data(), *h_sigma = h_mu + Rc, *h_val = h_mu + 2 * (size_t) Rc, *h_dmu = h_mu + 3 * (size_t) Rc,
                             *h_dsigma = h_dmu + (size_t) D * Rc, *h_grad = h_dsigma + (size_t) D * Rc;
                for (int r = 0; r < Rn; ++r)
                {
                    const size_t m = (size_t) index[(size_t) r0 + r];
                    if (job.mu) job.mu[m] = h_mu[r];
                    if (job.sigma) job.sigma[m] = h_sigma[r];
                    if (job.val) job.val[m] = h_val[r];
                    if (job.dmu) std::memcpy(job.dmu + m * D, h_dmu + (size_t) r * D, sizeof(double) * D);
                    if (job.dsigma) std::memcpy(job.dsigma + m * D, h_dsigma + (size_t) r * D, sizeof(double) * D);
                    if (job.grad) std::memcpy(job.grad + m * D, h_grad + (size_t) r * D, sizeof(double) * D);
                }
            }
            else
            {
                SweepOut dst;
                dst.mu = job.mu, dst.sigma = job.sigma, dst.val = job.val, dst.dmu = job.dmu, dst.dsigma = job.dsigma, dst.grad = job.grad;
                refine_scatter_kernel<<<(Rn * (D + 1) + 255) / 256, 256, 0, main>>>(idx, Rn, D, ro, dst);
                LAUNCH_CHECK();
            }
        }
        return SLSGP_OK;
    }

    slsgp_status run_sweep(slsgp_ctx* ctx, const SweepJob& job)
    {
        if (is_tensor_mode(ctx->sweep_mode) && ctx->refine_tau > 0.0 && job.slice_len == 0 && ctx->dense_version == ctx->model_version)
        {
            // the last sweep over this model sent most of its candidates to the second tier (refine_listed): skip the tensor pass
            const int user_mode = ctx->sweep_mode;
            ctx->sweep_mode     = SLSGP_SWEEP_FP64, ctx->Mcap = 0;
            slsgp_status st     = ensure_sweep_workspace(ctx, job.M);
            if (st == SLSGP_OK) st = run_sweep(ctx, job);
            ctx->sweep_mode = user_mode, ctx->Mcap = 0; // the next call re-establishes the tensor workspace (buffers only ever grow)
            return st;
        }
        const bool      tensor = is_tensor_mode(ctx->sweep_mode);
        const int       D = ctx->D;
        const long long cap = ctx->Mcap, n_shards = (job.M + cap - 1) / cap;
        if (n_shards == 0) return SLSGP_OK;
        static const bool overlap = !(std::getenv("SLSGP_PIPELINE") && std::atoi(std::getenv("SLSGP_PIPELINE")) == 0);
        // k* of shard s + 1 on `pre`, under the contraction of shard s (see kstar16_strip_kernel)
        // (only when the contraction is long enough to hide it: below ~1000 observations a shard is launch-latency bound and the
        // one-CTA-per-SM generator would be the slower of the two; measured on the D = 64 optimiser loop, N <= 120: 74 vs 88 ms)
        // Opt-in (SLSGP_KSTAR_OVERLAP=1). The sweep is power-bound at N = 2048 (sw_power_cap, ~1.2-1.4 GHz): running the generator
        // under the contraction moves its energy, not its time, and the whole step gains 2 % (39.7 vs 38.9 M candidates/s) while
        // the contraction launch itself stretches from 0.75 to 0.90 ms; the default keeps the two kernels back to back.
        static const bool kstar_overlap_env = std::getenv("SLSGP_KSTAR_OVERLAP") && std::atoi(std::getenv("SLSGP_KSTAR_OVERLAP")) != 0;
        const bool        kstar_overlap     = kstar_overlap_env && tensor && ctx->ldt >= 1024;
        // a single shard has nothing to overlap with: keep it on one stream (one-candidate calls are latency-bound)
        const bool   multi = overlap && n_shards > 1;
        cudaStream_t main = ctx->stream, pre = multi ? ctx->pre_stream : main, post = multi ? ctx->post_stream : main;
        if (multi)
        {
            CUDA_TRY(cudaEventRecord(ctx->ev_start, main));
            CUDA_TRY(cudaStreamWaitEvent(pre, ctx->ev_start, 0));
            CUDA_TRY(cudaStreamWaitEvent(post, ctx->ev_start, 0));
        }

        // Host buffers in pageable memory (what a C++ caller's Eigen matrices are): cudaMemcpyAsync would stage them through the
        // driver synchronously, shard after shard, with no overlap. Instead they go through a pinned ring of two shard-sized slots
        // each way: this thread copies (memcpy) while the copy engines and the SMs work on the neighbouring shards.
        const bool staged = (job.h_Xq || job.host_out) &&
                            (is_pageable(job.h_Xq) || is_pageable(job.mu) || is_pageable(job.sigma) || is_pageable(job.val) || is_pageable(job.dmu) ||
                             is_pageable(job.dsigma) || is_pageable(job.grad));
        size_t out_off[7] = {0, 0, 0, 0, 0, 0, 0}; // slot layout in doubles: mu | sigma | val | dmu | dsigma | grad (requested ones only)
        if (staged)
        {
            double* const outs[6] = {job.mu, job.sigma, job.val, job.dmu, job.dsigma, job.grad};
            for (int i = 0; i < 6; ++i) out_off[i + 1] = out_off[i] + (outs[i] ? (size_t) cap * (i < 3 ? 1 : D) : 0);
            if (job.h_Xq) TRY(ensure_pinned(ctx, &ctx->pin_in, &ctx->pin_in_bytes, 2 * sizeof(double) * (size_t) D * cap));
            if (job.host_out && out_off[6]) TRY(ensure_pinned(ctx, &ctx->pin_out, &ctx->pin_out_bytes, 2 * sizeof(double) * out_off[6]));
        }
        // copy the results of shard s from its pinned slot to the caller's arrays once their transfer has completed
        auto drain = [&](long long s) -> slsgp_status {
            const int       b  = (int) (s & 1);
            const long long m0 = s * cap, Mc = std::min<long long>(cap, job.M - m0);
            CUDA_TRY(cudaEventSynchronize(ctx->ev_out[b]));
            const double* slot = ctx->pin_out + (size_t) b * out_off[6];
            double* const outs[6] = {job.mu, job.sigma, job.val, job.dmu, job.dsigma, job.grad};
            for (int i = 0; i < 6; ++i)
                if (outs[i])
                {
                    const size_t w = i < 3 ? 1 : (size_t) D;
                    std::memcpy(outs[i] + (size_t) m0 * w, slot + out_off[i], sizeof(double) * (size_t) Mc * w);
                }
            return SLSGP_OK;
        };

        auto xq_of = [&](long long s) -> const double* {
            return job.d_Xq ? job.d_Xq + (size_t) (s * cap) * D : dp(ctx->Xq) + (size_t) (s & 1) * D * cap;
        };
        auto stage_in = [&](long long s) -> slsgp_status {
            const int       b  = (int) (s & 1);
            const long long m0 = s * cap, Mc = std::min<long long>(cap, job.M - m0);
            if (s >= 2) CUDA_TRY(cudaStreamWaitEvent(pre, ctx->ev_main[b], 0)); // shard s - 2 no longer reads buffer b
            double* xq = dp(ctx->Xq) + (size_t) b * D * cap;
            if (job.h_Xq && staged)
            {
                // pageable caller memory: this thread copies the shard into the pinned slot (free once the transfer of shard
                // s - 2 out of it has completed), the DMA engine takes it from there while earlier shards compute
                if (s >= 2) CUDA_TRY(cudaEventSynchronize(ctx->ev_in[b]));
                double* slot = ctx->pin_in + (size_t) b * D * cap;
                std::memcpy(slot, job.h_Xq + (size_t) m0 * D, sizeof(double) * (size_t) Mc * D);
                CUDA_TRY(cudaMemcpyAsync(xq, slot, sizeof(double) * (size_t) Mc * D, cudaMemcpyHostToDevice, pre));
            }
            else if (job.h_Xq)
                CUDA_TRY(cudaMemcpyAsync(xq, job.h_Xq + (size_t) m0 * D, sizeof(double) * (size_t) Mc * D, cudaMemcpyHostToDevice, pre));
            else if (job.generate)
            {
                candidates_kernel<<<(unsigned) ((Mc * D + 255) / 256), 256, 0, pre>>>(job.seed, job.first + m0, Mc, D, xq);
                LAUNCH_CHECK();
            }
            // the k* generator rides on `pre` too: it co-resides with the persistent contraction CTAs (64 registers per thread)
            if (tensor && kstar_overlap) TRY(tensor_kstar(ctx, b, xq_of(s), Mc, pre, multi && s > 0));
            if (multi || staged) CUDA_TRY(cudaEventRecord(ctx->ev_in[b], pre));
            return SLSGP_OK;
        };

        // Two-tier precision (tensor modes, jobs that return per-candidate arrays): sweep_finish_kernel lists the candidates with
        // sigma^2 < tau * a; they are re-evaluated in IEEE double after the last shard (refine_listed below).
        // (arg-max jobs defer the listed candidates: they cannot win until their IEEE-double value is folded in; the slice
        // winners of slsgp_acq_maximize are only starting points and are left alone)
        const bool      refine_on  = tensor && ctx->refine_tau > 0.0 && job.slice_len == 0 &&
                                (job.argmax || job.mu || job.sigma || job.val || job.dmu || job.dsigma || job.grad);
        const long long refine_cap = std::min<long long>(job.M, 1LL << 21);
        if (refine_on)
        {
            TRY(ensure(ctx, ctx->rf_count, sizeof(int)));
            TRY(ensure(ctx, ctx->rf_index, sizeof(long long) * (size_t) refine_cap));
            CUDA_TRY(cudaMemsetAsync(ctx->rf_count.p, 0, sizeof(int), main));
        }

        TRY(stage_in(0));
        for (long long s = 0; s < n_shards; ++s)
        {
            const int       b  = (int) (s & 1);
            const long long m0 = s * cap, Mc = std::min<long long>(cap, job.M - m0);
            if (s + 1 < n_shards) TRY(stage_in(s + 1));

            // ---- main
            if (multi) CUDA_TRY(cudaStreamWaitEvent(main, ctx->ev_in[b], 0));
            if (job.host_out && s >= 2) CUDA_TRY(cudaStreamWaitEvent(main, ctx->ev_out[b], 0)); // result buffer b drained
            SweepOut o;
            if (job.host_out)
            {
                const size_t ov = (size_t) b * cap, og = ov * D;
                o.mu = job.mu ? dp(ctx->o_mu) + ov : nullptr, o.sigma = job.sigma ? dp(ctx->o_sigma) + ov : nullptr;
                o.val = job.val ? dp(ctx->o_val) + ov : nullptr, o.dmu = job.dmu ? dp(ctx->o_dmu) + og : nullptr;
                o.dsigma = job.dsigma ? dp(ctx->o_dsigma) + og : nullptr, o.grad = job.grad ? dp(ctx->o_grad) + og : nullptr;
            }
            else
            {
                o.mu = job.mu ? job.mu + m0 : nullptr, o.sigma = job.sigma ? job.sigma + m0 : nullptr;
                o.val = job.val ? job.val + m0 : nullptr, o.dmu = job.dmu ? job.dmu + (size_t) m0 * D : nullptr;
                o.dsigma = job.dsigma ? job.dsigma + (size_t) m0 * D : nullptr, o.grad = job.grad ? job.grad + (size_t) m0 * D : nullptr;
            }
            if ((job.argmax || job.slice_len > 0) && !o.val) o.val = dp(ctx->o_val) + (size_t) b * cap;
            if (tensor)
            {
                if (!kstar_overlap) TRY(tensor_kstar(ctx, b, xq_of(s), Mc, main, false));
                RefineList rl{nullptr, nullptr, 0, 0, 0.0, 0};
                if (refine_on) rl = RefineList{ptr<int>(ctx->rf_count), ptr<long long>(ctx->rf_index), refine_cap, m0, ctx->refine_tau, job.argmax ? 1 : 0};
                TRY(tensor_main(ctx, b, job.acq_type, job.ucb_beta, xq_of(s), Mc, o, rl));
            }
            else
                TRY(sweep_shard(ctx, job.acq_type, job.ucb_beta, xq_of(s), Mc, o));
            if (job.argmax)
            {
                const int nblk = (int) std::min<long long>(1024, (Mc + 255) / 256);
                argmax_partial_kernel<<<nblk, 256, 0, main>>>(o.val, Mc, job.first + m0, ptr<ArgMax>(ctx->am_part));
                LAUNCH_CHECK();
                argmax_final_kernel<<<1, 256, 0, main>>>(ptr<ArgMax>(ctx->am_part), nblk, ptr<ArgMax>(ctx->am_acc));
                LAUNCH_CHECK();
            }
            if (job.slice_len > 0)
            {
                const long long s_lo = m0 / job.slice_len, s_hi = (m0 + Mc - 1) / job.slice_len;
                slice_argmax_kernel<<<(unsigned) (s_hi - s_lo + 1), 256, 0, main>>>(o.val, Mc, job.first + m0, job.first, job.slice_len,
                                                                                   s_lo, job.slice_best);
                LAUNCH_CHECK();
            }
            if (multi) CUDA_TRY(cudaEventRecord(ctx->ev_main[b], main));

            // ---- out
            if (job.host_out)
            {
                if (multi) CUDA_TRY(cudaStreamWaitEvent(post, ctx->ev_main[b], 0));
                const size_t sv = sizeof(double) * (size_t) Mc, sg = sv * D;
                if (staged)
                {
                    if (s >= 2) TRY(drain(s - 2)); // frees pinned slot b (the host is at most two shards behind the device)
                    double*       slot = ctx->pin_out + (size_t) b * out_off[6];
                    const double* src[6] = {o.mu, o.sigma, o.val, o.dmu, o.dsigma, o.grad};
                    double* const outs[6] = {job.mu, job.sigma, job.val, job.dmu, job.dsigma, job.grad};
                    for (int i = 0; i < 6; ++i)
                        if (outs[i]) CUDA_TRY(cudaMemcpyAsync(slot + out_off[i], src[i], i < 3 ? sv : sg, cudaMemcpyDeviceToHost, post));
                    CUDA_TRY(cudaEventRecord(ctx->ev_out[b], post));
                    continue;
                }
                if (job.mu) CUDA_TRY(cudaMemcpyAsync(job.mu + m0, o.mu, sv, cudaMemcpyDeviceToHost, post));
                if (job.sigma) CUDA_TRY(cudaMemcpyAsync(job.sigma + m0, o.sigma, sv, cudaMemcpyDeviceToHost, post));
                if (job.val) CUDA_TRY(cudaMemcpyAsync(job.val + m0, o.val, sv, cudaMemcpyDeviceToHost, post));
                if (job.dmu) CUDA_TRY(cudaMemcpyAsync(job.dmu + (size_t) m0 * D, o.dmu, sg, cudaMemcpyDeviceToHost, post));
                if (job.dsigma) CUDA_TRY(cudaMemcpyAsync(job.dsigma + (size_t) m0 * D, o.dsigma, sg, cudaMemcpyDeviceToHost, post));
                if (job.grad) CUDA_TRY(cudaMemcpyAsync(job.grad + (size_t) m0 * D, o.grad, sg, cudaMemcpyDeviceToHost, post));
                if (multi) CUDA_TRY(cudaEventRecord(ctx->ev_out[b], post));
            }
        }
        if (job.host_out && staged)
            for (long long s = std::max<long long>(0, n_shards - 2); s < n_shards; ++s) TRY(drain(s));
        if (job.host_out && (multi || staged))
            for (long long s = std::max<long long>(0, n_shards - 2); s < n_shards; ++s)
                CUDA_TRY(cudaStreamWaitEvent(main, ctx->ev_out[s & 1], 0));
        if (refine_on) TRY(refine_listed(ctx, job, refine_cap));
        return SLSGP_OK;
    }
    // ---------------------------------------------------------------------------------------------------------
    // Multi-GPU group: the fitted model of the primary context is replicated onto the peers by peer-to-peer copies
    // (NVLink / NVSwitch: X, K^-1, alpha, f_best, hyper-parameters: N^2 doubles dominate, 33.5 MB at N = 2048), once per
    // model version; candidate ranges are then split over the group, one host thread per device, and the per-device winners
    // (value, index, point: 8 (2 + D) bytes each) are reduced on the host: highest value, lowest candidate index on ties, so
    // the result does not depend on the number of devices. Gram, Cholesky, inverse and the MAP objectives stay on the primary
    // (replicas only).
    // ---------------------------------------------------------------------------------------------------------
    void adopt_sweep_mode(slsgp_ctx* c, int mode)
    {
        if (is_tensor_mode(mode) != is_tensor_mode(c->sweep_mode)) c->Mcap = 0; // the per-shard scratch differs between the two families
        c->sweep_mode = mode;
    }

    slsgp_status replicate_model(slsgp_ctx* ctx /* primary */, slsgp_ctx* dst)
    {
        if (dst->replica_of == ctx->model_version && dst->has_alpha) return SLSGP_OK;
        if (!ctx->has_alpha) return fail(ctx, SLSGP_ERR_STATE, "multi-GPU sweep needs a fitted model on the primary context");
        CUDA_TRY(cudaSetDevice(dst->device));
        if (!dst->peer_access_tried)
        {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dst->device, ctx->device) == cudaSuccess && can)
                if (cudaDeviceEnablePeerAccess(ctx->device, 0) != cudaSuccess) cudaGetLastError(); // already enabled is fine
            dst->peer_access_tried = true;
        }
        const bool same_shape = dst->has_data && dst->N == ctx->N && dst->D == ctx->D;
        if (!same_shape)
        {
            dst->N = ctx->N, dst->D = ctx->D, dst->ld = ctx->ld, dst->Dp = ctx->Dp, dst->ldx = ctx->ldx;
            dst->Mcap = 0, dst->tc_Mcap = 0;
            dst->P = 0, dst->pref_total = 0;
            const slsgp_status st = ensure_model_buffers(dst);
            if (st != SLSGP_OK) return fail(ctx, st, "replica on device " + std::to_string(dst->device) + ": " + dst->err);
        }
        dst->tc_ready = false;
        dst->kernel_type = ctx->kernel_type, dst->noise = ctx->noise, dst->theta_host = ctx->theta_host, dst->logdet_host = ctx->logdet_host;
        dst->compat = ctx->compat, dst->refine_tau = ctx->refine_tau;
        adopt_sweep_mode(dst, ctx->sweep_mode);
        const size_t ld = ctx->ld, D = ctx->D, mat = sizeof(double) * ld * ld;
        struct Piece
        {
            DevBuf *to, *from;
            size_t  bytes;
        } pieces[] = {{&dst->X, &ctx->X, sizeof(double) * ld * D},      {&dst->Xpad, &ctx->Xpad, sizeof(double) * (size_t) ctx->Dp * ld},
                      {&dst->XT1, &ctx->XT1, sizeof(double) * ld * ctx->ldx}, {&dst->theta, &ctx->theta, sizeof(double) * (D + 1)},
                      {&dst->inv_l, &ctx->inv_l, sizeof(double) * D},   {&dst->Kinv, &ctx->Kinv, mat},
                      {&dst->alpha, &ctx->alpha, sizeof(double) * ld},   {&dst->y, &ctx->y, sizeof(double) * ld},
                      {&dst->fbest, &ctx->fbest, sizeof(double)},        {&dst->fbest_idx, &ctx->fbest_idx, sizeof(int)}};
        for (const Piece& pc : pieces)
            CUDA_TRY(cudaMemcpyPeerAsync(pc.to->p, dst->device, pc.from->p, ctx->device, pc.bytes, dst->stream));
        CUDA_TRY(cudaStreamSynchronize(dst->stream));
        // a replica serves sweeps only: no factor to update, no Gram matrix to read back
        dst->dense_version = ~0ull; // a new model: the second-tier statistics of the old one do not carry over
        dst->has_data = true, dst->has_gram = false, dst->has_factor = false, dst->has_W = false, dst->has_inverse = true, dst->has_alpha = true;
        dst->replica_of = ctx->model_version;
        return SLSGP_OK;
    }

    // every member of the group with its share [first_g, first_g + count_g) of [first, first + count)
    struct GroupShare
    {
        slsgp_ctx* ctx;
        long long  first, count;
    };
    std::vector<GroupShare> split_over_group(slsgp_ctx* ctx, long long first, long long count)
    {
        const long long n = 1 + (long long) ctx->peers.size(), base = count / n, rem = count % n;
        std::vector<GroupShare> out;
        long long               at = first;
        for (long long g = 0; g < n; ++g)
        {
            const long long c = base + (g < rem ? 1 : 0);
            if (c > 0) out.push_back({g == 0 ? ctx : ctx->peers[(size_t) g - 1], at, c});
            at += c;
        }
        return out;
    }
    // ranges smaller than this stay on the primary: a replica costs a peer copy of K^-1 and a thread hand-off
    long long group_min_count() { return 1LL << 16; }

    template <typename F> slsgp_status run_on_group(slsgp_ctx* ctx, const std::vector<GroupShare>& shares, F&& body)
    {
        for (const GroupShare& sh : shares)
            if (sh.ctx != ctx)
            {
                adopt_sweep_mode(sh.ctx, ctx->sweep_mode);
                sh.ctx->refine_tau = ctx->refine_tau, sh.ctx->compat = ctx->compat;
                TRY(replicate_model(ctx, sh.ctx));
            }
        std::vector<slsgp_status> st(shares.size(), SLSGP_OK);
        std::vector<std::thread>  workers;
        for (size_t g = 1; g < shares.size(); ++g) workers.emplace_back([&, g]() { st[g] = body(shares[g], g); });
        st[0] = body(shares[0], 0);
        for (auto& w : workers) w.join();
        CUDA_TRY(cudaSetDevice(ctx->device));
        for (size_t g = 0; g < shares.size(); ++g)
            if (st[g] != SLSGP_OK) return fail(ctx, st[g], "device " + std::to_string(shares[g].ctx->device) + ": " + shares[g].ctx->err);
        return SLSGP_OK;
    }
} // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C"
{
    const char* slsgp_status_string(slsgp_status s)
    {
        switch (s)
        {
            case SLSGP_OK: return "ok";
            case SLSGP_ERR_INVALID: return "invalid argument";
            case SLSGP_ERR_STATE: return "call-order error";
            case SLSGP_ERR_NOT_SPD: return "matrix is not symmetric positive definite";
            case SLSGP_ERR_NAN: return "non-finite input";
            case SLSGP_ERR_CUDA: return "CUDA error";
            case SLSGP_ERR_NOMEM: return "out of device memory";
        }
        return "unknown";
    }

    slsgp_status slsgp_ctx_create(int device, slsgp_ctx** ctx_out)
    {
        if (!ctx_out) return SLSGP_ERR_INVALID;
        *ctx_out = nullptr;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
        {
            cudaGetLastError();
            return SLSGP_ERR_CUDA; // no CPU fallback by design
        }
        slsgp_ctx* ctx = new slsgp_ctx;
        ctx->device    = device;
        // every failure below goes through slsgp_ctx_destroy, which releases whatever had been created by then
        bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
        ctx->stream = ctx->own_stream;
        ok = ok && cudaStreamCreateWithFlags(&ctx->pre_stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&ctx->post_stream, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming) == cudaSuccess;
        for (int b = 0; b < 2 && ok; ++b)
            ok = cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ctx->ev_main[b], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming) == cudaSuccess;
        if (const char* e = std::getenv("SLSGP_TC_PAIR")) ctx->tc_ncta = std::atoi(e) ? 2 : 1;
        if (const char* e = std::getenv("SLSGP_CHOL_TWO_LEVEL_FROM")) ctx->chol_two_level_from = std::max(2, std::atoi(e));
        if (const char* e = std::getenv("SLSGP_CHOL_PANEL")) ctx->chol_panel = std::min(16, std::max(2, std::atoi(e)));
        if (const char* e = std::getenv("SLSGP_CHOL_LOOKAHEAD")) ctx->chol_look_ahead = std::atoi(e) != 0;
        if (const char* e = std::getenv("SLSGP_CHOL_SWITCH_REM")) ctx->chol_switch_rem = std::max(0, std::atoi(e));
        if (const char* e = std::getenv("SLSGP_CHOL_PIVOTS")) ctx->chol_pivots = std::atoi(e);
        if (const char* e = std::getenv("SLSGP_REFINE_TAU")) ctx->refine_tau = std::min(1.0, std::max(0.0, std::atof(e)));
        ctx->pinned_bytes = 1 << 17; // second half: results of the one-synchronisation MAP objective (kMapOutOffset)
        ok                = ok && cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) == cudaSuccess;
        if (!ok)
        {
            cudaGetLastError();
            ctx->pinned = ok ? ctx->pinned : nullptr;
            slsgp_ctx_destroy(ctx);
            return SLSGP_ERR_CUDA;
        }
        *ctx_out = ctx;
        return SLSGP_OK;
    }

    slsgp_status slsgp_ctx_create_multi(const int* device_ids, int n_devices, slsgp_ctx** ctx_out)
    {
        if (!ctx_out) return SLSGP_ERR_INVALID;
        *ctx_out = nullptr;
        if (!device_ids || n_devices <= 0) return SLSGP_ERR_INVALID;
        for (int i = 0; i < n_devices; ++i)
            for (int j = 0; j < i; ++j)
                if (device_ids[i] == device_ids[j]) return SLSGP_ERR_INVALID; // one context per device
        slsgp_ctx*   primary = nullptr;
        slsgp_status st      = slsgp_ctx_create(device_ids[0], &primary);
        if (st != SLSGP_OK) return st;
        for (int i = 1; i < n_devices; ++i)
        {
            slsgp_ctx* peer = nullptr;
            st              = slsgp_ctx_create(device_ids[i], &peer);
            if (st != SLSGP_OK)
            {
                slsgp_ctx_destroy(primary);
                return st;
            }
            primary->peers.push_back(peer);
        }
        cudaSetDevice(device_ids[0]);
        *ctx_out = primary;
        return SLSGP_OK;
    }

    int slsgp_ctx_device_count(const slsgp_ctx* ctx) { return ctx ? 1 + (int) ctx->peers.size() : 0; }

    slsgp_status slsgp_ctx_destroy(slsgp_ctx* ctx)
    {
        if (!ctx) return SLSGP_OK;
        for (slsgp_ctx* peer : ctx->peers) slsgp_ctx_destroy(peer);
        ctx->peers.clear();
        cudaSetDevice(ctx->device);
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
        DevBuf* all[] = {&ctx->X, &ctx->Xpad, &ctx->XT1, &ctx->theta, &ctx->inv_l, &ctx->K, &ctx->L, &ctx->W,
                         &ctx->Kinv, &ctx->T, &ctx->y, &ctx->alpha, &ctx->Kalpha, &ctx->vec, &ctx->scalars,
                         &ctx->info, &ctx->fbest, &ctx->fbest_idx, &ctx->pref_off, &ctx->pref_idx, &ctx->slot_off,
                         &ctx->slot_list, &ctx->loglik, &ctx->contrib, &ctx->grad_y, &ctx->Ymat, &ctx->g_l, &ctx->Xq,
                         &ctx->Kstar, &ctx->Gstar, &ctx->Beta, &ctx->P1, &ctx->P2, &ctx->stats, &ctx->o_mu,
                         &ctx->o_sigma, &ctx->o_dmu, &ctx->o_dsigma, &ctx->o_val, &ctx->o_grad, &ctx->am_part,
                         &ctx->am_acc, &ctx->chol_flags, &ctx->Bmat, &ctx->Xt, &ctx->Xs32, &ctx->tcs, &ctx->Ks, &ctx->tc_err, &ctx->comb, &ctx->tc_qx, &ctx->tc_P2x, &ctx->mx_best, &ctx->mx_X, &ctx->mx_Xbest, &ctx->mx_Gbest, &ctx->mx_state, &ctx->mx_val, &ctx->mx_grad, &ctx->rf_count, &ctx->rf_index, &ctx->rf_X, &ctx->rf_out};
        for (DevBuf* b : all)
            if (b->p) cudaFree(b->p);
        for (auto& kv : ctx->phases)
        {
            if (kv.second.start) cudaEventDestroy(kv.second.start);
            if (kv.second.stop) cudaEventDestroy(kv.second.stop);
        }
        for (auto& kv : ctx->prof)
            for (auto& r : kv.second) ctx->prof_free.push_back(r);
        for (auto& r : ctx->prof_free) cudaEventDestroy(r.start), cudaEventDestroy(r.stop);
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        if (ctx->pin_in) cudaFreeHost(ctx->pin_in);
        if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
        for (int b = 0; b < 2; ++b)
        {
            if (ctx->ev_in[b]) cudaEventDestroy(ctx->ev_in[b]);
            if (ctx->ev_main[b]) cudaEventDestroy(ctx->ev_main[b]);
            if (ctx->ev_out[b]) cudaEventDestroy(ctx->ev_out[b]);
        }
        if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
        for (int b = 0; b < 3; ++b)
            if (ctx->ev_chol[b]) cudaEventDestroy(ctx->ev_chol[b]);
        if (ctx->chol_hi) cudaStreamDestroy(ctx->chol_hi);
        if (ctx->chol_side) cudaStreamDestroy(ctx->chol_side);
        if (ctx->pre_stream) cudaStreamDestroy(ctx->pre_stream);
        if (ctx->post_stream) cudaStreamDestroy(ctx->post_stream);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return SLSGP_OK;
    }

    // Release the per-shard sweep workspaces (the buffers that only ever grow with the largest batch seen) when they hold more
    // than keep_bytes; the model itself (X, K, L, W, K^-1, alpha) stays. They are re-allocated by the next sweep.
    slsgp_status slsgp_trim(slsgp_ctx* ctx, size_t keep_bytes)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        DevBuf* scratch[] = {&ctx->Xq, &ctx->Kstar, &ctx->Gstar, &ctx->Beta, &ctx->P1, &ctx->P2, &ctx->stats, &ctx->o_mu, &ctx->o_sigma, &ctx->o_dmu,
                             &ctx->o_dsigma, &ctx->o_val, &ctx->o_grad, &ctx->Ks, &ctx->comb, &ctx->tc_qx, &ctx->tc_P2x, &ctx->mx_best, &ctx->mx_X,
                             &ctx->mx_Xbest, &ctx->mx_Gbest, &ctx->mx_state, &ctx->mx_val, &ctx->mx_grad, &ctx->rf_index, &ctx->rf_X, &ctx->rf_out};
        size_t  total     = 0;
        for (DevBuf* b : scratch) total += b->bytes;
        if (total <= keep_bytes) return SLSGP_OK;
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        for (DevBuf* b : scratch)
        {
            if (b->p) cudaFree(b->p);
            b->p = nullptr, b->bytes = 0;
        }
        ctx->Mcap = 0, ctx->tc_Mcap = 0;
        return SLSGP_OK;
    }

    const char* slsgp_last_error(const slsgp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

    slsgp_status slsgp_set_compat_flags(slsgp_ctx* ctx, unsigned flags)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        ctx->compat = flags;
        return SLSGP_OK;
    }

    slsgp_status slsgp_set_sweep_mode(slsgp_ctx* ctx, slsgp_sweep_mode mode)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (mode != SLSGP_SWEEP_FP64 && !is_tensor_mode(mode)) return fail(ctx, SLSGP_ERR_INVALID, "unknown sweep mode");
        if (is_tensor_mode(mode) != is_tensor_mode(ctx->sweep_mode)) ctx->Mcap = 0; // the per-shard scratch differs between the two modes
        ctx->sweep_mode = mode;
        return SLSGP_OK;
    }

    slsgp_status slsgp_set_refine_threshold(slsgp_ctx* ctx, double tau)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!(tau >= 0.0) || tau > 1.0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_set_refine_threshold: tau must lie in [0, 1]");
        ctx->refine_tau = tau;
        return SLSGP_OK;
    }

    slsgp_status slsgp_get_sweep_mode(const slsgp_ctx* ctx, slsgp_sweep_mode* mode_out)
    {
        if (!ctx || !mode_out) return SLSGP_ERR_INVALID;
        *mode_out = (slsgp_sweep_mode) ctx->sweep_mode;
        return SLSGP_OK;
    }

    slsgp_status slsgp_set_stream(slsgp_ctx* ctx, void* cuda_stream)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
        return SLSGP_OK;
    }

    slsgp_status slsgp_synchronize(slsgp_ctx* ctx)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    uint64_t slsgp_launch_count(const slsgp_ctx* ctx) { return ctx ? ctx->launches : 0; }

    double slsgp_last_phase_ms(const slsgp_ctx* ctx, const char* phase)
    {
        if (!ctx || !phase) return -1.0;
        auto it = ctx->phases.find(phase);
        if (it == ctx->phases.end() || !it->second.valid) return -1.0;
        if (cudaEventSynchronize(it->second.stop) != cudaSuccess) return -1.0;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, it->second.start, it->second.stop) != cudaSuccess) return -1.0;
        return (double) ms;
    }

    slsgp_status slsgp_profile_enable(slsgp_ctx* ctx, int on)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        ctx->profile = on != 0;
        return SLSGP_OK;
    }

    slsgp_status slsgp_profile_read(slsgp_ctx* ctx, const char* kernel, double* total_ms_out, uint64_t* launches_out)
    {
        if (!ctx || !kernel) return SLSGP_ERR_INVALID;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        double   total = 0.0;
        uint64_t n     = 0;
        auto     it    = ctx->prof.find(kernel);
        if (it != ctx->prof.end())
        {
            for (auto& r : it->second)
            {
                float ms = 0.f;
                CUDA_TRY(cudaEventElapsedTime(&ms, r.start, r.stop));
                total += ms, ++n;
                ctx->prof_free.push_back(r);
            }
            it->second.clear();
        }
        if (total_ms_out) *total_ms_out = total;
        if (launches_out) *launches_out = n;
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    slsgp_status slsgp_set_data(slsgp_ctx* ctx, const double* X, int N, int D)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!X || N <= 0 || D <= 0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_set_data: X null or N, D not positive");
        if ((size_t) (2 * D + 2) * sizeof(double) > ctx->pinned_bytes)
            return fail(ctx, SLSGP_ERR_INVALID, "slsgp_set_data: D too large");
        for (size_t i = 0; i < (size_t) N * D; ++i)
            if (!std::isfinite(X[i])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_set_data: non-finite coordinate in X");
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (N != ctx->N) ctx->P = 0, ctx->pref_total = 0; // tuples index the columns of X: a new N voids them
        ctx->N = N, ctx->D = D, ctx->ld = round_up(N, TILE), ctx->Dp = round_up(D, TILE), ctx->ldx = round_up(D + 1, TILE);
        ctx->has_data = ctx->has_gram = ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = false;
        ctx->Mcap = 0; // sweep workspace depends on ld
        ctx->tc_Mcap = 0, ctx->tc_ready = false;
        TRY(ensure_model_buffers(ctx));
        const size_t ld = ctx->ld;
        CUDA_TRY(cudaMemcpyAsync(ctx->X.p, X, sizeof(double) * (size_t) N * D, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemsetAsync(ctx->y.p, 0, sizeof(double) * ld, ctx->stream));
        pack_x_kernel<<<(ctx->ld + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->X), N, D, ctx->ld, ctx->Dp, ctx->ldx,
                                                                     dp(ctx->Xpad), dp(ctx->XT1));
        LAUNCH_CHECK();
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->X_host.assign(X, X + (size_t) N * D);
        ctx->has_data = true;
        return SLSGP_OK;
    }

    // rows / columns >= p of an ld x ld matrix back to the identity padding every dense kernel relies on
    __global__ void truncate_to_identity_kernel(double* __restrict__ A, double* __restrict__ B, double* __restrict__ C, int p, int ld)
    {
        const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
        if (i >= ld || (i < p && j < p)) return;
        const double v = i == j ? 1.0 : 0.0;
        const size_t e = (size_t) i + (size_t) j * ld;
        A[e] = v, B[e] = v, C[e] = v;
    }

    slsgp_status slsgp_invalidate(slsgp_ctx* ctx)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        ctx->has_gram = ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = false;
        ctx->tc_ready = false;
        return SLSGP_OK;
    }

    slsgp_status slsgp_set_data_extend(slsgp_ctx* ctx, const double* X, int N, int D, int* n_kept_out)
    {
        if (n_kept_out) *n_kept_out = 0;
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!X || N <= 0 || D <= 0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_set_data_extend: X null or N, D not positive");
        const int N_old = ctx->N;
        int       p     = 0; // leading data points shared, bit for bit, by the model and the new data
        if (ctx->has_data && ctx->has_factor && D == ctx->D && N <= ctx->ld && ctx->X_host.size() == (size_t) N_old * D)
            while (p < std::min(N, N_old) && std::memcmp(X + (size_t) p * D, ctx->X_host.data() + (size_t) p * D, sizeof(double) * (size_t) D) == 0) ++p;
        // each appended point is one bordered update (a handful of small launches); beyond a few of them the rebuild is cheaper
        static const int max_new = std::getenv("SLSGP_EXTEND_MAX_NEW") ? std::atoi(std::getenv("SLSGP_EXTEND_MAX_NEW")) : 8;
        if (p == 0 || N - p > max_new) return slsgp_set_data(ctx, X, N, D);
        CUDA_TRY(cudaSetDevice(ctx->device));
        ctx->has_alpha = false; // the observations go: no alpha update per appended point
        if (p < N_old)
        {
            // Points beyond the common prefix have changed (the data manager merges coincident points: both leave their columns and
            // the midpoint is appended, src/preference-data-manager.cpp:14-86). The leading p x p blocks of K_y, L and L^-1 ARE the
            // model of the first p points; K_y^-1 is rebuilt from L^-1 on demand.
            const int ld = ctx->ld;
            TRY(do_trtri(ctx));
            truncate_to_identity_kernel<<<dim3((ld + 255) / 256, ld), 256, 0, ctx->stream>>>(dp(ctx->K), dp(ctx->L), dp(ctx->W), p, ld);
            LAUNCH_CHECK();
            logdet_kernel<<<1, 256, 0, ctx->stream>>>(dp(ctx->L), p, ld, dp(ctx->scalars) + 8);
            LAUNCH_CHECK();
            double logdet = 0.0;
            CUDA_TRY(cudaMemcpyAsync(&logdet, dp(ctx->scalars) + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            ctx->logdet_host = logdet;
            ctx->N           = p;
            ctx->X_host.resize((size_t) p * D);
            ctx->has_inverse = false, ctx->tc_ready = false;
            ctx->P = 0, ctx->pref_total = 0;
            if (p == N) // nothing to append: the packed copies of X still describe N_old points
            {
                pack_x_kernel<<<(ld + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->X), ctx->N, D, ld, ctx->Dp, ctx->ldx, dp(ctx->Xpad), dp(ctx->XT1));
                LAUNCH_CHECK();
            }
        }
        for (int i = p; i < N; ++i)
        {
            const slsgp_status s = slsgp_append_point(ctx, X + (size_t) i * D, 0.0, nullptr, nullptr);
            if (s == SLSGP_ERR_NOT_SPD) return slsgp_set_data(ctx, X, N, D); // e.g. a duplicated point with zero noise: let the rebuild report it
            if (s != SLSGP_OK) return s;
        }
        // as after slsgp_set_data: no observations
        CUDA_TRY(cudaMemsetAsync(ctx->y.p, 0, sizeof(double) * (size_t) ctx->ld, ctx->stream));
        ctx->has_alpha = false;
        if (n_kept_out) *n_kept_out = p;
        return SLSGP_OK;
    }

    // K_y depends on (X, kernel, theta, noise) only. X changes through slsgp_set_data (clears has_gram) or slsgp_append_point /
    // slsgp_set_data_extend (which update K_y, L, L^-1 and K_y^-1 consistently), and every internal rebuild (the MAP objectives)
    // records its hyper-parameters in the context, so a request for the matrix the context already holds is answered from it.
    static bool gram_is_current(const slsgp_ctx* ctx, int kernel_type, const double* theta, double noise)
    {
        if (!ctx->has_gram || kernel_type != ctx->kernel_type || noise != ctx->noise || (int) ctx->theta_host.size() != ctx->D + 1) return false;
        return std::memcmp(theta, ctx->theta_host.data(), sizeof(double) * (size_t) (ctx->D + 1)) == 0;
    }

    slsgp_status slsgp_gram(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* theta, double noise,
                            double* K_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!theta) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_gram: theta is null");
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (!gram_is_current(ctx, (int) kernel_type, theta, noise)) TRY(do_gram(ctx, (int) kernel_type, theta, noise));
        return copy_matrix_out(ctx, ctx->K, K_out, ctx->N);
    }

    slsgp_status slsgp_factor(slsgp_ctx* ctx, double* logdet_out, double* L_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (ctx->has_gram && ctx->has_factor) // the factor of the current K_y (do_gram clears has_factor)
        {
            if (logdet_out) *logdet_out = ctx->logdet_host;
        }
        else
            TRY(do_factor(ctx, logdet_out));
        return copy_matrix_out(ctx, ctx->L, L_out, ctx->N);
    }

    slsgp_status slsgp_inverse(slsgp_ctx* ctx, double* Kinv_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        CUDA_TRY(cudaSetDevice(ctx->device));
        TRY(do_inverse(ctx));
        return copy_matrix_out(ctx, ctx->Kinv, Kinv_out, ctx->N);
    }

    slsgp_status slsgp_solve_alpha(slsgp_ctx* ctx, const double* y, double* alpha_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!y) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_solve_alpha: y is null");
        if (!ctx->has_factor) return fail(ctx, SLSGP_ERR_STATE, "slsgp_solve_alpha before slsgp_factor");
        for (int i = 0; i < ctx->N; ++i)
            if (!std::isfinite(y[i])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_solve_alpha: non-finite y");
        CUDA_TRY(cudaSetDevice(ctx->device));
        CUDA_TRY(cudaMemcpyAsync(ctx->y.p, y, sizeof(double) * ctx->N, cudaMemcpyHostToDevice, ctx->stream));
        TRY(do_alpha(ctx));
        if (alpha_out)
            CUDA_TRY(cudaMemcpyAsync(alpha_out, ctx->alpha.p, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    slsgp_status slsgp_append_point(slsgp_ctx* ctx, const double* x, double y_new, double* K_col_out, double* Kinv_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!x) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_append_point: x is null");
        if (!ctx->has_factor) return fail(ctx, SLSGP_ERR_STATE, "slsgp_append_point before slsgp_factor");
        const int N = ctx->N, D = ctx->D, ld = ctx->ld;
        for (int d = 0; d < D; ++d)
            if (!std::isfinite(x[d])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_append_point: non-finite coordinate");
        if (!std::isfinite(y_new)) return fail(ctx, SLSGP_ERR_NAN, "slsgp_append_point: non-finite y");
        CUDA_TRY(cudaSetDevice(ctx->device));
        const bool had_alpha = ctx->has_alpha;

        if (N + 1 > ld)
        {
            // the padded leading dimension is exhausted (once every 64 appended points): rebuild at the next size
            std::vector<double> Xh((size_t) (N + 1) * D), yh((size_t) N + 1);
            CUDA_TRY(cudaMemcpyAsync(Xh.data(), ctx->X.p, sizeof(double) * (size_t) N * D, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(yh.data(), ctx->y.p, sizeof(double) * (size_t) N, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            for (int d = 0; d < D; ++d) Xh[(size_t) N * D + d] = x[d];
            yh[(size_t) N]                  = y_new;
            const std::vector<double> theta = ctx->theta_host;
            const int                 kt    = ctx->kernel_type;
            const double              noise = ctx->noise;
            TRY(slsgp_set_data(ctx, Xh.data(), N + 1, D));
            TRY(do_gram(ctx, kt, theta.data(), noise));
            TRY(do_factor(ctx, nullptr));
            TRY(do_inverse(ctx));
            if (had_alpha)
            {
                CUDA_TRY(cudaMemcpyAsync(ctx->y.p, yh.data(), sizeof(double) * (size_t) (N + 1), cudaMemcpyHostToDevice, ctx->stream));
                TRY(do_alpha(ctx));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            }
        }
        else
        {
            TRY(do_inverse(ctx)); // W and Kinv of the current model
            TRY(phase_begin(ctx, "append"));
            double* h = ctx->pinned;
            for (int d = 0; d < D; ++d) h[d] = x[d];
            h[D] = y_new;
            CUDA_TRY(cudaMemcpyAsync(dp(ctx->X) + (size_t) N * D, h, sizeof(double) * D, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(ctx->info.p, 0, sizeof(int), ctx->stream));
            double *k = dp(ctx->vec), *l = dp(ctx->Kalpha), *u = dp(ctx->grad_y);
            append_kvec_kernel<<<(ld + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->X), N, D, ld, dp(ctx->theta), dp(ctx->inv_l), ctx->noise,
                                                                         ctx->kernel_type, k, dp(ctx->scalars));
            LAUNCH_CHECK();
            const int blocks = (ld * 32 + 255) / 256;
            gemv_kernel<false><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->W), ld, ld, k, l, 1); // l = W k   (rows j <= i)
            LAUNCH_CHECK();
            gemv_kernel<true><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->W), ld, ld, l, u, 1);  // u = W^T l (columns j >= i)
            LAUNCH_CHECK();
            append_rows_kernel<<<1, 1024, 0, ctx->stream>>>(N, ld, l, u, dp(ctx->L), dp(ctx->W), dp(ctx->scalars), ptr<int>(ctx->info));
            LAUNCH_CHECK();
            int    info = 0;
            double s    = 0.0;
            CUDA_TRY(cudaMemcpyAsync(&info, ctx->info.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(&s, dp(ctx->scalars) + 10, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (info != 0)
                return fail(ctx, SLSGP_ERR_NOT_SPD,
                            "slsgp_append_point: non-positive Schur complement " + std::to_string(s) + " (the model is unchanged)");
            append_commit_kernel<<<dim3((N + 1 + 255) / 256, N + 1), 256, 0, ctx->stream>>>(N, ld, k, u, dp(ctx->scalars), dp(ctx->K),
                                                                                          dp(ctx->Kinv));
            LAUNCH_CHECK();
            CUDA_TRY(cudaMemcpyAsync(dp(ctx->y) + N, h + D, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            ctx->N = N + 1;
            ctx->logdet_host += std::log(s);
            ctx->P = 0, ctx->pref_total = 0; // the per-point tuple lists were built for N points
            ctx->tc_ready  = false;
            ctx->has_alpha = false;
            pack_x_kernel<<<(ld + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->X), ctx->N, D, ld, ctx->Dp, ctx->ldx, dp(ctx->Xpad), dp(ctx->XT1));
            LAUNCH_CHECK();
            TRY(phase_end(ctx, "append"));
            if (had_alpha) TRY(do_alpha(ctx));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream)); // the pinned staging area is reused by later calls
        }
        if (ctx->X_host.size() == (size_t) N * D) ctx->X_host.insert(ctx->X_host.end(), x, x + D); // (the rebuild path went through slsgp_set_data)
        if (K_col_out)
            CUDA_TRY(cudaMemcpyAsync(K_col_out, dp(ctx->K) + (size_t) (ctx->N - 1) * ctx->ld, sizeof(double) * (size_t) ctx->N,
                                     cudaMemcpyDeviceToHost, ctx->stream));
        if (K_col_out) CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return copy_matrix_out(ctx, ctx->Kinv, Kinv_out, ctx->N);
    }

    slsgp_status slsgp_get_f_best(slsgp_ctx* ctx, double* f_best_out, int* index_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        TRY(require_model(ctx));
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (f_best_out)
            CUDA_TRY(cudaMemcpyAsync(f_best_out, ctx->fbest.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (index_out)
            CUDA_TRY(cudaMemcpyAsync(index_out, ctx->fbest_idx.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // K4
    // ---------------------------------------------------------------------------------------------------------
    slsgp_status slsgp_acq_batch_device(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta,
                                        const double* d_Xq, int64_t M, double* d_mu, double* d_sigma,
                                        double* d_dmu, double* d_dsigma, double* d_val, double* d_grad)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (M < 0 || (M > 0 && !d_Xq)) return fail(ctx, SLSGP_ERR_INVALID, "acq_batch: bad M or null Xq");
        if (acq_type != SLSGP_ACQ_EXPECTED_IMPROVEMENT && acq_type != SLSGP_ACQ_GP_UCB)
            return fail(ctx, SLSGP_ERR_INVALID, "unknown acq_type");
        TRY(require_model(ctx));
        CUDA_TRY(cudaSetDevice(ctx->device));
        TRY(ensure_sweep_workspace(ctx, M));
        TRY(phase_begin(ctx, "sweep"));
        SweepJob job;
        job.acq_type = (int) acq_type, job.ucb_beta = ucb_beta, job.M = M, job.d_Xq = d_Xq;
        job.mu = d_mu, job.sigma = d_sigma, job.dmu = d_dmu, job.dsigma = d_dsigma, job.val = d_val, job.grad = d_grad;
        TRY(run_sweep(ctx, job));
        TRY(phase_end(ctx, "sweep"));
        return SLSGP_OK;
    }

    // Host-buffer form shared by slsgp_posterior_batch and slsgp_acq_batch: per shard, H2D the candidates, sweep,
    // D2H whichever outputs were asked for.
    static slsgp_status host_sweep(slsgp_ctx* ctx, int acq_type, double ucb_beta, const double* Xq, int64_t M,
                                   double* mu, double* sigma, double* dmu, double* dsigma, double* val, double* grad)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (M < 0 || (M > 0 && !Xq)) return fail(ctx, SLSGP_ERR_INVALID, "batch: bad M or null Xq");
        if (acq_type != SLSGP_ACQ_EXPECTED_IMPROVEMENT && acq_type != SLSGP_ACQ_GP_UCB)
            return fail(ctx, SLSGP_ERR_INVALID, "unknown acq_type");
        TRY(require_model(ctx));
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (!ctx->peers.empty() && M >= group_min_count())
        {
            // the group: contiguous parts of the host batch, each through the host-buffer path of its own device
            const std::vector<GroupShare> shares = split_over_group(ctx, 0, M);
            const size_t                  D = (size_t) ctx->D;
            return run_on_group(ctx, shares, [&](const GroupShare& sh, size_t) {
                slsgp_ctx* c = sh.ctx;
                std::vector<slsgp_ctx*> none;
                none.swap(c->peers);
                const size_t       o  = (size_t) sh.first;
                const slsgp_status st = host_sweep(c, acq_type, ucb_beta, Xq + o * D, sh.count, mu ? mu + o : nullptr, sigma ? sigma + o : nullptr,
                                                   dmu ? dmu + o * D : nullptr, dsigma ? dsigma + o * D : nullptr, val ? val + o : nullptr,
                                                   grad ? grad + o * D : nullptr);
                none.swap(c->peers);
                return st;
            });
        }
        TRY(ensure_sweep_workspace(ctx, M));
        TRY(phase_begin(ctx, "sweep"));
        SweepJob job;
        job.acq_type = acq_type, job.ucb_beta = ucb_beta, job.M = M, job.h_Xq = Xq, job.host_out = true;
        job.mu = mu, job.sigma = sigma, job.dmu = dmu, job.dsigma = dsigma, job.val = val, job.grad = grad;
        TRY(run_sweep(ctx, job));
        TRY(phase_end(ctx, "sweep"));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    slsgp_status slsgp_posterior_batch(slsgp_ctx* ctx, const double* Xq, int64_t M, double* mu_out,
                                       double* sigma_out, double* dmu_out, double* dsigma_out)
    {
        return host_sweep(ctx, SLSGP_ACQ_EXPECTED_IMPROVEMENT, 0.0, Xq, M, mu_out, sigma_out, dmu_out, dsigma_out,
                          nullptr, nullptr);
    }

    slsgp_status slsgp_acq_batch(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, const double* Xq,
                                 int64_t M, double* val_out, double* grad_out)
    {
        return host_sweep(ctx, (int) acq_type, ucb_beta, Xq, M, nullptr, nullptr, nullptr, nullptr, val_out, grad_out);
    }

    slsgp_status slsgp_acq_from_posterior(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, double f_best, int D,
                                          int64_t M, const double* mu, const double* sigma, const double* dmu,
                                          const double* dsigma, double* val_out, double* grad_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (M < 0 || D <= 0 || (M > 0 && (!mu || !sigma)) || (grad_out && (!dmu || !dsigma)))
            return fail(ctx, SLSGP_ERR_INVALID, "slsgp_acq_from_posterior: bad arguments");
        if (acq_type != SLSGP_ACQ_EXPECTED_IMPROVEMENT && acq_type != SLSGP_ACQ_GP_UCB)
            return fail(ctx, SLSGP_ERR_INVALID, "unknown acq_type");
        if (M == 0) return SLSGP_OK;
        CUDA_TRY(cudaSetDevice(ctx->device));
        // scratch: [mu | sigma | val] (M each) and, with gradients, [dmu | dsigma | grad] (D x M each)
        const size_t sv = sizeof(double) * (size_t) M, sg = sv * D;
        TRY(ensure(ctx, ctx->comb, 3 * sv + (grad_out ? 3 * sg : 0)));
        double *d_mu = dp(ctx->comb), *d_sigma = d_mu + M, *d_val = d_sigma + M;
        double *d_dmu = d_val + M, *d_dsigma = d_dmu + (size_t) M * D, *d_grad = d_dsigma + (size_t) M * D;
        CUDA_TRY(cudaMemcpyAsync(d_mu, mu, sv, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(d_sigma, sigma, sv, cudaMemcpyHostToDevice, ctx->stream));
        if (grad_out)
        {
            CUDA_TRY(cudaMemcpyAsync(d_dmu, dmu, sg, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(d_dsigma, dsigma, sg, cudaMemcpyHostToDevice, ctx->stream));
        }
        acq_combine_kernel<<<(unsigned) ((M + 255) / 256), 256, 0, ctx->stream>>>(
            d_mu, d_sigma, d_dmu, d_dsigma, D, M, f_best, (int) acq_type, ucb_beta, val_out ? d_val : nullptr,
            grad_out ? d_grad : nullptr);
        LAUNCH_CHECK();
        if (val_out) CUDA_TRY(cudaMemcpyAsync(val_out, d_val, sv, cudaMemcpyDeviceToHost, ctx->stream));
        if (grad_out) CUDA_TRY(cudaMemcpyAsync(grad_out, d_grad, sg, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    // Schonlau's batch criterion as one device-resident arg-max: candidate i of [first, first + count) from the counter-based
    // generator, mu from `ctx` (the original model), sigma from `ctx_sigma` (the model that already holds the pending points),
    // the acquisition formula, a running arg-max; nothing but the winner leaves the device. Chunks of 2^17 candidates.
    slsgp_status slsgp_pair_acq_argmax(slsgp_ctx* ctx, slsgp_ctx* ctx_sigma, slsgp_acq_type acq_type, double ucb_beta, uint64_t seed,
                                       int64_t first, int64_t count, double* x_best_out, double* val_best_out, int64_t* index_best_out)
    {
        if (!ctx || !ctx_sigma) return SLSGP_ERR_INVALID;
        if (count <= 0 || first < 0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_pair_acq_argmax: empty candidate range");
        if (acq_type != SLSGP_ACQ_EXPECTED_IMPROVEMENT && acq_type != SLSGP_ACQ_GP_UCB) return fail(ctx, SLSGP_ERR_INVALID, "unknown acq_type");
        if (ctx->device != ctx_sigma->device) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_pair_acq_argmax: both models must live on one device");
        if (ctx->D != ctx_sigma->D) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_pair_acq_argmax: the two models differ in dimension");
        TRY(require_model(ctx));
        {
            const slsgp_status st = require_model(ctx_sigma);
            if (st != SLSGP_OK) return fail(ctx, st, "slsgp_pair_acq_argmax: " + ctx_sigma->err);
        }
        CUDA_TRY(cudaSetDevice(ctx->device));
        const int       D = ctx->D;
        const long long chunk = std::min<long long>(count, 1LL << 17);
        // scratch in the first context: candidates (D x chunk) | mu | sigma | val
        TRY(ensure(ctx, ctx->comb, sizeof(double) * (size_t) (D + 3) * chunk));
        double *d_Xq = dp(ctx->comb), *d_mu = d_Xq + (size_t) D * chunk, *d_sigma = d_mu + chunk, *d_val = d_sigma + chunk;
        TRY(ensure(ctx, ctx->am_part, sizeof(ArgMax) * 1024));
        TRY(ensure(ctx, ctx->am_acc, sizeof(ArgMax)));
        ArgMax init;
        init.v = 0.0, init.i = -1;
        std::memcpy(ctx->pinned, &init, sizeof(init));
        CUDA_TRY(cudaMemcpyAsync(ctx->am_acc.p, ctx->pinned, sizeof(ArgMax), cudaMemcpyHostToDevice, ctx->stream));
        double f_best = 0.0;
        CUDA_TRY(cudaMemcpyAsync(&f_best, ctx->fbest.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        for (long long m0 = 0; m0 < count; m0 += chunk)
        {
            const long long Mc = std::min(chunk, count - m0);
            candidates_kernel<<<(unsigned) ((Mc * D + 255) / 256), 256, 0, ctx->stream>>>(seed, first + m0, Mc, D, d_Xq);
            LAUNCH_CHECK();
            TRY(slsgp_acq_batch_device(ctx, acq_type, ucb_beta, d_Xq, Mc, d_mu, nullptr, nullptr, nullptr, nullptr, nullptr));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream)); // the candidates are complete before the second context reads them
            {
                const slsgp_status st = slsgp_acq_batch_device(ctx_sigma, acq_type, ucb_beta, d_Xq, Mc, nullptr, d_sigma, nullptr, nullptr, nullptr, nullptr);
                if (st != SLSGP_OK) return fail(ctx, st, "slsgp_pair_acq_argmax: " + ctx_sigma->err);
                CUDA_TRY(cudaStreamSynchronize(ctx_sigma->stream));
            }
            acq_combine_kernel<<<(unsigned) ((Mc + 255) / 256), 256, 0, ctx->stream>>>(d_mu, d_sigma, nullptr, nullptr, D, Mc, f_best, (int) acq_type, ucb_beta,
                                                                                    d_val, nullptr);
            LAUNCH_CHECK();
            const int nblk = (int) std::min<long long>(1024, (Mc + 255) / 256);
            argmax_partial_kernel<<<nblk, 256, 0, ctx->stream>>>(d_val, Mc, first + m0, ptr<ArgMax>(ctx->am_part));
            LAUNCH_CHECK();
            argmax_final_kernel<<<1, 256, 0, ctx->stream>>>(ptr<ArgMax>(ctx->am_part), nblk, ptr<ArgMax>(ctx->am_acc));
            LAUNCH_CHECK();
        }
        ArgMax best;
        CUDA_TRY(cudaMemcpyAsync(&best, ctx->am_acc.p, sizeof(ArgMax), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (best.i < 0) return fail(ctx, SLSGP_ERR_NAN, "slsgp_pair_acq_argmax: every candidate evaluated to NaN");
        if (x_best_out)
            for (int d = 0; d < D; ++d) x_best_out[d] = candidate_coord(seed, best.i, d);
        if (val_best_out) *val_best_out = best.v;
        if (index_best_out) *index_best_out = best.i;
        return SLSGP_OK;
    }

    slsgp_status slsgp_argmax_device(slsgp_ctx* ctx, const double* d_val, int64_t count, int64_t index0,
                                     double* val_best_out, int64_t* index_best_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!d_val || count <= 0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_argmax_device: null values or empty range");
        CUDA_TRY(cudaSetDevice(ctx->device));
        TRY(ensure(ctx, ctx->am_part, sizeof(ArgMax) * 1024));
        TRY(ensure(ctx, ctx->am_acc, sizeof(ArgMax)));
        ArgMax init;
        init.v = 0.0, init.i = -1;
        std::memcpy(ctx->pinned, &init, sizeof(init));
        CUDA_TRY(cudaMemcpyAsync(ctx->am_acc.p, ctx->pinned, sizeof(ArgMax), cudaMemcpyHostToDevice, ctx->stream));
        const int nblk = (int) std::min<long long>(1024, (count + 255) / 256);
        argmax_partial_kernel<<<nblk, 256, 0, ctx->stream>>>(d_val, count, index0, ptr<ArgMax>(ctx->am_part));
        LAUNCH_CHECK();
        argmax_final_kernel<<<1, 256, 0, ctx->stream>>>(ptr<ArgMax>(ctx->am_part), nblk, ptr<ArgMax>(ctx->am_acc));
        LAUNCH_CHECK();
        ArgMax best;
        CUDA_TRY(cudaMemcpyAsync(&best, ctx->am_acc.p, sizeof(ArgMax), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (best.i < 0) return fail(ctx, SLSGP_ERR_NAN, "slsgp_argmax_device: every value is NaN");
        if (val_best_out) *val_best_out = best.v;
        if (index_best_out) *index_best_out = best.i;
        return SLSGP_OK;
    }

    slsgp_status slsgp_candidates(slsgp_ctx* ctx, uint64_t seed, int64_t first, int64_t count, double* Xq_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_data || !Xq_out || count < 0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_candidates: bad arguments");
        for (int64_t i = 0; i < count; ++i)
            for (int d = 0; d < ctx->D; ++d) Xq_out[(size_t) i * ctx->D + d] = candidate_coord(seed, first + i, d);
        return SLSGP_OK;
    }

    slsgp_status slsgp_acq_argmax(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, uint64_t seed,
                                  int64_t first, int64_t count, double* x_best_out, double* val_best_out,
                                  int64_t* index_best_out, double* grad_best_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (count <= 0 || first < 0) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_acq_argmax: empty candidate range");
        if (acq_type != SLSGP_ACQ_EXPECTED_IMPROVEMENT && acq_type != SLSGP_ACQ_GP_UCB)
            return fail(ctx, SLSGP_ERR_INVALID, "unknown acq_type");
        TRY(require_model(ctx));
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (!ctx->peers.empty() && count >= group_min_count())
        {
            // the group: every device takes a contiguous part of the candidate range; winners reduced on the host
            const std::vector<GroupShare> shares = split_over_group(ctx, first, count);
            const int                     D      = ctx->D;
            std::vector<double>           val(shares.size(), 0.0), xs(shares.size() * (size_t) D, 0.0);
            std::vector<int64_t>          idx(shares.size(), -1);
            TRY(run_on_group(ctx, shares, [&](const GroupShare& sh, size_t g) {
                slsgp_ctx* c = sh.ctx;
                std::vector<slsgp_ctx*> none;
                none.swap(c->peers); // the member runs the single-device path
                const slsgp_status st = slsgp_acq_argmax(c, acq_type, ucb_beta, seed, sh.first, sh.count, &xs[g * (size_t) D], &val[g], &idx[g], nullptr);
                none.swap(c->peers);
                return st;
            }));
            size_t best = 0;
            for (size_t g = 1; g < shares.size(); ++g)
                if (val[g] > val[best] || (val[g] == val[best] && idx[g] < idx[best])) best = g;
            if (x_best_out) std::memcpy(x_best_out, &xs[best * (size_t) D], sizeof(double) * (size_t) D);
            if (val_best_out) *val_best_out = val[best];
            if (index_best_out) *index_best_out = idx[best];
            if (grad_best_out) TRY(slsgp_acq_batch(ctx, acq_type, ucb_beta, &xs[best * (size_t) D], 1, nullptr, grad_best_out));
            return SLSGP_OK;
        }
        TRY(ensure_sweep_workspace(ctx, count));
        const int D = ctx->D;
        ArgMax    init;
        init.v = 0.0, init.i = -1;
        std::memcpy(ctx->pinned, &init, sizeof(init));
        CUDA_TRY(cudaMemcpyAsync(ctx->am_acc.p, ctx->pinned, sizeof(ArgMax), cudaMemcpyHostToDevice, ctx->stream));
        TRY(phase_begin(ctx, "sweep"));
        SweepJob job; // values only: the winner's gradient is evaluated once at the end
        job.acq_type = (int) acq_type, job.ucb_beta = ucb_beta, job.M = count, job.generate = true, job.seed = seed;
        job.first = first, job.argmax = true;
        TRY(run_sweep(ctx, job));
        TRY(phase_end(ctx, "sweep"));
        ArgMax best;
        CUDA_TRY(cudaMemcpyAsync(&best, ctx->am_acc.p, sizeof(ArgMax), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (best.i < 0) return fail(ctx, SLSGP_ERR_NAN, "slsgp_acq_argmax: every candidate evaluated to NaN");
        std::vector<double> xb((size_t) D);
        for (int d = 0; d < D; ++d) xb[(size_t) d] = candidate_coord(seed, best.i, d);
        if (x_best_out) std::memcpy(x_best_out, xb.data(), sizeof(double) * (size_t) D);
        if (val_best_out) *val_best_out = best.v;
        if (index_best_out) *index_best_out = best.i;
        if (grad_best_out) TRY(slsgp_acq_batch(ctx, acq_type, ucb_beta, xb.data(), 1, nullptr, grad_best_out));
        return SLSGP_OK;
    }

    slsgp_status slsgp_acq_maximize(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, uint64_t seed, int64_t first,
                                    int64_t count, int n_starts, int n_iters, double* x_best_out, double* val_best_out,
                                    double* grad_best_out, double* val_sweep_best_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (count <= 0 || first < 0 || n_starts <= 0 || n_iters < 0)
            return fail(ctx, SLSGP_ERR_INVALID, "slsgp_acq_maximize: empty candidate range or bad n_starts / n_iters");
        if (acq_type != SLSGP_ACQ_EXPECTED_IMPROVEMENT && acq_type != SLSGP_ACQ_GP_UCB)
            return fail(ctx, SLSGP_ERR_INVALID, "unknown acq_type");
        TRY(require_model(ctx));
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (!ctx->peers.empty() && count >= group_min_count())
        {
            // every device maximises over its own part of the candidate range with its share of the starts; best value wins
            const std::vector<GroupShare> shares = split_over_group(ctx, first, count);
            const int                     D = ctx->D, ng = (int) shares.size();
            std::vector<double>           val(shares.size(), 0.0), vs(shares.size(), 0.0), xs(shares.size() * (size_t) D, 0.0), gs(shares.size() * (size_t) D, 0.0);
            TRY(run_on_group(ctx, shares, [&](const GroupShare& sh, size_t g) {
                slsgp_ctx* c = sh.ctx;
                std::vector<slsgp_ctx*> none;
                none.swap(c->peers);
                const int          starts = std::max(1, n_starts / ng + ((int) g < n_starts % ng ? 1 : 0));
                const slsgp_status st     = slsgp_acq_maximize(c, acq_type, ucb_beta, seed, sh.first, sh.count, starts, n_iters, &xs[g * (size_t) D], &val[g],
                                                               &gs[g * (size_t) D], &vs[g]);
                none.swap(c->peers);
                return st;
            }));
            size_t best = 0;
            double sweep_best = vs[0];
            for (size_t g = 1; g < shares.size(); ++g)
            {
                if (val[g] > val[best]) best = g; // ties: the lower part of the range wins
                sweep_best = std::max(sweep_best, vs[g]);
            }
            if (x_best_out) std::memcpy(x_best_out, &xs[best * (size_t) D], sizeof(double) * (size_t) D);
            if (grad_best_out) std::memcpy(grad_best_out, &gs[best * (size_t) D], sizeof(double) * (size_t) D);
            if (val_best_out) *val_best_out = val[best];
            if (val_sweep_best_out) *val_sweep_best_out = sweep_best;
            return SLSGP_OK;
        }
        const int       D         = ctx->D;
        const long long slice_len = (count + std::min<long long>(std::min<long long>(n_starts, 16384), count) - 1) /
                                    std::min<long long>(std::min<long long>(n_starts, 16384), count);
        const int       K         = (int) ((count + slice_len - 1) / slice_len);
        TRY(ensure(ctx, ctx->mx_best, sizeof(ArgMax) * (size_t) K));
        TRY(ensure(ctx, ctx->mx_X, sizeof(double) * (size_t) K * D));
        TRY(ensure(ctx, ctx->mx_Xbest, sizeof(double) * (size_t) K * D));
        TRY(ensure(ctx, ctx->mx_Gbest, sizeof(double) * (size_t) K * D));
        TRY(ensure(ctx, ctx->mx_grad, sizeof(double) * (size_t) K * D));
        TRY(ensure(ctx, ctx->mx_val, sizeof(double) * (size_t) K));
        TRY(ensure(ctx, ctx->mx_state, sizeof(AscentState) * (size_t) K));
        TRY(ensure(ctx, ctx->am_acc, sizeof(ArgMax)));
        {
            std::vector<ArgMax> init((size_t) K);
            for (auto& a : init) a.v = 0.0, a.i = -1;
            CUDA_TRY(cudaMemcpyAsync(ctx->mx_best.p, init.data(), sizeof(ArgMax) * (size_t) K, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream)); // `init` is pageable and dies with this scope
        }
        TRY(phase_begin(ctx, "maximize"));
        // (1) global stage: dense sweep in the caller's sweep mode, best candidate per slice
        TRY(ensure_sweep_workspace(ctx, count));
        {
            SweepJob job;
            job.acq_type = (int) acq_type, job.ucb_beta = ucb_beta, job.M = count, job.generate = true, job.seed = seed, job.first = first;
            job.slice_len = slice_len, job.slice_best = ptr<ArgMax>(ctx->mx_best);
            TRY(run_sweep(ctx, job));
        }
        starts_from_slices_kernel<<<(K + 127) / 128, 128, 0, ctx->stream>>>(ptr<ArgMax>(ctx->mx_best), K, D, seed, 0.05, dp(ctx->mx_X),
                                                                           dp(ctx->mx_Xbest), ptr<AscentState>(ctx->mx_state));
        LAUNCH_CHECK();
        // (2) local stage in IEEE double whatever the sweep mode: n_iters + 1 evaluations, n_iters moves
        const int    user_mode = ctx->sweep_mode;
        slsgp_status st        = SLSGP_OK;
        ctx->sweep_mode        = SLSGP_SWEEP_FP64;
        if (is_tensor_mode(user_mode)) ctx->Mcap = 0;
        st = ensure_sweep_workspace(ctx, K);
        for (int it = 0; it <= n_iters && st == SLSGP_OK; ++it)
        {
            SweepJob job;
            job.acq_type = (int) acq_type, job.ucb_beta = ucb_beta, job.M = K, job.d_Xq = dp(ctx->mx_X);
            job.val = dp(ctx->mx_val), job.grad = dp(ctx->mx_grad);
            st = run_sweep(ctx, job);
            if (st != SLSGP_OK) break;
            ascent_update_kernel<<<(K + 3) / 4, 128, 0, ctx->stream>>>(K, D, dp(ctx->mx_val), dp(ctx->mx_grad), dp(ctx->mx_X),
                                                                          dp(ctx->mx_Xbest), dp(ctx->mx_Gbest),
                                                                          ptr<AscentState>(ctx->mx_state), 1.6, 0.35, 0.25);
            ++ctx->launches;
            if (cudaGetLastError() != cudaSuccess) st = fail(ctx, SLSGP_ERR_CUDA, "ascent_update_kernel launch failed");
        }
        ctx->sweep_mode = user_mode;
        if (is_tensor_mode(user_mode)) ctx->Mcap = 0;
        TRY(st);
        ascent_winner_kernel<<<1, 256, 0, ctx->stream>>>(ptr<AscentState>(ctx->mx_state), K, ptr<ArgMax>(ctx->am_acc));
        LAUNCH_CHECK();
        TRY(phase_end(ctx, "maximize"));
        ArgMax win;
        CUDA_TRY(cudaMemcpyAsync(&win, ctx->am_acc.p, sizeof(ArgMax), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (win.i < 0) return fail(ctx, SLSGP_ERR_NAN, "slsgp_acq_maximize: every start evaluated to NaN");
        if (x_best_out)
            CUDA_TRY(cudaMemcpyAsync(x_best_out, dp(ctx->mx_Xbest) + (size_t) win.i * D, sizeof(double) * D, cudaMemcpyDeviceToHost, ctx->stream));
        if (grad_best_out)
            CUDA_TRY(cudaMemcpyAsync(grad_best_out, dp(ctx->mx_Gbest) + (size_t) win.i * D, sizeof(double) * D, cudaMemcpyDeviceToHost, ctx->stream));
        if (val_sweep_best_out)
        {
            // the best value the global stage alone found (diagnostic: what the local stage added)
            std::vector<ArgMax> slices((size_t) K);
            CUDA_TRY(cudaMemcpyAsync(slices.data(), ctx->mx_best.p, sizeof(ArgMax) * (size_t) K, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            double b = -INFINITY;
            for (const auto& a : slices)
                if (a.i >= 0 && a.v > b) b = a.v;
            *val_sweep_best_out = b;
        }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (val_best_out) *val_best_out = win.v;
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // L1 arrays on request (l1.cuh). theta / x go through the scratch buffer `comb`, so a fitted model is left untouched.
    // ---------------------------------------------------------------------------------------------------------
    static slsgp_status l1_upload(slsgp_ctx* ctx, const double* theta, const double* x, double** d_theta, double** d_inv_l, double** d_x)
    {
        const int D = ctx->D;
        for (int i = 0; i <= D; ++i)
            if (!std::isfinite(theta[i])) return fail(ctx, SLSGP_ERR_NAN, "non-finite kernel hyper-parameter");
        TRY(ensure(ctx, ctx->comb, sizeof(double) * (size_t) (3 * D + 1)));
        double* h = ctx->pinned;
        for (int i = 0; i <= D; ++i) h[i] = theta[i];
        for (int i = 0; i < D; ++i) h[D + 1 + i] = 1.0 / theta[1 + i];
        for (int i = 0; i < D; ++i) h[2 * D + 1 + i] = x ? x[i] : 0.0;
        CUDA_TRY(cudaMemcpyAsync(ctx->comb.p, h, sizeof(double) * (size_t) (3 * D + 1), cudaMemcpyHostToDevice, ctx->stream));
        *d_theta = dp(ctx->comb), *d_inv_l = dp(ctx->comb) + D + 1, *d_x = dp(ctx->comb) + 2 * D + 1;
        return SLSGP_OK;
    }

    slsgp_status slsgp_small_k(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* theta, const double* x, double* k_out,
                               double* dk_dx_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_data) return fail(ctx, SLSGP_ERR_STATE, "slsgp_small_k before slsgp_set_data");
        if (!theta || !x || (kernel_type != 0 && kernel_type != 1)) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_small_k: bad arguments");
        const int N = ctx->N, D = ctx->D;
        for (int d = 0; d < D; ++d)
            if (!std::isfinite(x[d])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_small_k: non-finite x");
        CUDA_TRY(cudaSetDevice(ctx->device));
        double *d_theta, *d_inv_l, *d_x;
        TRY(l1_upload(ctx, theta, x, &d_theta, &d_inv_l, &d_x));
        DevBuf out; // k (N) followed by dk/dx (D x N)
        TRY(ensure(ctx, out, sizeof(double) * (size_t) (D + 1) * N));
        double *d_k = dp(out), *d_dk = dp(out) + N;
        small_k_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->X), N, D, d_x, d_theta, d_inv_l, (int) kernel_type,
                                                                 (ctx->compat & SLSGP_COMPAT_SE_XGRAD_2X) ? 2.0 : 1.0, k_out ? d_k : nullptr,
                                                                 dk_dx_out ? d_dk : nullptr);
        ++ctx->launches;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess && k_out) e = cudaMemcpyAsync(k_out, d_k, sizeof(double) * (size_t) N, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && dk_dx_out) e = cudaMemcpyAsync(dk_dx_out, d_dk, sizeof(double) * (size_t) N * D, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(out.p);
        if (e != cudaSuccess) return fail(ctx, SLSGP_ERR_CUDA, std::string("slsgp_small_k: ") + cudaGetErrorString(e));
        return SLSGP_OK;
    }

    slsgp_status slsgp_gram_theta_derivative(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* theta, double* out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_data) return fail(ctx, SLSGP_ERR_STATE, "slsgp_gram_theta_derivative before slsgp_set_data");
        if (!theta || !out || (kernel_type != 0 && kernel_type != 1)) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_gram_theta_derivative: bad arguments");
        const int N = ctx->N, D = ctx->D;
        CUDA_TRY(cudaSetDevice(ctx->device));
        double *d_theta, *d_inv_l, *d_x;
        TRY(l1_upload(ctx, theta, nullptr, &d_theta, &d_inv_l, &d_x));
        DevBuf       planes;
        const size_t bytes = sizeof(double) * (size_t) (D + 1) * N * N;
        TRY(ensure(ctx, planes, bytes));
        gram_theta_derivative_kernel<<<dim3((N + 15) / 16, (N + 15) / 16), dim3(16, 16), 0, ctx->stream>>>(dp(ctx->X), N, D, d_theta, d_inv_l,
                                                                                                          (int) kernel_type, dp(planes));
        ++ctx->launches;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, planes.p, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        cudaFree(planes.p);
        if (e != cudaSuccess) return fail(ctx, SLSGP_ERR_CUDA, std::string("slsgp_gram_theta_derivative: ") + cudaGetErrorString(e));
        return SLSGP_OK;
    }

    // ---------------------------------------------------------------------------------------------------------
    // K5 / K6
    // ---------------------------------------------------------------------------------------------------------
    slsgp_status slsgp_set_preferences(slsgp_ctx* ctx, const uint32_t* offsets, const uint32_t* idx, int P)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_data) return fail(ctx, SLSGP_ERR_STATE, "slsgp_set_preferences before slsgp_set_data");
        if (P < 0 || (P > 0 && (!offsets || !idx))) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_set_preferences: bad arguments");
        CUDA_TRY(cudaSetDevice(ctx->device));
        const int total = P > 0 ? (int) offsets[P] : 0;
        for (int t = 0; t < P; ++t)
            if (offsets[t + 1] < offsets[t] + 2) return fail(ctx, SLSGP_ERR_INVALID, "preference tuple with fewer than 2 members");
        std::vector<uint32_t> slot_off((size_t) ctx->N + 1, 0), slot_list((size_t) std::max(total, 1));
        for (int s = 0; s < total; ++s)
        {
            if (idx[s] >= (uint32_t) ctx->N) return fail(ctx, SLSGP_ERR_INVALID, "preference index out of range");
            ++slot_off[idx[s] + 1];
        }
        for (int i = 0; i < ctx->N; ++i) slot_off[(size_t) i + 1] += slot_off[(size_t) i];
        std::vector<uint32_t> cursor(slot_off.begin(), slot_off.end() - 1);
        for (int s = 0; s < total; ++s) slot_list[cursor[idx[s]]++] = (uint32_t) s;
        TRY(ensure(ctx, ctx->pref_off, sizeof(uint32_t) * ((size_t) P + 1)));
        TRY(ensure(ctx, ctx->pref_idx, sizeof(uint32_t) * (size_t) std::max(total, 1)));
        TRY(ensure(ctx, ctx->slot_off, sizeof(uint32_t) * ((size_t) ctx->N + 1)));
        TRY(ensure(ctx, ctx->slot_list, sizeof(uint32_t) * (size_t) std::max(total, 1)));
        TRY(ensure(ctx, ctx->loglik, sizeof(double) * (size_t) std::max(P, 1)));
        TRY(ensure(ctx, ctx->contrib, sizeof(double) * (size_t) std::max(total, 1)));
        if (P > 0)
        {
            CUDA_TRY(cudaMemcpyAsync(ctx->pref_off.p, offsets, sizeof(uint32_t) * ((size_t) P + 1), cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(ctx->pref_idx.p, idx, sizeof(uint32_t) * (size_t) total, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(ctx->slot_list.p, slot_list.data(), sizeof(uint32_t) * (size_t) total, cudaMemcpyHostToDevice, ctx->stream));
        }
        CUDA_TRY(cudaMemcpyAsync(ctx->slot_off.p, slot_off.data(), sizeof(uint32_t) * ((size_t) ctx->N + 1), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->P = P, ctx->pref_total = total;
        return SLSGP_OK;
    }

    // GetLogOfLogNormalDist{,Derivative} (mathtoolbox probability-distributions.cpp:47-58): O(D) scalar prior terms,
    // evaluated on the host from the caller's x.
    static double log_lognormal(double x, double mu, double s2)
    {
        const double lx = std::log(x), r = lx - mu;
        return -lx - 0.5 * std::log(2.0 * kPi * s2) - 0.5 * (r * r) / s2;
    }
    static double log_lognormal_derivative(double x, double mu, double s2) { return (mu - std::log(x) - s2) / (x * s2); }

    // Shared GP part of both objectives, for y already in ctx->y and the model (K, L, [Kinv]) current:
    //   alpha = Kinv y; returns -1/2 y.alpha - 1/2 logdet - N/2 log(2 pi); if want_hyper also the data-fit part of the
    //   gradient wrt (a, b, l_1..l_D) into g_hyper (D + 2 values, reference ordering a, b, r).
    // host part of the GP term: value and hyper-gradient from the device scalars (y.alpha, alpha.alpha, tr Kinv) and g_l
    static void gp_term_host(const slsgp_ctx* ctx, double logdet, const double* sc, const double* gl, bool want_hyper, double* value,
                             double* g_hyper)
    {
        const int    N = ctx->N, D = ctx->D;
        const double y_alpha = sc[0], alpha_alpha = sc[1], tr_kinv = sc[2];
        *value = -0.5 * y_alpha + -0.5 * logdet + -0.5 * N * std::log(2.0 * kPi);
        if (want_hyper)
        {
            const double a = ctx->theta_host[0], b = ctx->noise;
            // dK/da = K_f / a and K_f = K_y - b I:  1/2 alpha^T K_f alpha - 1/2 tr(Kinv K_f), all over a
            g_hyper[0] = 0.5 / a * ((y_alpha - b * alpha_alpha) - (N - b * tr_kinv));
            g_hyper[1] = 0.5 * alpha_alpha - 0.5 * tr_kinv; // dK/db = I
            for (int t = 0; t < D; ++t) g_hyper[2 + t] = gl[(size_t) t];
        }
    }

    // Results of one MAP objective evaluation on the general path, in the second half of the pinned staging area, so that the
    // whole evaluation needs ONE stream synchronisation (device-to-host copies into pageable memory synchronise by themselves):
    //   [0..2] y.alpha, alpha.alpha, tr K^-1   [3] logdet   [4] info (int)   [5] BTL log-likelihood   [8 .. 8 + D) d/dl   [256 .. 256 + N) d/dy
    constexpr int kMapOutOffset = 8192, kMapOutGl = 8, kMapOutGy = 256, kMapOutMaxD = 248, kMapOutMaxN = 8192 - 256;

    // slices of the contraction Y = Wm XT1 (nt = ld / 64 tiles along it): as many as divide nt, at most 8
    static int map_y_splits(int nt)
    {
        for (int s = 8; s > 1; --s)
            if (nt % s == 0) return s;
        return 1;
    }

    static slsgp_status gp_term_kernels(slsgp_ctx* ctx, bool want_hyper)
    {
        const int ld = ctx->ld, N = ctx->N, D = ctx->D;
        TRY(do_alpha(ctx));
        gp_scalars_kernel<<<1, 256, 0, ctx->stream>>>(dp(ctx->y), dp(ctx->alpha), dp(ctx->Kinv), N, ld, dp(ctx->scalars));
        LAUNCH_CHECK();
        if (want_hyper)
        {
            const int nt = ld / TILE;
            gram_tile_kernel<1><<<nt * (nt + 1) / 2, 256, 0, ctx->stream>>>(
                dp(ctx->X), N, D, ld, dp(ctx->theta), dp(ctx->inv_l), ctx->noise, ctx->kernel_type, dp(ctx->T),
                dp(ctx->Kinv), dp(ctx->alpha));
            LAUNCH_CHECK();
            // Y = Wm XT1 has ld / 64 x 1 output tiles only: the contraction is split into `splits` slices (batch index), each
            // CTA writes its own partial Y and lengthscale_grad_kernel adds them up in a fixed order
            const int splits = map_y_splits(nt), kc = ld / splits;
            GemmArgs  g      = gemm_args(dp(ctx->T), dp(ctx->XT1), dp(ctx->Ymat), ld, ctx->ldx, kc, ld, ld, ld, 1.0, 0.0);
            g.sA = (long long) kc * ld, g.sB = kc, g.sC = (long long) ld * ctx->ldx;
            TRY((launch_gemm<false, false>(ctx, g, splits)));
            lengthscale_grad_kernel<<<D, 256, 0, ctx->stream>>>(dp(ctx->XT1), dp(ctx->Ymat), N, ld, D, dp(ctx->theta), dp(ctx->g_l), splits,
                                                                (size_t) ld * ctx->ldx);
            LAUNCH_CHECK();
        }
        return SLSGP_OK;
    }

    // All results of one evaluation gathered by ONE small kernel that writes the pinned result block directly (pinned host memory
    // is device-addressable): no device-to-host copy operations at all, where the multi-copy form paid ~5 us for each of six.
    __global__ void __launch_bounds__(256)
        map_pack_kernel(const double* __restrict__ scalars, const int* __restrict__ info, const double* __restrict__ g_l,
                        const double* __restrict__ grad_y, int D, int N, int want_hyper, int want_gy, double* __restrict__ out)
    {
        const int i = blockIdx.x * 256 + threadIdx.x;
        if (i < 3) out[i] = scalars[i];                               // y.alpha, alpha.alpha, tr K^-1 (gp_scalars_kernel)
        if (i == 3) out[3] = scalars[8];                              // log-determinant (logdet_kernel)
        if (i == 4) *reinterpret_cast<int*>(out + 4) = *info;         // pivot check of the factorisation
        if (i == 5) out[5] = scalars[4];                              // BTL log-likelihood (sum_kernel)
        if (want_hyper && i < D) out[kMapOutGl + i] = g_l[i];
        if (want_gy && i < N) out[kMapOutGy + i] = grad_y[i];
    }

    // Enqueue form of gp_term: kernels only; map_results_enqueue (after the BTL kernels) sends everything home.
    static slsgp_status gp_term_enqueue(slsgp_ctx* ctx, bool want_hyper) { return gp_term_kernels(ctx, want_hyper); }

    static slsgp_status map_results_enqueue(slsgp_ctx* ctx, bool want_hyper, bool want_gy)
    {
        double* out_dev = nullptr;
        CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&out_dev), ctx->pinned + kMapOutOffset, 0));
        const int n = std::max(std::max(want_gy ? ctx->N : 0, want_hyper ? ctx->D : 0), 8);
        map_pack_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dp(ctx->scalars), ptr<int>(ctx->info), dp(ctx->g_l), dp(ctx->grad_y), ctx->D, ctx->N,
                                                                  want_hyper ? 1 : 0, want_gy ? 1 : 0, out_dev);
        LAUNCH_CHECK();
        return SLSGP_OK;
    }

    // After the synchronisation: the deferred pivot check of do_factor(defer_check), then the host part of the GP term.
    static slsgp_status gp_term_collect(slsgp_ctx* ctx, double* logdet, bool want_hyper, double* value, double* g_hyper)
    {
        const double* out = ctx->pinned + kMapOutOffset;
        if (ctx->factor_pending)
        {
            ctx->factor_pending = false;
            const int info = *reinterpret_cast<const int*>(out + 4);
            if (info != 0)
            {
                ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = false;
                return fail(ctx, SLSGP_ERR_NOT_SPD,
                            "Cholesky: non-positive pivot at index " + std::to_string(info - 1) + " (K_y is not SPD)");
            }
            ctx->logdet_host = *logdet = out[3];
        }
        gp_term_host(ctx, *logdet, out, out + kMapOutGl, want_hyper, value, g_hyper);
        return SLSGP_OK;
    }

    static slsgp_status gp_term(slsgp_ctx* ctx, double logdet, bool want_hyper, double* value, double* g_hyper)
    {
        const int D = ctx->D;
        TRY(gp_term_kernels(ctx, want_hyper));
        double sc[3];
        CUDA_TRY(cudaMemcpyAsync(sc, ctx->scalars.p, sizeof(sc), cudaMemcpyDeviceToHost, ctx->stream));
        std::vector<double> gl((size_t) D);
        if (want_hyper) CUDA_TRY(cudaMemcpyAsync(gl.data(), ctx->g_l.p, sizeof(double) * (size_t) D, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        gp_term_host(ctx, logdet, sc, gl.data(), want_hyper, value, g_hyper);
        return SLSGP_OK;
    }

    // Small models (small.cuh): hyper-parameters and y up, ONE launch, scalars down. Leaves the context exactly as
    // slsgp_gram + slsgp_factor + slsgp_inverse + slsgp_solve_alpha would (K, L, W, Kinv, alpha, f_best, flags).
    static bool small_model_applies(const slsgp_ctx* ctx)
    {
        static const bool enabled = !(std::getenv("SLSGP_SMALL_FUSED") && std::atoi(std::getenv("SLSGP_SMALL_FUSED")) == 0);
        return enabled && ctx->has_data && ctx->ld == SMALL_N && ctx->D <= SMALL_DMAX;
    }

    constexpr int kSmallOutOffset = 1024; // doubles into the pinned staging area: results of the small-model kernel

    // Enqueue: ONE host-to-device copy of [theta | 1 / l | y], the kernel, ONE device-to-host copy of its results.
    static slsgp_status small_model_launch(slsgp_ctx* ctx, int kernel_type, const double* theta, double noise, const double* y_host, bool want_hyper,
                                           int btl_P = -1, double btl_scale = 1.0, bool btl_grad = false)
    {
        const int N = ctx->N, D = ctx->D;
        if (kernel_type != 0 && kernel_type != 1) return fail(ctx, SLSGP_ERR_INVALID, "unknown kernel_type");
        if (!std::isfinite(noise)) return fail(ctx, SLSGP_ERR_NAN, "non-finite noise level");
        for (int i = 0; i <= D; ++i)
            if (!std::isfinite(theta[i])) return fail(ctx, SLSGP_ERR_NAN, "non-finite kernel hyper-parameter");
        ctx->theta_host.assign(theta, theta + D + 1);
        ctx->kernel_type = kernel_type, ctx->noise = noise;
        ctx->has_gram = ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = false;
        ctx->tc_ready = false;
        double* h = ctx->pinned; // consumed before the API call returns (every caller ends in a stream synchronisation)
        for (int i = 0; i <= D; ++i) h[i] = theta[i];
        for (int i = 0; i < D; ++i) h[D + 1 + i] = 1.0 / theta[1 + i];
        for (int i = 0; i < N; ++i) h[2 * D + 1 + i] = y_host[i];
        double *in_dev = dp(ctx->Ymat), *out_dev = dp(ctx->T);
        CUDA_TRY(cudaMemcpyAsync(in_dev, h, sizeof(double) * (size_t) (2 * D + 1 + N), cudaMemcpyHostToDevice, ctx->stream));
        static bool attr[64] = {};
        if (!attr[ctx->device & 63])
        {
            CUDA_TRY(cudaFuncSetAttribute(small_model_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM_BYTES));
            attr[ctx->device & 63] = true;
        }
        SmallModelArgs a;
        a.X = dp(ctx->X), a.in = in_dev;
        a.N = N, a.D = D, a.kernel_type = kernel_type, a.want_hyper = want_hyper ? 1 : 0, a.noise = noise;
        a.theta = dp(ctx->theta), a.inv_l = dp(ctx->inv_l), a.y = dp(ctx->y);
        a.K = dp(ctx->K), a.L = dp(ctx->L), a.W = dp(ctx->W), a.Kinv = dp(ctx->Kinv), a.alpha = dp(ctx->alpha), a.Kalpha = dp(ctx->Kalpha);
        a.out = out_dev, a.fbest = dp(ctx->fbest), a.fbest_idx = ptr<int>(ctx->fbest_idx), a.info = ptr<int>(ctx->info);
        a.btl_P = btl_P, a.btl_grad = btl_grad ? 1 : 0, a.btl_scale = btl_scale;
        a.pref_off = ptr<uint32_t>(ctx->pref_off), a.pref_idx = ptr<uint32_t>(ctx->pref_idx), a.slot_off = ptr<uint32_t>(ctx->slot_off);
        a.slot_list = ptr<uint32_t>(ctx->slot_list), a.contrib = dp(ctx->contrib);
        {
            ProfScope ps(ctx, "small_model");
            small_model_kernel<<<1, 256, SMALL_SMEM_BYTES, ctx->stream>>>(a);
            LAUNCH_CHECK();
        }
        CUDA_TRY(cudaMemcpyAsync(ctx->pinned + kSmallOutOffset, out_dev, sizeof(double) * (size_t) (5 + D + (btl_P >= 0 ? 1 + N : 0)), cudaMemcpyDeviceToHost,
                                 ctx->stream));
        return SLSGP_OK;
    }

    // After the stream has been synchronised: status, flags, value and hyper-gradient of the GP term.
    static slsgp_status small_model_collect(slsgp_ctx* ctx, bool want_hyper, double* logdet_out, double* value, double* g_hyper)
    {
        const double* r = ctx->pinned + kSmallOutOffset;
        if (r[4] != 0.0)
            return fail(ctx, SLSGP_ERR_NOT_SPD,
                        "Cholesky: non-positive pivot at index " + std::to_string((int) r[4] - 1) + " (K_y is not SPD)");
        ctx->logdet_host = r[3];
        ctx->has_gram = ctx->has_factor = ctx->has_W = ctx->has_inverse = ctx->has_alpha = true;
        ++ctx->model_version;
        if (logdet_out) *logdet_out = r[3];
        gp_term_host(ctx, r[3], r, r + 5, want_hyper, value, g_hyper);
        return SLSGP_OK;
    }

    slsgp_status slsgp_map_objective_pref(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* x, int n_x,
                                          int use_map, double default_a, double default_r, double default_b,
                                          double prior_var, double btl_scale, double* f_out, double* grad_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_data) return fail(ctx, SLSGP_ERR_STATE, "slsgp_map_objective_pref before slsgp_set_data");
        const int N = ctx->N, D = ctx->D;
        if (!x || !f_out || n_x != (use_map ? N + 2 + D : N))
            return fail(ctx, SLSGP_ERR_INVALID, "slsgp_map_objective_pref: x/f_out null or n_x != N (+ 2 + D)");
        for (int i = 0; i < n_x; ++i)
            if (!std::isfinite(x[i])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_map_objective_pref: non-finite x");
        CUDA_TRY(cudaSetDevice(ctx->device));
        TRY(phase_begin(ctx, "map"));
        double              logdet = 0.0, gp = 0.0;
        std::vector<double> gh((size_t) D + 2);
        const bool          want_hyper = grad_out && use_map, small = use_map && small_model_applies(ctx);
        // general path: every result travels through the pinned block and the evaluation synchronises once (SLSGP_MAP_ONE_SYNC=0: A/B)
        static const bool one_sync_enabled = !(std::getenv("SLSGP_MAP_ONE_SYNC") && std::atoi(std::getenv("SLSGP_MAP_ONE_SYNC")) == 0);
        const bool        one_sync = one_sync_enabled && !small && D <= kMapOutMaxD && N <= kMapOutMaxN;
        if (ctx->factor_pending) ctx->factor_pending = ctx->has_factor = false; // left over from a call that failed half-way
        // SLSGP_COMPAT_NOISELESS: the reference's SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION build (:48-52, 139, 178-192, 237):
        // K = K_f (b fixed at 0 whatever x holds), no prior on b, d/db = 0
        const bool   noiseless = use_map && (ctx->compat & SLSGP_COMPAT_NOISELESS);
        const double b_used    = noiseless ? 0.0 : (use_map ? x[N + 1] : 0.0);
        if (use_map)
        {
            std::vector<double> theta((size_t) D + 1);
            theta[0] = x[N + 0];
            for (int i = 0; i < D; ++i) theta[(size_t) 1 + i] = x[N + 2 + i];
            if (small) // ONE launch and two copies: model, alpha, GP scalars, length-scale gradient AND the BTL terms (small.cuh)
                TRY(small_model_launch(ctx, (int) kernel_type, theta.data(), b_used, x, want_hyper, ctx->P, btl_scale, grad_out != nullptr));
            else
            {
                TRY(do_gram(ctx, (int) kernel_type, theta.data(), b_used));
                TRY(do_factor(ctx, &logdet, one_sync));
                CUDA_TRY(cudaMemcpyAsync(ctx->y.p, x, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
                if (one_sync)
                    TRY(gp_term_enqueue(ctx, want_hyper));
                else
                    TRY(gp_term(ctx, logdet, want_hyper, &gp, gh.data()));
            }
        }
        else
        {
            if (!ctx->has_factor)
                return fail(ctx, SLSGP_ERR_STATE, "use_map_hyperparams == 0 needs slsgp_gram + slsgp_factor first");
            logdet = ctx->logdet_host;
            CUDA_TRY(cudaMemcpyAsync(ctx->y.p, x, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
            if (one_sync)
                TRY(gp_term_enqueue(ctx, want_hyper));
            else
                TRY(gp_term(ctx, logdet, want_hyper, &gp, gh.data()));
        }
        double* const map_out = ctx->pinned + kMapOutOffset;

        // BTL likelihood of the tuples (the small-model launch above already holds them)
        double loglik = 0.0;
        if (ctx->P > 0 && !small)
        {
            btl_tuple_kernel<<<(ctx->P + 127) / 128, 128, 0, ctx->stream>>>(
                dp(ctx->y), ptr<uint32_t>(ctx->pref_off), ptr<uint32_t>(ctx->pref_idx), ctx->P, btl_scale,
                dp(ctx->loglik), dp(ctx->contrib), grad_out ? 1 : 0);
            LAUNCH_CHECK();
            sum_kernel<<<1, 256, 0, ctx->stream>>>(dp(ctx->loglik), ctx->P, dp(ctx->scalars) + 4);
            LAUNCH_CHECK();
            if (!one_sync) CUDA_TRY(cudaMemcpyAsync(&loglik, dp(ctx->scalars) + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (grad_out && !small)
        {
            btl_gather_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(
                dp(ctx->contrib), ptr<uint32_t>(ctx->slot_off), ptr<uint32_t>(ctx->slot_list), dp(ctx->alpha), N,
                dp(ctx->grad_y));
            LAUNCH_CHECK();
            if (!one_sync) CUDA_TRY(cudaMemcpyAsync(grad_out, ctx->grad_y.p, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (one_sync) TRY(map_results_enqueue(ctx, want_hyper, grad_out != nullptr));
        TRY(phase_end(ctx, "map"));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (one_sync)
        {
            TRY(gp_term_collect(ctx, &logdet, want_hyper, &gp, gh.data()));
            if (ctx->P > 0) loglik = map_out[5];
            if (grad_out) std::memcpy(grad_out, map_out + kMapOutGy, sizeof(double) * (size_t) N);
        }
        if (small)
        {
            TRY(small_model_collect(ctx, want_hyper, &logdet, &gp, gh.data()));
            const double* r = ctx->pinned + kSmallOutOffset + 5 + D;
            loglik          = r[0];
            if (grad_out) std::memcpy(grad_out, r + 1, sizeof(double) * (size_t) N);
        }

        double obj = loglik + gp;
        if (use_map) // log-normal hyper-priors centred on the defaults (:175-192) and their derivatives (:103-112, :69-73)
        {
            const double a = x[N], b = x[N + 1];
            obj += log_lognormal(a, std::log(default_a), prior_var);
            if (!noiseless) obj += log_lognormal(b, std::log(default_b), prior_var);
            for (int i = 0; i < D; ++i) obj += log_lognormal(x[N + 2 + i], std::log(default_r), prior_var);
            if (grad_out)
            {
                grad_out[N + 0] = gh[0] + log_lognormal_derivative(a, std::log(default_a), prior_var);
                grad_out[N + 1] = noiseless ? 0.0 : gh[1] + log_lognormal_derivative(b, std::log(default_b), prior_var);
                for (int i = 0; i < D; ++i)
                    grad_out[N + 2 + i] = gh[(size_t) 2 + i] + log_lognormal_derivative(x[N + 2 + i], std::log(default_r), prior_var);
            }
        }
        *f_out = obj;
        return SLSGP_OK;
    }

    slsgp_status slsgp_map_objective_pref_whitened(slsgp_ctx* ctx, const double* z, double btl_scale, double* f_out,
                                                   double* grad_z_out, double* y_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_factor) return fail(ctx, SLSGP_ERR_STATE, "slsgp_map_objective_pref_whitened needs slsgp_gram + slsgp_factor first");
        if (!z || !f_out) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_map_objective_pref_whitened: z / f_out null");
        const int N = ctx->N, ld = ctx->ld;
        double    zz = 0.0;
        for (int i = 0; i < N; ++i)
        {
            if (!std::isfinite(z[i])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_map_objective_pref_whitened: non-finite z");
            zz += z[i] * z[i];
        }
        CUDA_TRY(cudaSetDevice(ctx->device));
        if (N <= 1024 && (size_t) (2 * N + 1) * sizeof(double) <= ctx->pinned_bytes)
        {
            // small problems: one single-block launch, operands and results through the pinned staging area
            TRY(ensure(ctx, ctx->comb, sizeof(double) * (size_t) (2 * N + 1)));
            TRY(phase_begin(ctx, "map"));
            std::memcpy(ctx->pinned, z, sizeof(double) * N);
            CUDA_TRY(cudaMemcpyAsync(ctx->vec.p, ctx->pinned, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
            const size_t smem = sizeof(double) * (size_t) (3 * N + 32);
            map_whitened_fused_kernel<<<1, 1024, smem, ctx->stream>>>(dp(ctx->L), N, ld, dp(ctx->vec), ptr<uint32_t>(ctx->pref_off),
                                                                      ptr<uint32_t>(ctx->pref_idx), ctx->P, btl_scale,
                                                                      ptr<uint32_t>(ctx->slot_off), ptr<uint32_t>(ctx->slot_list),
                                                                      dp(ctx->contrib), dp(ctx->y), grad_z_out ? 1 : 0, dp(ctx->comb));
            LAUNCH_CHECK();
            ctx->has_alpha = false; // y changed under the cached alpha
            TRY(phase_end(ctx, "map"));
            CUDA_TRY(cudaMemcpyAsync(ctx->pinned, ctx->comb.p, sizeof(double) * (size_t) (2 * N + 1), cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            const double* res = ctx->pinned;
            if (grad_z_out) std::memcpy(grad_z_out, res + 1, sizeof(double) * N);
            if (y_out) std::memcpy(y_out, res + 1 + N, sizeof(double) * N);
            *f_out = res[0] + -0.5 * zz + -0.5 * ctx->logdet_host + -0.5 * N * std::log(2.0 * kPi);
            return SLSGP_OK;
        }
        TRY(phase_begin(ctx, "map"));
        const int blocks = (ld * 32 + 255) / 256;
        CUDA_TRY(cudaMemsetAsync(ctx->vec.p, 0, sizeof(double) * ld, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->vec.p, z, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
        gemv_kernel<false><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->L), ld, ld, dp(ctx->vec), dp(ctx->y), 1); // y = L z
        LAUNCH_CHECK();
        ctx->has_alpha = false; // y changed under the cached alpha
        double loglik = 0.0;
        if (ctx->P > 0)
        {
            btl_tuple_kernel<<<(ctx->P + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->y), ptr<uint32_t>(ctx->pref_off), ptr<uint32_t>(ctx->pref_idx),
                                                                          ctx->P, btl_scale, dp(ctx->loglik), dp(ctx->contrib), grad_z_out ? 1 : 0);
            LAUNCH_CHECK();
            sum_kernel<<<1, 256, 0, ctx->stream>>>(dp(ctx->loglik), ctx->P, dp(ctx->scalars) + 4);
            LAUNCH_CHECK();
            CUDA_TRY(cudaMemcpyAsync(&loglik, dp(ctx->scalars) + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (grad_z_out)
        {
            CUDA_TRY(cudaMemsetAsync(ctx->grad_y.p, 0, sizeof(double) * ld, ctx->stream));
            if (ctx->P > 0)
            {
                btl_gather_kernel<<<(N + 127) / 128, 128, 0, ctx->stream>>>(dp(ctx->contrib), ptr<uint32_t>(ctx->slot_off),
                                                                           ptr<uint32_t>(ctx->slot_list), nullptr, N, dp(ctx->grad_y));
                LAUNCH_CHECK();
            }
            gemv_kernel<true><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->L), ld, ld, dp(ctx->grad_y), dp(ctx->Kalpha), 1); // L^T g
            LAUNCH_CHECK();
            CUDA_TRY(cudaMemcpyAsync(grad_z_out, ctx->Kalpha.p, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (y_out) CUDA_TRY(cudaMemcpyAsync(y_out, ctx->y.p, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(phase_end(ctx, "map"));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (grad_z_out)
            for (int i = 0; i < N; ++i) grad_z_out[i] -= z[i];
        *f_out = loglik + -0.5 * zz + -0.5 * ctx->logdet_host + -0.5 * N * std::log(2.0 * kPi);
        return SLSGP_OK;
    }

    slsgp_status slsgp_whiten(slsgp_ctx* ctx, const double* y, double* z_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_factor) return fail(ctx, SLSGP_ERR_STATE, "slsgp_whiten needs slsgp_gram + slsgp_factor first");
        if (!y || !z_out) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_whiten: null argument");
        CUDA_TRY(cudaSetDevice(ctx->device));
        TRY(do_trtri(ctx)); // W = L^-1
        const int N = ctx->N, ld = ctx->ld, blocks = (ld * 32 + 255) / 256;
        CUDA_TRY(cudaMemsetAsync(ctx->vec.p, 0, sizeof(double) * ld, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(ctx->vec.p, y, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
        gemv_kernel<false><<<blocks, 256, 0, ctx->stream>>>(dp(ctx->W), ld, ld, dp(ctx->vec), dp(ctx->Kalpha), 1);
        LAUNCH_CHECK();
        CUDA_TRY(cudaMemcpyAsync(z_out, ctx->Kalpha.p, sizeof(double) * N, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return SLSGP_OK;
    }

    slsgp_status slsgp_map_objective_gpr(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* y,
                                         const double* x, double* f_out, double* grad_out)
    {
        if (!ctx) return SLSGP_ERR_INVALID;
        if (!ctx->has_data) return fail(ctx, SLSGP_ERR_STATE, "slsgp_map_objective_gpr before slsgp_set_data");
        if (!y || !x || !f_out) return fail(ctx, SLSGP_ERR_INVALID, "slsgp_map_objective_gpr: null argument");
        const int N = ctx->N, D = ctx->D;
        for (int i = 0; i < D + 2; ++i)
            if (!std::isfinite(x[i])) return fail(ctx, SLSGP_ERR_NAN, "slsgp_map_objective_gpr: non-finite x");
        CUDA_TRY(cudaSetDevice(ctx->device));
        TRY(phase_begin(ctx, "map"));
        std::vector<double> theta((size_t) D + 1);
        theta[0] = x[0];
        for (int i = 0; i < D; ++i) theta[(size_t) 1 + i] = x[2 + i];
        double              logdet = 0.0, gp = 0.0;
        std::vector<double> gh((size_t) D + 2);
        if (small_model_applies(ctx))
        {
            TRY(small_model_launch(ctx, (int) kernel_type, theta.data(), x[1], y, grad_out != nullptr));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            TRY(small_model_collect(ctx, grad_out != nullptr, &logdet, &gp, gh.data()));
        }
        else
        {
            static const bool one_sync_enabled = !(std::getenv("SLSGP_MAP_ONE_SYNC") && std::atoi(std::getenv("SLSGP_MAP_ONE_SYNC")) == 0);
            const bool        one_sync = one_sync_enabled && D <= kMapOutMaxD; // as in slsgp_map_objective_pref
            if (ctx->factor_pending) ctx->factor_pending = ctx->has_factor = false;
            TRY(do_gram(ctx, (int) kernel_type, theta.data(), x[1]));
            TRY(do_factor(ctx, &logdet, one_sync));
            CUDA_TRY(cudaMemcpyAsync(ctx->y.p, y, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
            if (one_sync)
            {
                TRY(gp_term_enqueue(ctx, grad_out != nullptr));
                TRY(map_results_enqueue(ctx, grad_out != nullptr, false));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                TRY(gp_term_collect(ctx, &logdet, grad_out != nullptr, &gp, gh.data()));
            }
            else
                TRY(gp_term(ctx, logdet, grad_out != nullptr, &gp, gh.data()));
        }
        TRY(phase_end(ctx, "map"));
        // fixed log-normal priors of src/gaussian-process-regressor.cpp:18-24
        const double a_mu = std::log(0.500), a_s2 = 0.50, b_mu = std::log(1e-04), b_s2 = 0.50, r_mu = std::log(0.500), r_s2 = 0.50;
        double       reg  = log_lognormal(x[0], a_mu, a_s2) + log_lognormal(x[1], b_mu, b_s2);
        for (int i = 0; i < D; ++i) reg += log_lognormal(x[2 + i], r_mu, r_s2);
        *f_out = gp + reg;
        if (grad_out)
        {
            grad_out[0] = gh[0] + log_lognormal_derivative(x[0], a_mu, a_s2);
            grad_out[1] = gh[1] + log_lognormal_derivative(x[1], b_mu, b_s2);
            for (int i = 0; i < D; ++i) grad_out[2 + i] = gh[(size_t) 2 + i] + log_lognormal_derivative(x[2 + i], r_mu, r_s2);
        }
        return SLSGP_OK;
    }
}
