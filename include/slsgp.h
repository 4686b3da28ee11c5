/* slsgp.h — C ABI of libslsgp, the B200 (sm_100a) implementation of the Gaussian-process hot path of
 * yuki-koyama/sequential-line-search.
 *
 * The reference has no plugin / FFI seam for this path; its replaceable seams are C++-level
 * (the `Regressor` virtuals, the L1 free functions of regressor.hpp, `acquisition_func::*`, the NLopt objective
 * callbacks). Each entry point below names the reference code it stands in for (paths relative to the upstream
 * repository). The host-side C++ that keeps the reference's class API calls these functions and nothing else;
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; all matrices column-major (Eigen's default); X is D x N, one observation per column;
 *     theta = (a, l_1 .. l_D) is the reference's `kernel_hyperparams` vector, `noise` its `noise_level` b.
 *   - every function returns an `slsgp_status`; nothing throws across the boundary; `slsgp_last_error(ctx)` gives
 *     the message for the last non-zero status on that context.
 *   - pointers named `*_out` may be NULL when the caller does not want that result copied back to the host.
 *   - a context owns all device memory it uses and one CUDA stream; calls on one context are serialised by the
 *     caller; distinct contexts may be used from distinct threads.
 *   - there is NO CPU fallback: if no CUDA device is usable `slsgp_ctx_create` fails with SLSGP_ERR_CUDA.
 */
#ifndef SLSGP_H
#define SLSGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

#define SLSGP_VERSION 100

    typedef struct slsgp_ctx slsgp_ctx;

    typedef enum
    {
        SLSGP_OK          = 0,
        SLSGP_ERR_INVALID = 1, /* bad argument (null pointer, non-positive size, unknown enum value) */
        SLSGP_ERR_STATE   = 2, /* call order: e.g. slsgp_factor before slsgp_gram */
        SLSGP_ERR_NOT_SPD = 3, /* Cholesky met a non-positive pivot; slsgp_last_error names the index */
        SLSGP_ERR_NAN     = 4, /* NaN / Inf in an input the reference would have asserted on */
        SLSGP_ERR_CUDA    = 5, /* CUDA runtime error, or no usable device */
        SLSGP_ERR_NOMEM   = 6
    } slsgp_status;

    /* sequential_line_search::KernelType (include/sequential-line-search/kernel-type.hpp:8-12), same order */
    typedef enum
    {
        SLSGP_KERNEL_ARD_SQUARED_EXP = 0,
        SLSGP_KERNEL_ARD_MATERN52    = 1
    } slsgp_kernel_type;

    /* sequential_line_search::AcquisitionFuncType (include/sequential-line-search/acquisition-function.hpp:11-15) */
    typedef enum
    {
        SLSGP_ACQ_EXPECTED_IMPROVEMENT = 0,
        SLSGP_ACQ_GP_UCB               = 1
    } slsgp_acq_type;

    /* Arithmetic used by the candidate sweep (slsgp_posterior_batch / slsgp_acq_batch / slsgp_acq_argmax).
     * Gram build, factorisation, inverse, alpha and the MAP objectives are always IEEE double. */
    typedef enum
    {
        SLSGP_SWEEP_FP64      = 0, /* IEEE double throughout; parity 1e-5 relative (north_star "FP64") */
        /* fp16 operands on the tcgen05 tensor pipe, fp32 accumulation in TMEM (both library kernels; D <= 67, 66 for Matern,
         * whose gradient weight travels as a second fp16 operand).
         * K^-1 and k* are each split into two fp16 terms; the contraction is evaluated as
         *   TENSOR     k16.A16 + k16.A_lo + k_lo.A16   fp32-class result, parity 1e-3 on every test distribution
         *   TENSOR_X2  k16.A16 + k16.A_lo              sigma / EI value fp32-class, gradients limited by k16 (~2e-3)
         *   TENSOR_X1  k16.A16                          fastest; 1e-3 only for well-conditioned K (cond <~ 1e3) */
        SLSGP_SWEEP_TENSOR    = 1,
        SLSGP_SWEEP_TENSOR_X2 = 2,
        SLSGP_SWEEP_TENSOR_X1 = 3
    } slsgp_sweep_mode;

    /* Reference quirks that parity has to reproduce; all on by default. */
#define SLSGP_COMPAT_SE_XGRAD_2X 1u /* mathtoolbox kernel-functions.cpp:92 returns -2 k (x_a-x_b)/l^2 */
    /* Off by default: the reference's compile-time option SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION (CMakeLists.txt:28-33,
     * src/preference-regressor.cpp:48-52,139,178-192,237): slsgp_map_objective_pref with use_map_hyperparams builds K = K_f (the
     * noise entry of x is ignored, b = 0), drops the prior on b and returns d/db = 0. */
#define SLSGP_COMPAT_NOISELESS 2u

    /* ---- context -------------------------------------------------------------------------------------------- */
    slsgp_status slsgp_ctx_create(int device, slsgp_ctx** ctx_out);
    /* A context that spans several GPUs of one node (device_ids[0] is the primary). Used exactly like a single-device context:
     * the model is built on the primary (Gram, Cholesky, inverse and the MAP objectives are "replicas only"); the
     * first sweep after a model change copies X, K^-1, alpha, f_best and the hyper-parameters to the other devices peer to peer
     * (NVLink / NVSwitch); slsgp_acq_argmax, slsgp_acq_maximize (candidate ranges >= 2^16) and slsgp_posterior_batch /
     * slsgp_acq_batch (batches >= 2^16) then split their candidates over the devices, one host thread each, and reduce the
     * per-device winners (value, index, point) on the host - highest value, lowest candidate index on ties, so the answer of
     * slsgp_acq_argmax does not depend on the number of devices. slsgp_ctx_destroy releases the whole group. */
    slsgp_status slsgp_ctx_create_multi(const int* device_ids, int n_devices, slsgp_ctx** ctx_out);
    int          slsgp_ctx_device_count(const slsgp_ctx* ctx);
    slsgp_status slsgp_ctx_destroy(slsgp_ctx* ctx);
    const char*  slsgp_last_error(const slsgp_ctx* ctx);
    const char*  slsgp_status_string(slsgp_status s);
    slsgp_status slsgp_set_compat_flags(slsgp_ctx* ctx, unsigned flags);
    slsgp_status slsgp_set_sweep_mode(slsgp_ctx* ctx, slsgp_sweep_mode mode);
    slsgp_status slsgp_get_sweep_mode(const slsgp_ctx* ctx, slsgp_sweep_mode* mode_out);
    /* Two-tier precision of the tensor modes: after a sweep that returns per-candidate arrays, the candidates whose posterior
     * variance came out below tau * a (a = signal variance; these sit next to data points, where sigma^2 = a - k K^-1 k and the
     * sums behind grad sigma cancel and the fp16 / fp32 round-off is amplified by a / sigma^2) are re-evaluated through the
     * IEEE-double sweep and overwritten. tau in [0, 1]; 0 switches the second tier off; default 0.1 (environment variable
     * SLSGP_REFINE_TAU). Costs one host synchronisation per call and nothing else when no candidate qualifies. When more than half
     * of a batch (>= 8192 candidates) qualifies - dense data: thousands of observations in a handful of dimensions, where
     * sigma^2 << a almost everywhere - the context remembers it for that model and sweeps later batches in IEEE double directly
     * (the tensor pass would be wasted work); the next model change resets this. */
    slsgp_status slsgp_set_refine_threshold(slsgp_ctx* ctx, double tau);
    /* Use a caller-owned CUDA stream (a cudaStream_t passed as void*) instead of the context's own; NULL restores
     * the context's stream. Lets a host framework order libslsgp work with its own copies and events. */
    slsgp_status slsgp_set_stream(slsgp_ctx* ctx, void* cuda_stream);
    slsgp_status slsgp_synchronize(slsgp_ctx* ctx);

    /* ---- data ------------------------------------------------------------------------------------------------
     * Replaces the regressors' `m_X` copy (src/preference-regressor.cpp:273, gaussian-process-regressor.cpp:203).
     * Uploads X (D x N, column-major, host memory) and invalidates every derived quantity. */
    slsgp_status slsgp_set_data(slsgp_ctx* ctx, const double* X, int N, int D);

    /* Incremental refit across iterations (the reference rebuilds everything per SubmitFeedbackData,
     * src/sequential-line-search.cpp:91-103). Like slsgp_set_data, but when the context holds a factored model and the first p >= 1
     * columns of X are, bit for bit, its first p data points (same D, N within the padded size of the model, at most 8 columns
     * beyond them: SLSGP_EXTEND_MAX_NEW) the model of those p points is kept - the leading blocks of K_y, L and L^-1 - and grown
     * by the remaining N - p columns with the bordered update of slsgp_append_point (O(N^2) per point) under the hyper-parameters
     * of the last slsgp_gram. (The data manager of the reference appends new points and, when it merges two coincident ones,
     * removes both and appends their midpoint: src/preference-data-manager.cpp:14-141; so p is usually the position of the point
     * that was merged away.) slsgp_gram called again with the same hyper-parameters, slsgp_factor and slsgp_inverse then find
     * their results current and only copy them out (they always do: a context never recomputes a matrix it already holds).
     * y, alpha and the preference tuples are dropped.
     * n_kept_out (may be NULL): p when the model was extended, 0 when it was replaced (plain slsgp_set_data). */
    slsgp_status slsgp_set_data_extend(slsgp_ctx* ctx, const double* X, int N, int D, int* n_kept_out);

    /* Forgets every matrix derived from the data (K_y, its factor, the inverses, alpha) so that the next slsgp_gram / slsgp_factor
     * / slsgp_inverse recompute them even for unchanged hyper-parameters; the data stay. For benchmarks that time the model build
     * repeatedly: a context otherwise never recomputes a matrix it already holds (see slsgp_set_data_extend). */
    slsgp_status slsgp_invalidate(slsgp_ctx* ctx);

    /* Frees the per-shard sweep workspaces of the context when they hold more than keep_bytes (they only ever grow with the
     * largest batch seen); the fitted model stays. For hosts that pool contexts. */
    slsgp_status slsgp_trim(slsgp_ctx* ctx, size_t keep_bytes);

    /* ---- L1 arrays on request --------------------------------------------------------------------------------------
     * The fused sweep and MAP kernels never materialise these; they are offered for callers of the reference's free functions.
     * Both work on the context's X with the GIVEN theta and leave a fitted model untouched.
     *   slsgp_small_k                CalcSmallK (src/regressor.cpp:45-59): k_out[i] = k(x, X_i), N values, and
     *                                CalcSmallKSmallXDerivative (:91-108): dk_dx_out (D x N, column i = d k(x, X_i) / d x);
     *                                either output may be NULL
     *   slsgp_gram_theta_derivative  CalcLargeKYThetaDerivative (:110-134): out = D + 1 matrices of N x N, back to back
     *                                (d K / d a, d K / d l_1, ...); CalcLargeKYNoiseLevelDerivative (:136-141) is the identity. */
    slsgp_status slsgp_small_k(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* theta, const double* x, double* k_out,
                               double* dk_dx_out);
    slsgp_status slsgp_gram_theta_derivative(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* theta, double* out);

    /* ---- K1: Gram matrix --------------------------------------------------------------------------------------
     * K_y = K_f + noise * I: CalcLargeKY / CalcLargeKF (src/regressor.cpp:61-89) with the per-pair kernels
     * GetArdSquaredExpKernel / GetArdMatern52Kernel (external/mathtoolbox/src/kernel-functions.cpp:7-20, 95-112).
     * Keeps (kernel_type, theta, noise) as the model's hyper-parameters. K_out: N x N or NULL. */
    slsgp_status slsgp_gram(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* theta, double noise,
                            double* K_out);

    /* ---- K2: Cholesky ------------------------------------------------------------------------------------------
     * K_y = L L^T: Eigen::LLT at src/preference-regressor.cpp:162,290,370; logdet = 2 sum log L_ii:
     * mathtoolbox CalcLogDetOfSymmetricPositiveDefiniteMatrix (src/log-determinant.cpp:8-11).
     * L_out: N x N lower triangle (upper part zero) or NULL. Returns SLSGP_ERR_NOT_SPD on a bad pivot (the
     * reference never checks LLT::info()).
     * From N = 6144 on the factorisation runs in its two-level form on two streams of the context's own (panels on a
     * high-priority stream, rank-k trailing updates on a second one); both are ordered against the context's stream by
     * events on entry and exit, so callers see ordinary stream semantics (slsgp_set_stream included). */
    slsgp_status slsgp_factor(slsgp_ctx* ctx, double* logdet_out, double* L_out);

    /* ---- K3: factor application --------------------------------------------------------------------------------
     * Explicit K_y^-1: GaussianProcessRegressor::m_K_y_inv = m_K_y.inverse() (src/gaussian-process-regressor.cpp:
     * 159,211,231) and LLT::solve(Identity) (src/preference-regressor.cpp:66). Computed from the Cholesky factor
     * (L^-1, then L^-T L^-1). Needed by the sweep and by the MAP hyper-parameter gradient; computed on demand if
     * the caller did not ask for it. Kinv_out: N x N or NULL. */
    slsgp_status slsgp_inverse(slsgp_ctx* ctx, double* Kinv_out);

    /* alpha = K_y^-1 y: the LLT::solve(m_y) that every PredictMu / PredictMuDerivative repeats
     * (src/preference-regressor.cpp:296,320; `m_K_y_inv * m_y`, src/gaussian-process-regressor.cpp:238,262).
     * Also caches f_best = max_i mu(X_i), i.e. mu(PredictMaximumPointFromData()) (src/regressor.cpp:29-43), which
     * Expected Improvement needs. y: N goodness / observed values. */
    slsgp_status slsgp_solve_alpha(slsgp_ctx* ctx, const double* y, double* alpha_out);
    slsgp_status slsgp_get_f_best(slsgp_ctx* ctx, double* f_best_out, int* index_out);

    /* Append one data point (x[D], y_new) to a factored model in O(N^2): bordered update of K_y, L, L^-1, K_y^-1 and of
     * the log-determinant, then alpha and f_best again if slsgp_solve_alpha had been called. Replaces the rebuild +
     * `.inverse()` of the temporary GaussianProcessRegressor that FindNextPoints grows by one pending point per option
     * (src/acquisition-function.cpp:281-296). Hyper-parameters, kernel and noise are those of the last slsgp_gram. The
     * preference tuples are dropped. K_col_out: the new last column of K_y (N + 1 values) or NULL; Kinv_out: the new
     * (N + 1) x (N + 1) inverse or NULL. SLSGP_ERR_NOT_SPD (model unchanged) when the Schur complement is not positive. */
    slsgp_status slsgp_append_point(slsgp_ctx* ctx, const double* x, double y_new, double* K_col_out, double* Kinv_out);

    /* ---- K4: batched posterior and acquisition sweep --------------------------------------------------------------
     * For each of M query points (Xq: D x M, column-major, host memory):
     *   mu      Regressor::PredictMu               (src/preference-regressor.cpp:293-297)
     *   sigma   Regressor::PredictSigma            (:299-313; clamps sigma^2 < 0 to 0)
     *   dmu     Regressor::PredictMuDerivative     (:315-321)            D x M
     *   dsigma  Regressor::PredictSigmaDerivative  (:323-330; divides by sigma unguarded, as the reference) D x M
     * through CalcSmallK / CalcSmallKSmallXDerivative (src/regressor.cpp:45-59, 91-108). */
    slsgp_status slsgp_posterior_batch(slsgp_ctx* ctx, const double* Xq, int64_t M, double* mu_out,
                                       double* sigma_out, double* dmu_out, double* dsigma_out);

    /*   val   acquisition_func::CalcAcquisitionValue            (src/acquisition-function.cpp:170-198)
     *   grad  acquisition_func::CalcAcquisitionValueDerivative  (:200-230)                                D x M
     * with mathtoolbox GetExpectedImprovement{,Derivative} / GetGaussianProcessUpperConfidenceBound{,Derivative}
     * (external/mathtoolbox/src/acquisition-functions.cpp:8-78). Returns 0 / zero vectors when the model holds no
     * data (:176-179, :206-209). */
    slsgp_status slsgp_acq_batch(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, const double* Xq,
                                 int64_t M, double* val_out, double* grad_out);

    /* The acquisition formulas alone, for callers that combine the posterior mean of one model with the deviation of
     * another: Schonlau's batch criterion, objective_for_multiple_points (src/acquisition-function.cpp:63-110), where mu
     * comes from the original regressor and sigma from a regressor that already holds the pending points. All arrays are
     * host memory, laid out as slsgp_posterior_batch writes them (dmu / dsigma / grad: D x M); f_best is
     * mu(PredictMaximumPointFromData()) of the original model (slsgp_get_f_best). grad_out (and then dmu / dsigma) may be
     * NULL. Evaluated on `ctx`'s device; the context needs no model. */
    slsgp_status slsgp_acq_from_posterior(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, double f_best, int D,
                                          int64_t M, const double* mu, const double* sigma, const double* dmu,
                                          const double* dsigma, double* val_out, double* grad_out);

    /* The global stage of FindNextPoints (src/acquisition-function.cpp:264-276 over objective_for_multiple_points :63-110) as one
     * device-resident arg-max: candidate i of [first, first + count) from the generator of slsgp_candidates, mu from `ctx`, sigma
     * from `ctx_sigma` (same device, same D), EI / UCB with f_best of `ctx`, lowest index wins ties; only the winner is copied out. */
    slsgp_status slsgp_pair_acq_argmax(slsgp_ctx* ctx, slsgp_ctx* ctx_sigma, slsgp_acq_type acq_type, double ucb_beta, uint64_t seed,
                                       int64_t first, int64_t count, double* x_best_out, double* val_best_out,
                                       int64_t* index_best_out);

    /* Same sweep with every buffer already resident on the context's device (device pointers): no copies.
     * Any output may be NULL. Asynchronous on the context's stream. */
    slsgp_status slsgp_acq_batch_device(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta,
                                        const double* d_Xq, int64_t M, double* d_mu, double* d_sigma,
                                        double* d_dmu, double* d_dsigma, double* d_val, double* d_grad);

    /* Global search stage of FindGlobalSolution (src/acquisition-function.cpp:112-167: DIRECT, or random-start
     * multi-start) as a dense sweep: candidate i in [first, first + count) is the point of [0,1]^D produced by the
     * counter-based generator slsgp_candidates(seed, i) — independent of how a range is split across GPUs. Evaluates
     * the acquisition value (+ gradient when grad_best_out != NULL) for all of them in `count`-sized shards on this
     * device and returns the arg-max (lowest index wins ties). x_best_out: D. */
    slsgp_status slsgp_acq_argmax(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, uint64_t seed,
                                  int64_t first, int64_t count, double* x_best_out, double* val_best_out,
                                  int64_t* index_best_out, double* grad_best_out);
    /* FindGlobalSolution (src/acquisition-function.cpp:112-167) as one device-resident procedure; the multi-start form of
     * the reference (random starts + NLopt L-BFGS on threads, :125-144) without a host round trip per objective evaluation:
     *   1. the dense sweep of slsgp_acq_argmax over candidates [first, first + count) in the context's sweep mode;
     *   2. the best candidate of each of `n_starts` equal slices of that range becomes a starting point (<= 16384);
     *   3. `n_iters` iterations of projected, normalised-gradient ascent with a per-start trust radius inside [0,1]^D,
     *      every start advanced by ONE batched IEEE-double sweep per iteration;
     *   4. the arg-max over the refined starts.
     * x_best_out: D. grad_best_out (D) and val_sweep_best_out (best value of step 1 alone) may be NULL.
     * Shards across GPUs like slsgp_acq_argmax: each rank maximises its own range, the winners are compared by value. */
    slsgp_status slsgp_acq_maximize(slsgp_ctx* ctx, slsgp_acq_type acq_type, double ucb_beta, uint64_t seed, int64_t first,
                                    int64_t count, int n_starts, int n_iters, double* x_best_out, double* val_best_out,
                                    double* grad_best_out, double* val_sweep_best_out);

    /* Arg-max (lowest index wins ties, NaN never wins) of `count` values already on the device, e.g. the d_val of
     * slsgp_acq_batch_device; index_best_out = index0 + position. Synchronises the context's stream. */
    slsgp_status slsgp_argmax_device(slsgp_ctx* ctx, const double* d_val, int64_t count, int64_t index0,
                                     double* val_best_out, int64_t* index_best_out);
    /* The generator itself, for hosts that want the same candidates (D x count, host memory). */
    slsgp_status slsgp_candidates(slsgp_ctx* ctx, uint64_t seed, int64_t first, int64_t count, double* Xq_out);

    /* ---- K5/K6: MAP objectives ---------------------------------------------------------------------------------------
     * Preference tuples in CSR form: tuple t = idx[offsets[t] .. offsets[t+1]), first member preferred
     * (include/sequential-line-search/preference.hpp:9-19). */
    slsgp_status slsgp_set_preferences(slsgp_ctx* ctx, const uint32_t* offsets, const uint32_t* idx, int P);

    /* `objective` of PreferenceRegressor (src/preference-regressor.cpp:129-259): BTL log-likelihood
     * (utils.hpp:25-52) + log N(y; 0, K_y) + log-normal hyper-priors, and its gradient
     * (CalcObjectiveThetaDerivative :77-115, CalcObjectiveNoiseLevelDerivative :53-74).
     * x = [y_1..y_N] or, when use_map_hyperparams, [y_1..y_N, a, b, r_1..r_D]; grad_out has the same length or is
     * NULL (derivative-free caller, `grad.empty()` in NLopt). With use_map_hyperparams == 0 the kernel matrix is the
     * one last built by slsgp_gram + slsgp_factor (the reference reuses m_K / m_K_llt, :160-162). */
    slsgp_status slsgp_map_objective_pref(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* x, int n_x,
                                          int use_map_hyperparams, double default_a, double default_r,
                                          double default_b, double prior_var, double btl_scale, double* f_out,
                                          double* grad_out);

    /* The same objective for FIXED hyper-parameters in whitened coordinates: y = L z with K_y = L L^T the factor last built
     * by slsgp_gram + slsgp_factor, so that log N(y; 0, K_y) = -1/2 |z|^2 - 1/2 logdet - N/2 log(2 pi) and
     *   F(z) = sum_tuples log BTL(L z) - 1/2 |z|^2 + const,     grad_z F = L^T grad_y(log BTL) - z.
     * F(z) equals slsgp_map_objective_pref(y = L z, use_map_hyperparams = 0). The prior Hessian is the identity in z, which is
     * what makes a quasi-Newton driver converge in tens instead of hundreds of evaluations; no K^-1 is needed.
     * z: N; grad_z_out and y_out (= L z) may be NULL. slsgp_whiten gives z = L^-1 y for a starting point. */
    slsgp_status slsgp_map_objective_pref_whitened(slsgp_ctx* ctx, const double* z, double btl_scale, double* f_out,
                                                   double* grad_z_out, double* y_out);
    slsgp_status slsgp_whiten(slsgp_ctx* ctx, const double* y, double* z_out);

    /* `objective` of GaussianProcessRegressor (src/gaussian-process-regressor.cpp:141-193, calc_grad :108-127, priors
     * :18-64): log marginal likelihood + fixed log-normal priors; x = (a, b, r_1..r_D); grad_out D+2 or NULL. */
    slsgp_status slsgp_map_objective_gpr(slsgp_ctx* ctx, slsgp_kernel_type kernel_type, const double* y,
                                         const double* x, double* f_out, double* grad_out);

    /* ---- introspection (bench / tests) ----------------------------------------------------------------------------- */
    /* Number of kernels this context has launched since creation (bench.py's `gpu_launches`). */
    uint64_t     slsgp_launch_count(const slsgp_ctx* ctx);
    /* Milliseconds (CUDA events on the context's stream) spent in the most recent call of the named phase:
     * "gram", "factor", "inverse", "alpha", "sweep", "map", "maximize", "append". Returns < 0 for an unknown name. */
    double       slsgp_last_phase_ms(const slsgp_ctx* ctx, const char* phase);
    /* Per-kernel device timing: while enabled, every launch of the named hot kernels is bracketed by CUDA events on
     * the context's stream. slsgp_profile_read synchronises, returns the summed duration and launch count of
     * `kernel` ("sweep_gemm", "sweep_kstar", "sweep_reduce", "sweep_grad_gemm", "sweep_finish", "tc_kstar",
     * "tc_gemm", "gram", "chol_step", "small_model") since the last read, and clears that kernel's records. */
    slsgp_status slsgp_profile_enable(slsgp_ctx* ctx, int on);
    slsgp_status slsgp_profile_read(slsgp_ctx* ctx, const char* kernel, double* total_ms_out, uint64_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* SLSGP_H */
