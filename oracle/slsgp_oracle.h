/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
 *
 * Plain-C (C11, libm only) CPU restatement of the Gaussian-process hot path of
 * yuki-koyama/sequential-line-search, written function by function after the reference sources (each function
 * in slsgp_oracle.c cites the reference file:line it follows). It exists so that tests can check the CUDA
 * library (include/slsgp.h) against an independent implementation on a box that has no /root/reference.
 *
 * Who may use it: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as the
 * CHECKER only. The product (libslsgp.so and the host C++ above it) never links or calls this file.
 *
 * Pinning: the reference ships no golden vectors or known-answer tests for this path.
 * The restatement is therefore pinned against the reference ITSELF: oracle/_ref/libsls_ref_probe.so is the
 * reference's unmodified sources compiled in this container (oracle/Makefile), tests/test_oracle_vs_ref.py
 * compares the two function by function, and tests/golden/ holds vectors generated from the reference build
 * (tests/golden/make_golden.py) for boxes where the reference tree is absent.
 *
 * Conventions: all matrices column-major (Eigen's default), X is D x N (one point per column),
 * theta = (a, l_1..l_D) is the kernel hyper-parameter vector, b is the noise level, kernel_type 0 = ARD squared
 * exponential, 1 = ARD Matern-5/2 (order of KernelType, include/sequential-line-search/kernel-type.hpp:8-12),
 * acq_type 0 = Expected Improvement, 1 = GP-UCB (acquisition-function.hpp:11-15).
 */
#ifndef SLSGP_ORACLE_H
#define SLSGP_ORACLE_H

#ifdef __cplusplus
extern "C"
{
#endif

    /* ---- kernels (external/mathtoolbox/src/kernel-functions.cpp) ---- */
    double slsgp_oracle_kernel(int kernel_type, int D, const double* xa, const double* xb, const double* theta);
    void   slsgp_oracle_kernel_theta_derivative(int kernel_type, int D, const double* xa, const double* xb,
                                                const double* theta, double* out /* D+1 */);
    void   slsgp_oracle_kernel_first_arg_derivative(int kernel_type, int D, const double* xa, const double* xb,
                                                    const double* theta, double* out /* D */);

    /* ---- kernel matrices (src/regressor.cpp) ---- */
    void slsgp_oracle_large_kf(int kernel_type, int D, int N, const double* X, const double* theta, double* K);
    void slsgp_oracle_large_ky(int kernel_type, int D, int N, const double* X, const double* theta, double b,
                               double* K);
    void slsgp_oracle_small_k(int kernel_type, int D, int N, const double* X, const double* theta,
                              const double* x, double* k /* N */);
    void slsgp_oracle_small_k_x_derivative(int kernel_type, int D, int N, const double* X, const double* theta,
                                           const double* x, double* J /* D x N */);
    void slsgp_oracle_large_ky_theta_derivative(int kernel_type, int D, int N, const double* X,
                                                const double* theta, double* out /* (D+1) x N x N */);

    /* ---- dense linear algebra standing in for Eigen::LLT / inverse() ---- */
    /* returns 0 on success, else 1 + index of the first non-positive pivot. L is lower, upper part zeroed. */
    int    slsgp_oracle_cholesky(int N, const double* K, double* L);
    void   slsgp_oracle_llt_solve(int N, const double* L, int nrhs, double* B /* N x nrhs, in/out */);
    double slsgp_oracle_logdet(int N, const double* L);
    int    slsgp_oracle_inverse(int N, const double* K, double* Kinv); /* via LLT; 0 on success */

    /* ---- regressor state shared by PreferenceRegressor and GaussianProcessRegressor predictions ---- */
    typedef struct
    {
        int           kernel_type, D, N;
        const double* X;     /* D x N */
        const double* theta; /* D+1 */
        double        b;
        const double* y; /* N: MAP goodness values (preference) or observed values (GPR) */
        const double* L; /* N x N Cholesky factor of K_y */
    } slsgp_oracle_model;

    double slsgp_oracle_predict_mu(const slsgp_oracle_model* m, const double* x);
    double slsgp_oracle_predict_sigma(const slsgp_oracle_model* m, const double* x);
    void   slsgp_oracle_predict_mu_derivative(const slsgp_oracle_model* m, const double* x, double* out);
    void   slsgp_oracle_predict_sigma_derivative(const slsgp_oracle_model* m, const double* x, double* out);
    /* index of argmax_i mu(X_i) (src/regressor.cpp:29-43); *f_best_out = mu at that point. O(N^3) as written. */
    int slsgp_oracle_predict_maximum_point_from_data(const slsgp_oracle_model* m, double* f_best_out);

    /* ---- acquisition (external/mathtoolbox/src/acquisition-functions.cpp, src/acquisition-function.cpp) ---- */
    /* f_best is passed in (the reference recomputes it on every call; same number). */
    double slsgp_oracle_acq_value(const slsgp_oracle_model* m, int acq_type, double ucb_beta, double f_best,
                                  const double* x);
    void   slsgp_oracle_acq_derivative(const slsgp_oracle_model* m, int acq_type, double ucb_beta, double f_best,
                                       const double* x, double* out);
    /* Batched convenience used by tests and by the "port" CPU baseline: caches alpha = K^-1 y (mathematically
     * what every PredictMu recomputes) and loops the functions above over M query points. Any out may be NULL. */
    void slsgp_oracle_acq_batch(const slsgp_oracle_model* m, int acq_type, double ucb_beta, double f_best,
                                long long M, const double* Xq /* D x M */, double* mu, double* sigma,
                                double* dmu /* D x M */, double* dsigma /* D x M */, double* val,
                                double* grad /* D x M */);

    /* ---- BTL likelihood (include/sequential-line-search/utils.hpp:25-52) ---- */
    double slsgp_oracle_btl(int n, const double* f, double scale);
    void   slsgp_oracle_btl_derivative(int n, const double* f, double scale, double* out);

    /* ---- log-normal prior (external/mathtoolbox/src/probability-distributions.cpp:47-58) ---- */
    double slsgp_oracle_log_lognormal(double x, double mu, double sigma2);
    double slsgp_oracle_log_lognormal_derivative(double x, double mu, double sigma2);

    /* ---- MAP objectives ---- */
    /* PreferenceRegressor objective (src/preference-regressor.cpp:129-259).
     * x = [y (N)] or [y (N), a, b, r_1..r_D] when use_map_hyperparams; tuples in CSR form.
     * grad (same length as x) may be NULL. Returns the objective (to be maximised). */
    double slsgp_oracle_map_objective_pref(int kernel_type, int D, int N, const double* X, int P,
                                           const unsigned* offsets, const unsigned* idx,
                                           int use_map_hyperparams, double default_a, double default_r,
                                           double default_b, double prior_var, double btl_scale,
                                           const double* x, double* grad);
    /* the same objective as the reference's SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION build evaluates it */
    double slsgp_oracle_map_objective_pref_noiseless(int kernel_type, int D, int N, const double* X, int P,
                                           const unsigned* offsets, const unsigned* idx,
                                           int use_map_hyperparams, double default_a, double default_r,
                                           double default_b, double prior_var, double btl_scale,
                                           const double* x, double* grad);
    /* GaussianProcessRegressor objective (src/gaussian-process-regressor.cpp:141-193), x = (a, b, r_1..r_D). */
    double slsgp_oracle_map_objective_gpr(int kernel_type, int D, int N, const double* X, const double* y,
                                          const double* x, double* grad);

#ifdef __cplusplus
}
#endif
#endif /* SLSGP_ORACLE_H */
