/* TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT. See slsgp_oracle.h for scope, pinning and who may call this.
 *
 * Plain-C restatement of the reference's GP hot path. Reference paths are relative to the upstream repository
 * (yuki-koyama/sequential-line-search @ cfdc4f14, mathtoolbox @ 2fd2302e). The dense linear algebra that the
 * reference delegates to Eigen (LLT, LLT::solve, inverse) — Eigen is a system dependency, absent from the
 * reference tree — is restated here as textbook column Cholesky + substitution.
 */
#include "slsgp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.14159265358979323846 /* mathtoolbox/constants.hpp */

/* r^2 = diff^T diag(1 / l^2) diff   (kernel-functions.cpp:16-17, 104-105) */
static double scaled_sq_dist(int D, const double* xa, const double* xb, const double* theta)
{
    double r2 = 0.0;
    for (int i = 0; i < D; ++i)
    {
        const double d   = xa[i] - xb[i];
        const double inv = 1.0 / (theta[1 + i] * theta[1 + i]);
        r2 += (d * inv) * d;
    }
    return r2;
}

/* GetArdSquaredExpKernel (kernel-functions.cpp:7-20); GetArdMatern52Kernel (:95-112) */
double slsgp_oracle_kernel(int kt, int D, const double* xa, const double* xb, const double* theta)
{
    const double a  = theta[0];
    const double r2 = scaled_sq_dist(D, xa, xb, theta);
    if (kt == 0) return a * exp(-0.5 * r2);
    const double s = sqrt(5.0 * r2);
    return a * (1.0 + s + (5.0 / 3.0) * r2) * exp(-s);
}

/* GetArdSquaredExpKernelThetaDerivative (:22-50); GetArdMatern52KernelThetaDerivative (:114-142) */
void slsgp_oracle_kernel_theta_derivative(int kt, int D, const double* xa, const double* xb, const double* theta,
                                          double* out)
{
    const double a  = theta[0];
    const double r2 = scaled_sq_dist(D, xa, xb, theta);
    if (kt == 0)
    {
        out[0]         = exp(-0.5 * r2);
        const double k = a * exp(-0.5 * r2);
        for (int i = 0; i < D; ++i)
        {
            const double r = xa[i] - xb[i], l = theta[1 + i];
            out[1 + i] = k * (r * r) / (l * l * l);
        }
        return;
    }
    const double s = sqrt(5.0 * r2), scale = 1.0 + s + (5.0 / 3.0) * r2, e = exp(-s);
    out[0] = scale * e;
    for (int i = 0; i < D; ++i)
    {
        const double d = xa[i] - xb[i], l = theta[1 + i];
        out[1 + i] = (5.0 / 3.0) * a * e * (1.0 + s) * (d * d) * (1.0 / (l * l * l));
    }
}

/* GetArdSquaredExpKernelFirstArgDerivative (:81-93) — NOTE the factor -2.0 at :92 (twice the analytic
 * derivative); parity with the reference means reproducing it.
 * GetArdMatern52KernelFirstArgDerivative (:179-212), including the zero-vector guard at :198. */
void slsgp_oracle_kernel_first_arg_derivative(int kt, int D, const double* xa, const double* xb,
                                              const double* theta, double* out)
{
    const double a  = theta[0];
    const double r2 = scaled_sq_dist(D, xa, xb, theta);
    if (kt == 0)
    {
        const double k = a * exp(-0.5 * r2);
        for (int i = 0; i < D; ++i)
        {
            const double inv = 1.0 / (theta[1 + i] * theta[1 + i]);
            out[i]           = (-2.0 * k) * inv * (xa[i] - xb[i]);
        }
        return;
    }
    const double s = sqrt(5.0 * r2), scale = 1.0 + s + (5.0 / 3.0) * r2, e = exp(-s);
    if (s < 1e-30)
    {
        for (int i = 0; i < D; ++i) out[i] = 0.0;
        return;
    }
    for (int i = 0; i < D; ++i)
    {
        const double inv  = 1.0 / (theta[1 + i] * theta[1 + i]);
        const double dr2  = 2.0 * inv * (xa[i] - xb[i]);          /* d r^2 / d x_a          (:203-204) */
        const double ds   = 0.5 * sqrt(5.0 / r2) * dr2;           /* d sqrt(5 r^2) / d x_a  (:205-206) */
        const double de   = -ds * e;                              /* (:207) */
        const double dsc  = ds + (5.0 / 3.0) * dr2;               /* (:208-209) */
        out[i]            = a * (dsc * e + scale * de);           /* (:211) */
    }
}

/* CalcLargeKF (src/regressor.cpp:72-89): upper-triangle loop, mirrored. */
void slsgp_oracle_large_kf(int kt, int D, int N, const double* X, const double* theta, double* K)
{
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j)
        {
            const double v = slsgp_oracle_kernel(kt, D, X + (size_t) i * D, X + (size_t) j * D, theta);
            K[(size_t) i + (size_t) j * N] = v;
            K[(size_t) j + (size_t) i * N] = v;
        }
}

/* CalcLargeKY (src/regressor.cpp:61-70): K_f + b I */
void slsgp_oracle_large_ky(int kt, int D, int N, const double* X, const double* theta, double b, double* K)
{
    slsgp_oracle_large_kf(kt, D, N, X, theta, K);
    for (int i = 0; i < N; ++i) K[(size_t) i * N + i] += b * 1.0;
}

/* CalcSmallK (src/regressor.cpp:45-59) */
void slsgp_oracle_small_k(int kt, int D, int N, const double* X, const double* theta, const double* x, double* k)
{
    for (int i = 0; i < N; ++i) k[i] = slsgp_oracle_kernel(kt, D, x, X + (size_t) i * D, theta);
}

/* CalcSmallKSmallXDerivative (src/regressor.cpp:91-108): column i = d k(x, X_i) / d x */
void slsgp_oracle_small_k_x_derivative(int kt, int D, int N, const double* X, const double* theta,
                                       const double* x, double* J)
{
    for (int i = 0; i < N; ++i)
        slsgp_oracle_kernel_first_arg_derivative(kt, D, x, X + (size_t) i * D, theta, J + (size_t) i * D);
}

/* CalcLargeKYThetaDerivative (src/regressor.cpp:110-134) */
void slsgp_oracle_large_ky_theta_derivative(int kt, int D, int N, const double* X, const double* theta,
                                            double* out)
{
    double*      g  = (double*) malloc(sizeof(double) * (size_t) (D + 1));
    const size_t NN = (size_t) N * N;
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j)
        {
            slsgp_oracle_kernel_theta_derivative(kt, D, X + (size_t) i * D, X + (size_t) j * D, theta, g);
            for (int k = 0; k <= D; ++k)
            {
                out[k * NN + (size_t) i + (size_t) j * N] = g[k];
                out[k * NN + (size_t) j + (size_t) i * N] = g[k];
            }
        }
    free(g);
}

/* Eigen::LLT<MatrixXd>(K) as used at src/preference-regressor.cpp:162,290,370 */
int slsgp_oracle_cholesky(int N, const double* K, double* L)
{
    int fail = 0;
    memset(L, 0, sizeof(double) * (size_t) N * N);
    for (int j = 0; j < N; ++j)
    {
        double d = K[(size_t) j + (size_t) j * N];
        for (int k = 0; k < j; ++k) d -= L[(size_t) j + (size_t) k * N] * L[(size_t) j + (size_t) k * N];
        if (!(d > 0.0) && !fail) fail = 1 + j;
        const double ljj              = sqrt(d);
        L[(size_t) j + (size_t) j * N] = ljj;
        for (int i = j + 1; i < N; ++i)
        {
            double s = K[(size_t) i + (size_t) j * N];
            for (int k = 0; k < j; ++k) s -= L[(size_t) i + (size_t) k * N] * L[(size_t) j + (size_t) k * N];
            L[(size_t) i + (size_t) j * N] = s / ljj;
        }
    }
    return fail;
}

/* LLT::solve (src/preference-regressor.cpp:165,296,309,320,329; matrix right-hand sides at :66,:97) */
void slsgp_oracle_llt_solve(int N, const double* L, int nrhs, double* B)
{
    for (int c = 0; c < nrhs; ++c)
    {
        double* x = B + (size_t) c * N;
        for (int j = 0; j < N; ++j)
        {
            const double* lj = L + (size_t) j * N;
            x[j] /= lj[j];
            const double xj = x[j];
            for (int i = j + 1; i < N; ++i) x[i] -= lj[i] * xj;
        }
        for (int j = N - 1; j >= 0; --j)
        {
            const double* lj = L + (size_t) j * N;
            double        s  = x[j];
            for (int i = j + 1; i < N; ++i) s -= lj[i] * x[i];
            x[j] = s / lj[j];
        }
    }
}

/* CalcLogDetOfSymmetricPositiveDefiniteMatrix (mathtoolbox/src/log-determinant.cpp:8-11): 2 sum log L_ii */
double slsgp_oracle_logdet(int N, const double* L)
{
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += log(L[(size_t) i + (size_t) i * N]);
    return 2.0 * s;
}

/* K_y.inverse() (src/gaussian-process-regressor.cpp:159,211,231). The reference's inverse() is Eigen's
 * partial-pivot LU; for the SPD K_y the result is the same matrix up to rounding, so the restatement goes
 * through the Cholesky factor. */
int slsgp_oracle_inverse(int N, const double* K, double* Kinv)
{
    double*   L    = (double*) malloc(sizeof(double) * (size_t) N * N);
    const int fail = slsgp_oracle_cholesky(N, K, L);
    memset(Kinv, 0, sizeof(double) * (size_t) N * N);
    for (int i = 0; i < N; ++i) Kinv[(size_t) i + (size_t) i * N] = 1.0;
    slsgp_oracle_llt_solve(N, L, N, Kinv);
    free(L);
    return fail;
}

/* ------------------------------------------------------------------------------------------------------------
 * Predictions. PreferenceRegressor::Predict* (src/preference-regressor.cpp:293-330) and
 * GaussianProcessRegressor::Predict* (src/gaussian-process-regressor.cpp:234-272) compute the same quantities
 * (k^T K^-1 y etc.), the former through LLT::solve, the latter through the explicit inverse.
 * ---------------------------------------------------------------------------------------------------------- */
static double dot(int n, const double* a, const double* b)
{
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* PredictMu: k^T LLT.solve(y) (:293-297) */
double slsgp_oracle_predict_mu(const slsgp_oracle_model* m, const double* x)
{
    const int N = m->N;
    double*   k = (double*) malloc(sizeof(double) * 2 * (size_t) N);
    double*   a = k + N;
    slsgp_oracle_small_k(m->kernel_type, m->D, N, m->X, m->theta, x, k);
    memcpy(a, m->y, sizeof(double) * (size_t) N);
    slsgp_oracle_llt_solve(N, m->L, 1, a);
    const double mu = dot(N, k, a);
    free(k);
    return mu;
}

/* PredictSigma: sqrt(max(0, theta_0 - k^T LLT.solve(k))) (:299-313) */
double slsgp_oracle_predict_sigma(const slsgp_oracle_model* m, const double* x)
{
    const int N = m->N;
    double*   k = (double*) malloc(sizeof(double) * 2 * (size_t) N);
    double*   s = k + N;
    slsgp_oracle_small_k(m->kernel_type, m->D, N, m->X, m->theta, x, k);
    memcpy(s, k, sizeof(double) * (size_t) N);
    slsgp_oracle_llt_solve(N, m->L, 1, s);
    const double sigma_2 = m->theta[0] - dot(N, k, s);
    free(k);
    return sigma_2 < 0 ? 0.0 : sqrt(sigma_2);
}

/* PredictMuDerivative: (dk/dx) LLT.solve(y) (:315-321) */
void slsgp_oracle_predict_mu_derivative(const slsgp_oracle_model* m, const double* x, double* out)
{
    const int N = m->N, D = m->D;
    double*   J = (double*) malloc(sizeof(double) * ((size_t) D * N + (size_t) N));
    double*   a = J + (size_t) D * N;
    slsgp_oracle_small_k_x_derivative(m->kernel_type, D, N, m->X, m->theta, x, J);
    memcpy(a, m->y, sizeof(double) * (size_t) N);
    slsgp_oracle_llt_solve(N, m->L, 1, a);
    for (int d = 0; d < D; ++d)
    {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += J[(size_t) d + (size_t) i * D] * a[i];
        out[d] = s;
    }
    free(J);
}

/* PredictSigmaDerivative: -(1/sigma) (dk/dx) LLT.solve(k) (:323-330); no guard on sigma == 0. */
void slsgp_oracle_predict_sigma_derivative(const slsgp_oracle_model* m, const double* x, double* out)
{
    const int N = m->N, D = m->D;
    double*   J = (double*) malloc(sizeof(double) * ((size_t) D * N + (size_t) N));
    double*   s = J + (size_t) D * N;
    slsgp_oracle_small_k_x_derivative(m->kernel_type, D, N, m->X, m->theta, x, J);
    slsgp_oracle_small_k(m->kernel_type, D, N, m->X, m->theta, x, s);
    const double sigma = slsgp_oracle_predict_sigma(m, x);
    slsgp_oracle_llt_solve(N, m->L, 1, s);
    for (int d = 0; d < D; ++d)
    {
        double acc = 0.0;
        for (int i = 0; i < N; ++i) acc += J[(size_t) d + (size_t) i * D] * s[i];
        out[d] = -(1.0 / sigma) * acc;
    }
    free(J);
}

/* Regressor::PredictMaximumPointFromData (src/regressor.cpp:29-43): first index of max_i mu(X_i). */
int slsgp_oracle_predict_maximum_point_from_data(const slsgp_oracle_model* m, double* f_best_out)
{
    int    best = 0;
    double fb   = 0.0;
    for (int i = 0; i < m->N; ++i)
    {
        const double f = slsgp_oracle_predict_mu(m, m->X + (size_t) i * m->D);
        if (i == 0 || f > fb) fb = f, best = i;
    }
    if (f_best_out) *f_best_out = fb;
    return best;
}

/* ------------------------------------------------------------------------------------------------------------
 * Acquisition functions (mathtoolbox/src/acquisition-functions.cpp, probability-distributions.cpp:6-20)
 * ---------------------------------------------------------------------------------------------------------- */
static double std_normal_pdf(double x) { return (1.0 / sqrt(2.0 * PI)) * exp(-0.5 * x * x); }   /* :6-9   */
static double std_normal_pdf_derivative(double x) { return -x * std_normal_pdf(x); }            /* :11-15 */
static double std_normal_cdf(double x) { return 0.5 * (1.0 + erf(x / sqrt(2.0))); }             /* :17-20 */

/* GetExpectedImprovement (:8-24), GetGaussianProcessUpperConfidenceBound (:57-65);
 * dispatch as CalcAcquisitionValue (src/acquisition-function.cpp:170-198). */
static double acq_from_moments(int acq_type, double beta, double f_best, double mu, double sigma)
{
    if (acq_type == 1) return mu + beta * sigma;
    const double diff = mu - f_best;
    const double Z    = diff / sigma;
    const double EI   = diff * std_normal_cdf(Z) + sigma * std_normal_pdf(Z);
    return (sigma < 1e-16 || isnan(EI)) ? 0.0 : EI;
}

/* GetExpectedImprovementDerivative (:26-55), GetGaussianProcessUpperConfidenceBoundDerivative (:67-78) */
static void acq_grad_from_moments(int acq_type, double beta, double f_best, int D, double mu, double sigma,
                                  const double* dmu, const double* dsigma, double* out)
{
    if (acq_type == 1)
    {
        for (int d = 0; d < D; ++d) out[d] = dmu[d] + beta * dsigma[d];
        return;
    }
    const double diff = mu - f_best;
    const double Z    = diff / sigma;
    const double Phi = std_normal_cdf(Z), phi = std_normal_pdf(Z), dphi = std_normal_pdf_derivative(Z);
    int          has_nan = 0;
    for (int d = 0; d < D; ++d)
    {
        const double dZ = (dmu[d] - Z * dsigma[d]) / sigma;
        out[d]          = dmu[d] * Phi + diff * dZ * phi + dsigma[d] * phi + sigma * dZ * dphi;
        if (isnan(out[d])) has_nan = 1;
    }
    if (sigma < 1e-16 || has_nan)
        for (int d = 0; d < D; ++d) out[d] = 0.0;
}

double slsgp_oracle_acq_value(const slsgp_oracle_model* m, int acq_type, double beta, double f_best,
                              const double* x)
{
    if (m->N == 0) return 0.0; /* src/acquisition-function.cpp:176-179 */
    return acq_from_moments(acq_type, beta, f_best, slsgp_oracle_predict_mu(m, x),
                            slsgp_oracle_predict_sigma(m, x));
}

void slsgp_oracle_acq_derivative(const slsgp_oracle_model* m, int acq_type, double beta, double f_best,
                                 const double* x, double* out)
{
    const int D = m->D;
    if (m->N == 0) /* :206-209 */
    {
        for (int d = 0; d < D; ++d) out[d] = 0.0;
        return;
    }
    double* dmu = (double*) malloc(sizeof(double) * 2 * (size_t) D);
    double* dsg = dmu + D;
    slsgp_oracle_predict_mu_derivative(m, x, dmu);
    slsgp_oracle_predict_sigma_derivative(m, x, dsg);
    acq_grad_from_moments(acq_type, beta, f_best, D, slsgp_oracle_predict_mu(m, x),
                          slsgp_oracle_predict_sigma(m, x), dmu, dsg, out);
    free(dmu);
}

void slsgp_oracle_acq_batch(const slsgp_oracle_model* m, int acq_type, double beta, double f_best, long long M,
                            const double* Xq, double* mu, double* sigma, double* dmu, double* dsigma, double* val,
                            double* grad)
{
    const int N = m->N, D = m->D;
    double*   alpha = (double*) malloc(sizeof(double) * ((size_t) 2 * N + (size_t) D * N + 2 * (size_t) D));
    double*   k     = alpha + N;
    double*   J     = k + N;
    double*   gm    = J + (size_t) D * N;
    double*   gs    = gm + D;
    memcpy(alpha, m->y, sizeof(double) * (size_t) N);
    slsgp_oracle_llt_solve(N, m->L, 1, alpha); /* LLT.solve(m_y): identical on every call in the reference */
    for (long long q = 0; q < M; ++q)
    {
        const double* x = Xq + (size_t) q * D;
        slsgp_oracle_small_k(m->kernel_type, D, N, m->X, m->theta, x, k);
        slsgp_oracle_small_k_x_derivative(m->kernel_type, D, N, m->X, m->theta, x, J);
        const double mu_x = dot(N, k, alpha);
        for (int d = 0; d < D; ++d)
        {
            double s = 0.0;
            for (int i = 0; i < N; ++i) s += J[(size_t) d + (size_t) i * D] * alpha[i];
            gm[d] = s;
        }
        double* beta_v = (double*) malloc(sizeof(double) * (size_t) N);
        memcpy(beta_v, k, sizeof(double) * (size_t) N);
        slsgp_oracle_llt_solve(N, m->L, 1, beta_v);
        const double s2 = m->theta[0] - dot(N, k, beta_v);
        const double sg = s2 < 0 ? 0.0 : sqrt(s2);
        for (int d = 0; d < D; ++d)
        {
            double s = 0.0;
            for (int i = 0; i < N; ++i) s += J[(size_t) d + (size_t) i * D] * beta_v[i];
            gs[d] = -(1.0 / sg) * s;
        }
        free(beta_v);
        if (mu) mu[q] = mu_x;
        if (sigma) sigma[q] = sg;
        if (dmu) memcpy(dmu + (size_t) q * D, gm, sizeof(double) * (size_t) D);
        if (dsigma) memcpy(dsigma + (size_t) q * D, gs, sizeof(double) * (size_t) D);
        if (val) val[q] = acq_from_moments(acq_type, beta, f_best, mu_x, sg);
        if (grad) acq_grad_from_moments(acq_type, beta, f_best, D, mu_x, sg, gm, gs, grad + (size_t) q * D);
    }
    free(alpha);
}

/* ------------------------------------------------------------------------------------------------------------
 * BTL likelihood (include/sequential-line-search/utils.hpp:25-52); un-stabilised exp as in the reference.
 * ---------------------------------------------------------------------------------------------------------- */
double slsgp_oracle_btl(int n, const double* f, double scale)
{
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += exp((1.0 / scale) * f[i]);
    return exp((1.0 / scale) * f[0]) / sum;
}

void slsgp_oracle_btl_derivative(int n, const double* f, double scale, double* out)
{
    const double btl = slsgp_oracle_btl(n, f, scale);
    const double tmp = -btl * btl / scale;
    double       sum = 0.0;
    for (int i = 1; i < n; ++i) sum += exp((f[i] - f[0]) / scale);
    out[0] = tmp * (-sum);
    for (int i = 1; i < n; ++i) out[i] = tmp * exp((f[i] - f[0]) / scale);
}

/* GetLogOfLogNormalDist / ...Derivative (mathtoolbox/src/probability-distributions.cpp:47-58) */
double slsgp_oracle_log_lognormal(double x, double mu, double sigma2)
{
    const double lx = log(x), r = lx - mu;
    return -lx - 0.5 * log(2.0 * PI * sigma2) - 0.5 * (r * r) / sigma2;
}
double slsgp_oracle_log_lognormal_derivative(double x, double mu, double sigma2)
{
    return (mu - log(x) - sigma2) / (x * sigma2);
}

/* Shared by both MAP objectives: 1/2 alpha^T dK alpha - 1/2 tr(K^-1 dK) for every kernel hyper-parameter
 * (CalcObjectiveThetaDerivative, src/preference-regressor.cpp:77-115; calc_grad_theta,
 * src/gaussian-process-regressor.cpp:79-106) and for the noise level, where dK = I
 * (CalcObjectiveNoiseLevelDerivative :53-74; calc_grad_b :66-77). */
static void gp_hyper_gradient(int kt, int D, int N, const double* X, const double* theta, const double* Kinv,
                              const double* alpha, double* g_theta /* D+1 */, double* g_b)
{
    const size_t NN     = (size_t) N * N;
    double*      tensor = (double*) malloc(sizeof(double) * NN * (size_t) (D + 1));
    slsgp_oracle_large_ky_theta_derivative(kt, D, N, X, theta, tensor);
    for (int t = 0; t <= D; ++t)
    {
        const double* dK    = tensor + (size_t) t * NN;
        double        quad  = 0.0, trace = 0.0;
        for (int j = 0; j < N; ++j)
        {
            double colsum = 0.0;
            for (int i = 0; i < N; ++i)
            {
                colsum += alpha[i] * dK[(size_t) i + (size_t) j * N];
                trace += Kinv[(size_t) j + (size_t) i * N] * dK[(size_t) i + (size_t) j * N];
            }
            quad += colsum * alpha[j];
        }
        g_theta[t] = 0.5 * quad - 0.5 * trace;
    }
    double tr = 0.0;
    for (int i = 0; i < N; ++i) tr += Kinv[(size_t) i + (size_t) i * N];
    *g_b = 0.5 * dot(N, alpha, alpha) - 0.5 * tr;
    free(tensor);
}

/* objective() of PreferenceRegressor (src/preference-regressor.cpp:129-259). noiseless != 0 restates the build with
 * SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION (:48-52, :141, :180-188, :237): with use_map the kernel matrix is K_f (b = 0
 * whatever x[N + 1] holds), b has no prior and d/db = 0; without use_map the option changes nothing (m_K keeps the default noise). */
static double map_objective_pref(int kt, int D, int N, const double* X, int P, const unsigned* offsets, const unsigned* idx,
                                 int use_map, double default_a, double default_r, double default_b, double prior_var,
                                 double btl_scale, const double* x, double* grad, int noiseless)
{
    const double* y     = x;
    double*       theta = (double*) malloc(sizeof(double) * (size_t) (D + 1));
    theta[0]            = use_map ? x[N + 0] : default_a;                       /* :139 */
    const double b      = use_map ? (noiseless ? 0.0 : x[N + 1]) : default_b;   /* :141-143 */
    for (int i = 0; i < D; ++i) theta[1 + i] = use_map ? x[N + 2 + i] : default_r; /* :145-147 */

    double obj = 0.0;
    double tmp[64];
    /* log likelihood of the preference tuples (:150-154, calc_log_likelihood :118-126) */
    for (int t = 0; t < P; ++t)
    {
        const int n = (int) (offsets[t + 1] - offsets[t]);
        for (int i = 0; i < n; ++i) tmp[i] = y[idx[offsets[t] + i]];
        obj += log(slsgp_oracle_btl(n, tmp, btl_scale));
    }

    /* GP prior on y (:160-170) */
    const size_t NN    = (size_t) N * N;
    double*      K     = (double*) malloc(sizeof(double) * (2 * NN + (size_t) N));
    double*      L     = K + NN;
    double*      alpha = L + NN;
    slsgp_oracle_large_ky(kt, D, N, X, theta, b, K);
    slsgp_oracle_cholesky(N, K, L);
    memcpy(alpha, y, sizeof(double) * (size_t) N);
    slsgp_oracle_llt_solve(N, L, 1, alpha);
    obj += -0.5 * dot(N, y, alpha) + -0.5 * slsgp_oracle_logdet(N, L) + -0.5 * N * log(2.0 * PI);

    /* log-normal hyper-priors centred on the defaults (:175-192) */
    if (use_map)
    {
        obj += slsgp_oracle_log_lognormal(theta[0], log(default_a), prior_var);
        if (!noiseless) obj += slsgp_oracle_log_lognormal(b, log(default_b), prior_var); /* :184-186 */
        for (int i = 0; i < D; ++i) obj += slsgp_oracle_log_lognormal(theta[1 + i], log(default_r), prior_var);
    }

    if (grad)
    {
        for (int i = 0; i < N; ++i) grad[i] = 0.0;
        double dbtl[64];
        for (int t = 0; t < P; ++t) /* :198-216 */
        {
            const int n = (int) (offsets[t + 1] - offsets[t]);
            for (int i = 0; i < n; ++i) tmp[i] = y[idx[offsets[t] + i]];
            slsgp_oracle_btl_derivative(n, tmp, btl_scale, dbtl);
            const double btl = slsgp_oracle_btl(n, tmp, btl_scale);
            for (int i = 0; i < n; ++i) grad[idx[offsets[t] + i]] += dbtl[i] / btl;
        }
        for (int i = 0; i < N; ++i) grad[i] += -alpha[i]; /* :219 */

        if (use_map) /* :223-257 */
        {
            double* Kinv = (double*) malloc(sizeof(double) * NN);
            memset(Kinv, 0, sizeof(double) * NN);
            for (int i = 0; i < N; ++i) Kinv[(size_t) i + (size_t) i * N] = 1.0;
            slsgp_oracle_llt_solve(N, L, N, Kinv);
            double* g_theta = (double*) malloc(sizeof(double) * (size_t) (D + 1));
            double  g_b;
            gp_hyper_gradient(kt, D, N, X, theta, Kinv, alpha, g_theta, &g_b);
            grad[N + 0] = g_theta[0] + slsgp_oracle_log_lognormal_derivative(theta[0], log(default_a), prior_var);
            grad[N + 1] = noiseless ? 0.0 : g_b + slsgp_oracle_log_lognormal_derivative(b, log(default_b), prior_var); /* :237 */
            for (int i = 0; i < D; ++i)
                grad[N + 2 + i] =
                    g_theta[1 + i] + slsgp_oracle_log_lognormal_derivative(theta[1 + i], log(default_r), prior_var);
            free(g_theta);
            free(Kinv);
        }
    }
    free(K);
    free(theta);
    return obj;
}

double slsgp_oracle_map_objective_pref(int kt, int D, int N, const double* X, int P, const unsigned* offsets,
                                       const unsigned* idx, int use_map, double default_a, double default_r,
                                       double default_b, double prior_var, double btl_scale, const double* x,
                                       double* grad)
{
    return map_objective_pref(kt, D, N, X, P, offsets, idx, use_map, default_a, default_r, default_b, prior_var, btl_scale, x, grad, 0);
}

double slsgp_oracle_map_objective_pref_noiseless(int kt, int D, int N, const double* X, int P, const unsigned* offsets,
                                                 const unsigned* idx, int use_map, double default_a, double default_r,
                                                 double default_b, double prior_var, double btl_scale, const double* x,
                                                 double* grad)
{
    return map_objective_pref(kt, D, N, X, P, offsets, idx, use_map, default_a, default_r, default_b, prior_var, btl_scale, x, grad, 1);
}

/* objective() of GaussianProcessRegressor (src/gaussian-process-regressor.cpp:141-193) with the hard-coded
 * log-normal priors of :18-24 */
double slsgp_oracle_map_objective_gpr(int kt, int D, int N, const double* X, const double* y, const double* x,
                                      double* grad)
{
    const double a_mu = log(0.500), a_s2 = 0.50, b_mu = log(1e-04), b_s2 = 0.50, r_mu = log(0.500), r_s2 = 0.50;
    const double a = x[0], b = x[1];
    double*      theta = (double*) malloc(sizeof(double) * (size_t) (D + 1));
    theta[0]           = a;
    for (int i = 0; i < D; ++i) theta[1 + i] = x[2 + i];

    const size_t NN    = (size_t) N * N;
    double*      K     = (double*) malloc(sizeof(double) * (3 * NN + (size_t) N));
    double*      L     = K + NN;
    double*      Kinv  = L + NN;
    double*      alpha = Kinv + NN;
    slsgp_oracle_large_ky(kt, D, N, X, theta, b, K);
    slsgp_oracle_cholesky(N, K, L);
    memset(Kinv, 0, sizeof(double) * NN);
    for (int i = 0; i < N; ++i) Kinv[(size_t) i + (size_t) i * N] = 1.0;
    slsgp_oracle_llt_solve(N, L, N, Kinv);
    for (int i = 0; i < N; ++i) /* alpha = K^-1 y */
    {
        double s = 0.0;
        for (int j = 0; j < N; ++j) s += Kinv[(size_t) i + (size_t) j * N] * y[j];
        alpha[i] = s;
    }

    if (grad) /* calc_grad (:108-127) */
    {
        double* g_theta = (double*) malloc(sizeof(double) * (size_t) (D + 1));
        double  g_b;
        gp_hyper_gradient(kt, D, N, X, theta, Kinv, alpha, g_theta, &g_b);
        grad[0] = g_theta[0] + slsgp_oracle_log_lognormal_derivative(a, a_mu, a_s2);
        grad[1] = g_b + slsgp_oracle_log_lognormal_derivative(b, b_mu, b_s2);
        for (int i = 0; i < D; ++i)
            grad[2 + i] = g_theta[1 + i] + slsgp_oracle_log_lognormal_derivative(theta[1 + i], r_mu, r_s2);
        free(g_theta);
    }

    const double term1 = -0.5 * dot(N, y, alpha);          /* :174 */
    const double term2 = -0.5 * slsgp_oracle_logdet(N, L); /* :175 */
    const double term3 = -0.5 * N * log(2.0 * PI);         /* :176 */
    double       reg   = slsgp_oracle_log_lognormal(a, a_mu, a_s2) + slsgp_oracle_log_lognormal(b, b_mu, b_s2);
    for (int i = 0; i < D; ++i) reg += slsgp_oracle_log_lognormal(theta[1 + i], r_mu, r_s2);
    free(K);
    free(theta);
    return term1 + term2 + term3 + reg;
}
