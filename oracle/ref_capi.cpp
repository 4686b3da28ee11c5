// TEST INFRASTRUCTURE (oracle/): extern "C" handles onto the *unmodified* reference sources.
//
// Built by oracle/Makefile together with /root/reference/src/*.cpp and the four mathtoolbox sources the hot path
// uses, against include/eigen-lite (Eigen is a system dependency of the reference and is not installed here) and
// oracle/nlopt_probe/nlopt.hpp. Output: oracle/_ref/libsls_ref_probe.so. Only tests/, the golden-vector generator
// and bench.py's reference/cpu_baseline legs load it. It contains no arithmetic of its own: every number that
// comes out of it is computed by the reference's code (with eigen-lite's LLT / LU standing in for Eigen's).
#include <chrono>
#include <cstring>
#include <memory>
#include <nlopt.hpp>
#include <sequential-line-search/acquisition-function.hpp>
#include <sequential-line-search/gaussian-process-regressor.hpp>
#include <sequential-line-search/preference-data-manager.hpp>
#include <sequential-line-search/preference-regressor.hpp>
#include <sequential-line-search/regressor.hpp>
#include <sequential-line-search/utils.hpp>
#include <mathtoolbox/kernel-functions.hpp>
#include <vector>

using Eigen::MatrixXd;
using Eigen::VectorXd;
using namespace sequential_line_search;

namespace
{
    KernelType to_kernel_type(int kt)
    {
        return kt == 0 ? KernelType::ArdSquaredExponentialKernel : KernelType::ArdMatern52Kernel;
    }
    AcquisitionFuncType to_acq_type(int t)
    {
        return t == 0 ? AcquisitionFuncType::ExpectedImprovement
                      : AcquisitionFuncType::GaussianProcessUpperConfidenceBound;
    }
    MatrixXd to_matrix(const double* p, int rows, int cols)
    {
        MatrixXd m(rows, cols);
        std::memcpy(m.data(), p, sizeof(double) * size_t(rows) * size_t(cols));
        return m;
    }
    VectorXd to_vector(const double* p, int n)
    {
        VectorXd v(n);
        std::memcpy(v.data(), p, sizeof(double) * size_t(n));
        return v;
    }

    // Kernel dispatch exactly as Regressor::Regressor does it (src/regressor.cpp:8-27).
    struct KernelProbe : public Regressor
    {
        KernelProbe(KernelType t) : Regressor(t) {}
        double          PredictMu(const VectorXd&) const override { return 0; }
        double          PredictSigma(const VectorXd&) const override { return 0; }
        VectorXd        PredictMuDerivative(const VectorXd& x) const override { return x; }
        VectorXd        PredictSigmaDerivative(const VectorXd& x) const override { return x; }
        const VectorXd& GetKernelHyperparams() const override { return v; }
        double          GetNoiseHyperparam() const override { return 0; }
        const MatrixXd& GetLargeX() const override { return m; }
        const VectorXd& GetSmallY() const override { return v; }
        VectorXd        v;
        MatrixXd        m;
    };

    struct PrefHandle
    {
        std::unique_ptr<PreferenceRegressor> reg;
        nlopt::vfunc                         objective = nullptr; // captured anonymous-namespace objective
        void*                                data      = nullptr;
        unsigned                             opt_dim   = 0;
    };
} // namespace

extern "C"
{
    // ---- L0/L1: kernels and kernel matrices --------------------------------------------------------------
    void ref_kernel(int kt, int D, const double* xa, const double* xb, const double* theta, double* k,
                    double* dtheta, double* dxa)
    {
        KernelProbe    p(to_kernel_type(kt));
        const VectorXd a = to_vector(xa, D), b = to_vector(xb, D), th = to_vector(theta, D + 1);
        if (k) *k = p.GetKernel()(a, b, th);
        if (dtheta)
        {
            const VectorXd g = p.GetKernelThetaDerivative()(a, b, th);
            std::memcpy(dtheta, g.data(), sizeof(double) * size_t(D + 1));
        }
        if (dxa)
        {
            const VectorXd g = p.GetKernelFirstArgDerivative()(a, b, th);
            std::memcpy(dxa, g.data(), sizeof(double) * size_t(D));
        }
    }

    void ref_calc_large_ky(int kt, int D, int N, const double* X, const double* theta, double b, double* K_out)
    {
        KernelProbe    p(to_kernel_type(kt));
        const MatrixXd K = CalcLargeKY(to_matrix(X, D, N), to_vector(theta, D + 1), b, p.GetKernel());
        std::memcpy(K_out, K.data(), sizeof(double) * size_t(N) * size_t(N));
    }

    // CalcLargeKY + Eigen::LLT exactly as PreferenceRegressor's constructor runs them (src/preference-regressor.cpp:289-290);
    // returns the wall time of the two statements in seconds (bench.py: the reference side of "Gram + Chol ms").
    double ref_gram_chol_seconds(int kt, int D, int N, const double* X, const double* theta, double b, double* L_out /* N x N or null */)
    {
        KernelProbe    p(to_kernel_type(kt));
        const MatrixXd Xm = to_matrix(X, D, N);
        const VectorXd th = to_vector(theta, D + 1);
        const auto     t0 = std::chrono::steady_clock::now();
        const MatrixXd             K = CalcLargeKY(Xm, th, b, p.GetKernel());
        const Eigen::LLT<MatrixXd> llt(K);
        const double               dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (L_out)
        {
            const MatrixXd L = llt.matrixL().toDenseMatrix();
            std::memcpy(L_out, L.data(), sizeof(double) * size_t(N) * size_t(N));
        }
        return dt;
    }

    void ref_calc_small_k(int kt, int D, int N, const double* X, const double* theta, const double* x,
                          double* k_out, double* dk_dx_out /* D x N col-major, may be null */)
    {
        KernelProbe    p(to_kernel_type(kt));
        const MatrixXd Xm = to_matrix(X, D, N);
        const VectorXd th = to_vector(theta, D + 1), xv = to_vector(x, D);
        const VectorXd k  = CalcSmallK(xv, Xm, th, p.GetKernel());
        std::memcpy(k_out, k.data(), sizeof(double) * size_t(N));
        if (dk_dx_out)
        {
            const MatrixXd J = CalcSmallKSmallXDerivative(xv, Xm, th, p.GetKernelFirstArgDerivative());
            std::memcpy(dk_dx_out, J.data(), sizeof(double) * size_t(D) * size_t(N));
        }
    }

    // (D+1) matrices of N x N, concatenated.
    void ref_calc_large_ky_theta_derivative(int kt, int D, int N, const double* X, const double* theta,
                                            double* out)
    {
        KernelProbe p(to_kernel_type(kt));
        const auto  tensor =
            CalcLargeKYThetaDerivative(to_matrix(X, D, N), to_vector(theta, D + 1), p.GetKernelThetaDerivative());
        for (size_t i = 0; i < tensor.size(); ++i)
            std::memcpy(out + i * size_t(N) * size_t(N), tensor[i].data(), sizeof(double) * size_t(N) * size_t(N));
    }

#ifndef SLS_REF_REAL_NLOPT // these reach the anonymous-namespace MAP objective through the probe NLopt's hook
    // ---- PreferenceRegressor -------------------------------------------------------------------------------
    // Tuples in CSR form: tuple t covers idx[offsets[t] .. offsets[t+1]). `solution` (length N, or N+2+D when
    // use_map != 0, layout [y, a, b, r_1..r_D] as in src/preference-regressor.cpp:137-147) is what the probe NLopt
    // "returns", i.e. the state the regressor ends up in.
    void* ref_pref_create(int kt, int D, int N, const double* X, int P, const unsigned* offsets,
                          const unsigned* idx, int use_map, double a, double r, double b, double prior_var,
                          double btl_scale, const double* solution)
    {
        std::vector<Preference> prefs;
        for (int t = 0; t < P; ++t)
            prefs.push_back(Preference(std::vector<unsigned>(idx + offsets[t], idx + offsets[t + 1])));

        auto* h = new PrefHandle;
        nlopt::probe::hook().fn =
            [&](const nlopt::opt& solver, nlopt::vfunc f, void* data, std::vector<double>& x, double& fval)
        {
            h->objective = f;
            h->data      = data;
            h->opt_dim   = solver.get_dimension();
            if (solution) std::memcpy(x.data(), solution, sizeof(double) * x.size());
            std::vector<double> no_grad;
            fval = f(x, no_grad, data);
        };
        h->reg.reset(new PreferenceRegressor(to_matrix(X, D, N), prefs, use_map != 0, a, r, b, prior_var, btl_scale,
                                             100, to_kernel_type(kt)));
        nlopt::probe::hook().fn = nullptr;
        return h;
    }
    void ref_pref_destroy(void* h) { delete static_cast<PrefHandle*>(h); }

    const void* ref_pref_regressor(void* h) { return static_cast<const Regressor*>(static_cast<PrefHandle*>(h)->reg.get()); }

    // The reference's MAP objective (src/preference-regressor.cpp:129-259) at an arbitrary point.
    double ref_pref_objective(void* hv, const double* x, double* grad /* may be null */)
    {
        auto*               h = static_cast<PrefHandle*>(hv);
        std::vector<double> xv(x, x + h->opt_dim);
        std::vector<double> g(grad ? h->opt_dim : 0);
        const double        f = h->objective(xv, g, h->data);
        if (grad) std::memcpy(grad, g.data(), sizeof(double) * g.size());
        return f;
    }
    unsigned ref_pref_opt_dim(void* hv) { return static_cast<PrefHandle*>(hv)->opt_dim; }

    void ref_pref_get_state(void* hv, double* y, double* theta, double* b, double* K, double* L)
    {
        const PreferenceRegressor& r = *static_cast<PrefHandle*>(hv)->reg;
        const size_t               N = size_t(r.m_X.cols());
        if (y) std::memcpy(y, r.GetSmallY().data(), sizeof(double) * N);
        if (theta) std::memcpy(theta, r.m_kernel_hyperparams.data(), sizeof(double) * size_t(r.m_kernel_hyperparams.size()));
        if (b) *b = r.m_noise_hyperparam;
        if (K) std::memcpy(K, r.m_K.data(), sizeof(double) * N * N);
        if (L)
        {
            const MatrixXd Lm = r.m_K_llt.matrixL().toDenseMatrix();
            std::memcpy(L, Lm.data(), sizeof(double) * N * N);
        }
    }
    void ref_pref_find_arg_max(void* hv, double* x_out)
    {
        const VectorXd x = static_cast<PrefHandle*>(hv)->reg->FindArgMax();
        std::memcpy(x_out, x.data(), sizeof(double) * size_t(x.size()));
    }

#endif

    // ---- GaussianProcessRegressor ----------------------------------------------------------------------------
    void* ref_gpr_create(int kt, int D, int N, const double* X, const double* y, const double* theta, double b)
    {
        return new GaussianProcessRegressor(to_matrix(X, D, N), to_vector(y, N), to_vector(theta, D + 1), b,
                                            to_kernel_type(kt));
    }
    void        ref_gpr_destroy(void* h) { delete static_cast<GaussianProcessRegressor*>(h); }
    const void* ref_gpr_regressor(void* h) { return static_cast<const Regressor*>(static_cast<GaussianProcessRegressor*>(h)); }
    void        ref_gpr_get_state(void* hv, double* K_y, double* K_y_inv)
    {
        const auto&  r = *static_cast<GaussianProcessRegressor*>(hv);
        const size_t N = size_t(r.m_K_y.rows());
        if (K_y) std::memcpy(K_y, r.m_K_y.data(), sizeof(double) * N * N);
        if (K_y_inv) std::memcpy(K_y_inv, r.m_K_y_inv.data(), sizeof(double) * N * N);
    }

#ifndef SLS_REF_REAL_NLOPT
    // The reference's GPR MAP objective (src/gaussian-process-regressor.cpp:141-193), variables (a, b, r_1..r_D),
    // evaluated at n_points points. It is only reachable from inside the MAP constructor, so the probe hook
    // evaluates it there; the constructor then finishes at the last probe point.
    void ref_gpr_objective(int kt, int D, int N, const double* X, const double* y, int n_points,
                           const double* points /* n_points x (D+2) row-major */, double* f_out,
                           double* grad_out /* n_points x (D+2) or null */)
    {
        const int dim           = D + 2;
        bool      done          = false;
        nlopt::probe::hook().fn =
            [&](const nlopt::opt&, nlopt::vfunc f, void* data, std::vector<double>& x, double& fval)
        {
            if (!done)
            {
                for (int p = 0; p < n_points; ++p)
                {
                    std::vector<double> xv(points + p * dim, points + (p + 1) * dim);
                    std::vector<double> g(grad_out ? dim : 0);
                    f_out[p] = f(xv, g, data);
                    if (grad_out) std::memcpy(grad_out + p * dim, g.data(), sizeof(double) * size_t(dim));
                }
                done = true;
            }
            std::memcpy(x.data(), points + (n_points - 1) * dim, sizeof(double) * size_t(dim));
            fval = f_out[n_points - 1];
        };
        GaussianProcessRegressor reg(to_matrix(X, D, N), to_vector(y, N), to_kernel_type(kt));
        nlopt::probe::hook().fn = nullptr;
    }
#endif

    // ---- Regressor virtual interface + acquisition (work on either regressor kind) -----------------------
    double ref_predict_mu(const void* r, int D, const double* x)
    {
        return static_cast<const Regressor*>(r)->PredictMu(to_vector(x, D));
    }
    double ref_predict_sigma(const void* r, int D, const double* x)
    {
        return static_cast<const Regressor*>(r)->PredictSigma(to_vector(x, D));
    }
    void ref_predict_mu_derivative(const void* r, int D, const double* x, double* out)
    {
        const VectorXd g = static_cast<const Regressor*>(r)->PredictMuDerivative(to_vector(x, D));
        std::memcpy(out, g.data(), sizeof(double) * size_t(D));
    }
    void ref_predict_sigma_derivative(const void* r, int D, const double* x, double* out)
    {
        const VectorXd g = static_cast<const Regressor*>(r)->PredictSigmaDerivative(to_vector(x, D));
        std::memcpy(out, g.data(), sizeof(double) * size_t(D));
    }
    void ref_predict_maximum_point_from_data(const void* r, int D, double* out)
    {
        const VectorXd x = static_cast<const Regressor*>(r)->PredictMaximumPointFromData();
        std::memcpy(out, x.data(), sizeof(double) * size_t(D));
    }
    double ref_acq_value(const void* r, int D, int acq_type, double ucb_beta, const double* x)
    {
        return acquisition_func::CalcAcquisitionValue(*static_cast<const Regressor*>(r), to_vector(x, D),
                                                      to_acq_type(acq_type), ucb_beta);
    }
    void ref_acq_derivative(const void* r, int D, int acq_type, double ucb_beta, const double* x, double* out)
    {
        const VectorXd g = acquisition_func::CalcAcquisitionValueDerivative(
            *static_cast<const Regressor*>(r), to_vector(x, D), to_acq_type(acq_type), ucb_beta);
        std::memcpy(out, g.data(), sizeof(double) * size_t(D));
    }

    // ---- BTL (include/sequential-line-search/utils.hpp:25-52) ------------------------------------------------
    double ref_btl(int n, const double* f, double scale, double* derivative /* n or null */)
    {
        const VectorXd fv = to_vector(f, n);
        if (derivative)
        {
            const VectorXd d = utils::CalcBtlDerivative(fv, scale);
            std::memcpy(derivative, d.data(), sizeof(double) * size_t(n));
        }
        return utils::CalcBtl(fv, scale);
    }
    // ---- PreferenceDataManager (src/preference-data-manager.cpp:88-141, with MergeClosePoints :14-86) -------------------
    // Same calling convention as the host facade's b200_data_manager_run.
    int ref_data_manager_run(int D, int n_batches, const int* sizes, const double* points, double eps, double* X_out,
                             unsigned* offsets_out, unsigned* idx_out)
    {
        PreferenceDataManager dm;
        const double*         p = points;
        for (int b = 0; b < n_batches; ++b)
        {
            const VectorXd        first = to_vector(p, D);
            std::vector<VectorXd> others;
            for (int k = 1; k < sizes[b]; ++k) others.push_back(to_vector(p + size_t(k) * D, D));
            p += size_t(sizes[b]) * D;
            dm.AddNewPoints(first, others, true, eps);
        }
        std::memcpy(X_out, dm.GetX().data(), sizeof(double) * size_t(D) * size_t(dm.GetNumDataPoints()));
        unsigned n     = 0;
        offsets_out[0] = 0;
        for (size_t t = 0; t < dm.GetD().size(); ++t)
        {
            for (unsigned i : dm.GetD()[t]) idx_out[n++] = i;
            offsets_out[t + 1] = n;
        }
        return dm.GetNumDataPoints();
    }
}
