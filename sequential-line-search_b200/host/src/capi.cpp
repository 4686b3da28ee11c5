// extern "C" handles onto the C++ host layer, mirroring the handles the test suite binds onto the reference classes, so the
// parity tests drive both class hierarchies through ctypes with the same arguments (prefix b200_ here, ref_ there).
// C++ exceptions never cross this boundary: a failing call stores the message (b200_last_error) and returns null / NaN.
#include "device.hpp"

#include <sequential-line-search/driver.hpp>
#include <sequential-line-search/optimizers.hpp>
#include <sequential-line-search/utils.hpp>

#include "optimizer.hpp"

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <string>

using Eigen::MatrixXd;
using Eigen::VectorXd;
using namespace sequential_line_search;

namespace
{
    thread_local std::string g_error;

    KernelType          kernel_type(int kt) { return kt == 0 ? KernelType::ArdSquaredExponentialKernel : KernelType::ArdMatern52Kernel; }
    AcquisitionFuncType acq_type(int t) { return t == 0 ? AcquisitionFuncType::ExpectedImprovement : AcquisitionFuncType::GaussianProcessUpperConfidenceBound; }
    MatrixXd            matrix(const double* p, int rows, int cols)
    {
        MatrixXd m = MatrixXd::Zero(rows, cols);
        if (rows && cols) std::memcpy(m.data(), p, sizeof(double) * (size_t) rows * (size_t) cols);
        return m;
    }
    VectorXd vector(const double* p, int n)
    {
        VectorXd v = VectorXd::Zero(n);
        if (n) std::memcpy(v.data(), p, sizeof(double) * (size_t) n);
        return v;
    }
    void store(const VectorXd& v, double* out) { std::memcpy(out, v.data(), sizeof(double) * (size_t) v.size()); }
    void store(const MatrixXd& m, double* out) { std::memcpy(out, m.data(), sizeof(double) * (size_t) m.rows() * (size_t) m.cols()); }
    // names loop_capi.inl expects
    KernelType          to_kernel(int kt) { return kernel_type(kt); }
    AcquisitionFuncType to_acq(int t) { return acq_type(t); }
    MatrixXd            to_mat(const double* p, int rows, int cols) { return matrix(p, rows, cols); }
    VectorXd            to_vec(const double* p, int n) { return vector(p, n); }
    void                put(const VectorXd& v, double* out) { store(v, out); }
    void                put(const MatrixXd& m, double* out) { store(m, out); }

    template <typename F> auto guarded(F&& f, decltype(f()) on_error) -> decltype(f())
    {
        try
        {
            g_error.clear();
            return f();
        }
        catch (const std::exception& e)
        {
            g_error = e.what();
            return on_error;
        }
    }
    const double kNaN = std::numeric_limits<double>::quiet_NaN();
} // namespace

extern "C"
{
    const char* b200_last_error() { return g_error.c_str(); }

    // ---- kernels handed out by Regressor::GetKernel() & co ------------------------------------------------------
    void b200_kernel(int kt, int D, const double* xa, const double* xb, const double* theta, double* k, double* dtheta, double* dxa)
    {
        GaussianProcessRegressor probe(MatrixXd(), VectorXd(), VectorXd(), 0.0, kernel_type(kt)); // no data: no device needed
        const VectorXd           a = vector(xa, D), b = vector(xb, D), th = vector(theta, D + 1);
        if (k) *k = probe.GetKernel()(a, b, th);
        if (dtheta) store(probe.GetKernelThetaDerivative()(a, b, th), dtheta);
        if (dxa) store(probe.GetKernelFirstArgDerivative()(a, b, th), dxa);
    }
    int b200_calc_large_ky(int kt, int D, int N, const double* X, const double* theta, double b, double* K_out)
    {
        return guarded(
            [&]() {
                GaussianProcessRegressor probe(MatrixXd(), VectorXd(), VectorXd(), 0.0, kernel_type(kt));
                store(CalcLargeKY(matrix(X, D, N), vector(theta, D + 1), b, probe.GetKernel()), K_out);
                return 0;
            },
            1);
    }

    // same signatures as the reference-side test handles ref_calc_small_k / ref_calc_large_ky_theta_derivative
    int b200_calc_small_k(int kt, int D, int N, const double* X, const double* theta, const double* x, double* k_out, double* dk_dx_out)
    {
        return guarded(
            [&]() {
                GaussianProcessRegressor probe(MatrixXd(), VectorXd(), VectorXd(), 0.0, kernel_type(kt));
                const MatrixXd           Xm = matrix(X, D, N);
                const VectorXd           th = vector(theta, D + 1), xv = vector(x, D);
                store(CalcSmallK(xv, Xm, th, probe.GetKernel()), k_out);
                if (dk_dx_out) store(CalcSmallKSmallXDerivative(xv, Xm, th, probe.GetKernelFirstArgDerivative()), dk_dx_out);
                return 0;
            },
            1);
    }
    int b200_calc_large_ky_theta_derivative(int kt, int D, int N, const double* X, const double* theta, double* out)
    {
        return guarded(
            [&]() {
                GaussianProcessRegressor probe(MatrixXd(), VectorXd(), VectorXd(), 0.0, kernel_type(kt));
                const auto tensor = CalcLargeKYThetaDerivative(matrix(X, D, N), vector(theta, D + 1), probe.GetKernelThetaDerivative());
                for (size_t i = 0; i < tensor.size(); ++i) store(tensor[i], out + i * (size_t) N * (size_t) N);
                const MatrixXd I = CalcLargeKYNoiseLevelDerivative(matrix(X, D, N), vector(theta, D + 1), 0.0);
                for (int r = 0; r < N; ++r)
                    for (int c = 0; c < N; ++c)
                        if (I(r, c) != (r == c ? 1.0 : 0.0)) throw std::logic_error("CalcLargeKYNoiseLevelDerivative is not the identity");
                return 0;
            },
            1);
    }
    void b200_release_device_resources() { ReleaseDeviceResources(); }

    // ---- GaussianProcessRegressor ---------------------------------------------------------------------------------
    void* b200_gpr_copy(void* h) // the copy constructor: a regressor of its own (deep copy)
    {
        return guarded([&]() -> void* { return new GaussianProcessRegressor(*static_cast<GaussianProcessRegressor*>(h)); }, nullptr);
    }
    int b200_gpr_num_points(void* h) { return (int) static_cast<GaussianProcessRegressor*>(h)->GetLargeX().cols(); }
    void* b200_gpr_create(int kt, int D, int N, const double* X, const double* y, const double* theta, double b)
    {
        return guarded([&]() -> void* { return new GaussianProcessRegressor(matrix(X, D, N), vector(y, N), vector(theta, D + 1), b, kernel_type(kt)); },
                       nullptr);
    }
    void* b200_gpr_create_map(int kt, int D, int N, const double* X, const double* y) // hyper-parameters by MAP estimation
    {
        return guarded([&]() -> void* { return new GaussianProcessRegressor(matrix(X, D, N), vector(y, N), kernel_type(kt)); }, nullptr);
    }
    int b200_gpr_append_point(void* h, int D, const double* x, double y)
    {
        return guarded(
            [&]() {
                static_cast<GaussianProcessRegressor*>(h)->AppendPoint(vector(x, D), y);
                return 0;
            },
            1);
    }
    void        b200_gpr_destroy(void* h) { delete static_cast<GaussianProcessRegressor*>(h); }
    const void* b200_gpr_regressor(void* h) { return static_cast<const Regressor*>(static_cast<GaussianProcessRegressor*>(h)); }
    void        b200_gpr_get_state(void* hv, double* K_y, double* K_y_inv, double* theta, double* b)
    {
        const auto& r = *static_cast<GaussianProcessRegressor*>(hv);
        if (K_y) store(r.m_K_y, K_y);
        if (K_y_inv) store(r.m_K_y_inv, K_y_inv);
        if (theta) store(r.GetKernelHyperparams(), theta);
        if (b) *b = r.GetNoiseHyperparam();
    }

    // ---- PreferenceRegressor (tuples in CSR form, first member preferred) ---------------------------------------------
    void* b200_pref_create(int kt, int D, int N, const double* X, int P, const unsigned* offsets, const unsigned* idx, int use_map,
                           double a, double r, double b, double prior_var, double btl_scale, unsigned num_iters)
    {
        return guarded(
            [&]() -> void* {
                std::vector<Preference> prefs;
                for (int t = 0; t < P; ++t) prefs.push_back(Preference(std::vector<unsigned>(idx + offsets[t], idx + offsets[t + 1])));
                return new PreferenceRegressor(matrix(X, D, N), prefs, use_map != 0, a, r, b, prior_var, btl_scale, num_iters, kernel_type(kt));
            },
            nullptr);
    }
    // same, warm-started from the state of a previous regressor `prev` (SequentialLineSearchOptimizer does this between iterations)
    void* b200_pref_create_warm(const void* prev, int kt, int D, int N, const double* X, int P, const unsigned* offsets, const unsigned* idx,
                                int use_map, double a, double r, double b, double prior_var, double btl_scale, unsigned num_iters)
    {
        return guarded(
            [&]() -> void* {
                const auto&  pr = *static_cast<const PreferenceRegressor*>(prev);
                MapWarmStart w;
                w.X = pr.GetLargeX(), w.y = pr.GetSmallY(), w.kernel_hyperparams = pr.GetKernelHyperparams(), w.noise_hyperparam = pr.GetNoiseHyperparam();
                std::vector<Preference> prefs;
                for (int t = 0; t < P; ++t) prefs.push_back(Preference(std::vector<unsigned>(idx + offsets[t], idx + offsets[t + 1])));
                return new PreferenceRegressor(matrix(X, D, N), prefs, use_map != 0, a, r, b, prior_var, btl_scale, num_iters, kernel_type(kt), &w);
            },
            nullptr);
    }
    void        b200_pref_destroy(void* h) { delete static_cast<PreferenceRegressor*>(h); }
    const void* b200_pref_regressor(void* h) { return static_cast<const Regressor*>(static_cast<PreferenceRegressor*>(h)); }
    double      b200_pref_objective(void* hv, const double* x, int n, double* grad /* may be null */)
    {
        return guarded(
            [&]() {
                VectorXd     g;
                const double f = static_cast<PreferenceRegressor*>(hv)->EvaluateMapObjective(vector(x, n), grad ? &g : nullptr);
                if (grad) store(g, grad);
                return f;
            },
            kNaN);
    }
    unsigned b200_pref_num_map_evaluations(void* hv) { return static_cast<PreferenceRegressor*>(hv)->GetNumMapEvaluations(); }
    int      b200_pref_num_points_kept(void* hv) { return static_cast<PreferenceRegressor*>(hv)->GetNumPointsKept(); }
    void     b200_pref_get_state(void* hv, double* y, double* theta, double* b, double* K, double* L)
    {
        const PreferenceRegressor& r = *static_cast<PreferenceRegressor*>(hv);
        if (y) store(r.GetSmallY(), y);
        if (theta) store(r.m_kernel_hyperparams, theta);
        if (b) *b = r.m_noise_hyperparam;
        if (K) store(r.m_K, K);
        if (L) store(r.m_L, L);
    }
    void b200_pref_find_arg_max(void* hv, double* x_out) { store(static_cast<PreferenceRegressor*>(hv)->FindArgMax(), x_out); }
    int  b200_pref_damp_data(void* hv, const char* dir, const char* prefix)
    {
        return guarded([&]() { return static_cast<PreferenceRegressor*>(hv)->DampData(dir, prefix), 0; }, 1);
    }

    // ---- PreferenceDataManager / Slider (host bookkeeping) -------------------------------------------------------------------
    // Feeds `n_batches` batches to AddNewPoints; batch b holds sizes[b] points (first = the preferred one), stored back to
    // back in `points` (D doubles each). Returns the final number of points; X_out (D x N) and the tuples in CSR form.
    int b200_data_manager_run(int D, int n_batches, const int* sizes, const double* points, double eps, double* X_out, unsigned* offsets_out,
                              unsigned* idx_out)
    {
        PreferenceDataManager dm;
        const double*         p = points;
        for (int b = 0; b < n_batches; ++b)
        {
            const VectorXd        first = vector(p, D);
            std::vector<VectorXd> others;
            for (int k = 1; k < sizes[b]; ++k) others.push_back(vector(p + (size_t) k * D, D));
            p += (size_t) sizes[b] * D;
            dm.AddNewPoints(first, others, true, eps);
        }
        store(dm.GetX(), X_out);
        unsigned n = 0;
        offsets_out[0] = 0;
        for (size_t t = 0; t < dm.GetD().size(); ++t)
        {
            for (unsigned i : dm.GetD()[t]) idx_out[n++] = i;
            offsets_out[t + 1] = n;
        }
        return dm.GetNumDataPoints();
    }
    void b200_slider(int D, const double* end_0, const double* end_1, int enlarge, double scale, double minimum_length, double* out_0, double* out_1)
    {
        const Slider s(vector(end_0, D), vector(end_1, D), enlarge != 0, scale, minimum_length);
        store(s.end_0, out_0);
        store(s.end_1, out_1);
    }

    // ---- the bound-constrained quasi-Newton driver on two closed-form problems (CPU-only test hook) ---------------------------
    // problem 0: sum_i w_i (x_i - t_i)^2 with w_i = 1 + i and targets t_i = 2 i / n - 0.5 (some outside the box [0, 1]^n);
    // problem 1: the Rosenbrock valley in n dimensions on [-2, 2]^n. Returns the number of evaluations; x_out holds the solution.
    // ---- utils (include/sequential-line-search/utils.hpp) ---------------------------------------------------------------------
    double b200_utils_btl(int n, const double* f, double scale, double* derivative /* n values or null */)
    {
        const VectorXd fv = vector(f, n);
        if (derivative) store(utils::CalcBtlDerivative(fv, scale), derivative);
        return utils::CalcBtl(fv, scale);
    }
    void b200_utils_random_vector(unsigned n, double* out) { store(utils::GenerateRandomVector(n), out); }
    int  b200_utils_export_csv(const char* path, int rows, int cols, const double* X)
    {
        return guarded(
            [&]() {
                utils::ExportMatrixToCsv(path, matrix(X, rows, cols));
                return 0;
            },
            1);
    }

    int b200_test_minimize(int problem, int n, const double* x0, unsigned max_evals, double* x_out, double* f_out)
    {
        std::vector<double> lo((size_t) n, problem == 0 ? 0.0 : -2.0), hi((size_t) n, problem == 0 ? 1.0 : 2.0), start(x0, x0 + n);
        const internal::Objective fun = [&](const std::vector<double>& x, std::vector<double>& g) {
            double f = 0.0;
            if (problem == 0)
                for (int i = 0; i < n; ++i)
                {
                    const double w = 1.0 + i, t = 2.0 * i / n - 0.5;
                    f += w * (x[(size_t) i] - t) * (x[(size_t) i] - t), g[(size_t) i] = 2.0 * w * (x[(size_t) i] - t);
                }
            else
            {
                for (auto& v : g) v = 0.0;
                for (int i = 0; i + 1 < n; ++i)
                {
                    const double a = x[(size_t) i + 1] - x[(size_t) i] * x[(size_t) i], b = 1.0 - x[(size_t) i];
                    f += 100.0 * a * a + b * b;
                    g[(size_t) i] += -400.0 * a * x[(size_t) i] - 2.0 * b, g[(size_t) i + 1] += 200.0 * a;
                }
            }
            return f;
        };
        const internal::MinimizeResult r = internal::minimize_bounded(fun, start, lo, hi, max_evals, 1e-9);
        std::memcpy(x_out, r.x.data(), sizeof(double) * (size_t) n);
        *f_out = r.f;
        return (int) r.evals;
    }

    // ---- search driver (sequential-line-search/driver.hpp): 0 Native, 1 Hybrid, 2 Reference -------------------------------------
    int b200_nlopt_available() { return IsNloptAvailable() ? 1 : 0; }
    int b200_get_search_driver() { return (int) GetSearchDriver(); }
    int b200_set_search_driver(int mode)
    {
        return guarded([&]() { return SetSearchDriver(mode == 0 ? SearchDriver::Native : mode == 1 ? SearchDriver::Hybrid : SearchDriver::Reference), 0; }, 1);
    }

    int b200_get_incremental_refit() { return GetIncrementalRefit() ? 1 : 0; }
    int b200_set_incremental_refit(int on) { return SetIncrementalRefit(on != 0), 0; }

    int b200_set_devices(const int* ids, int n)
    {
        return guarded([&]() { return SetDevices(std::vector<int>(ids, ids + n)), 0; }, 1);
    }
    int b200_get_device_count() { return (int) GetDevices().size(); }

    // ---- Regressor virtual interface ------------------------------------------------------------------------------------
    double b200_predict_mu(const void* r, int D, const double* x)
    {
        return guarded([&]() { return static_cast<const Regressor*>(r)->PredictMu(vector(x, D)); }, kNaN);
    }
    double b200_predict_sigma(const void* r, int D, const double* x)
    {
        return guarded([&]() { return static_cast<const Regressor*>(r)->PredictSigma(vector(x, D)); }, kNaN);
    }
    int b200_predict_mu_derivative(const void* r, int D, const double* x, double* out)
    {
        return guarded([&]() { return store(static_cast<const Regressor*>(r)->PredictMuDerivative(vector(x, D)), out), 0; }, 1);
    }
    int b200_predict_sigma_derivative(const void* r, int D, const double* x, double* out)
    {
        return guarded([&]() { return store(static_cast<const Regressor*>(r)->PredictSigmaDerivative(vector(x, D)), out), 0; }, 1);
    }
    int b200_predict_maximum_point_from_data(const void* r, int D, double* out)
    {
        return guarded([&]() { return store(static_cast<const Regressor*>(r)->PredictMaximumPointFromData(), out), 0; }, 1);
    }
    int b200_predict_batch(const void* r, int D, int M, const double* Xq, double* mu, double* sigma, double* dmu, double* dsigma)
    {
        return guarded(
            [&]() {
                const auto* d = dynamic_cast<const DeviceRegressor*>(static_cast<const Regressor*>(r));
                if (!d) throw std::invalid_argument("not a device-backed regressor");
                VectorXd m, s;
                MatrixXd dm, ds;
                d->PredictBatch(matrix(Xq, D, M), mu ? &m : nullptr, sigma ? &s : nullptr, dmu ? &dm : nullptr, dsigma ? &ds : nullptr);
                if (mu) store(m, mu);
                if (sigma) store(s, sigma);
                if (dmu) store(dm, dmu);
                if (dsigma) store(ds, dsigma);
                return 0;
            },
            1);
    }

    // Arithmetic of the candidate sweep behind this regressor's batched calls (slsgp_sweep_mode: 0 IEEE double, 1..3 tensor-core modes)
    int b200_regressor_set_sweep_mode(const void* r, int mode)
    {
        return guarded(
            [&]() {
                const auto* d = dynamic_cast<const DeviceRegressor*>(static_cast<const Regressor*>(r));
                if (!d || !d->HasModel()) throw std::invalid_argument("not a device-backed regressor with a model");
                slsgp_ctx* const            c = d->Device();
                std::lock_guard<std::mutex> lock(d->DeviceMutex());
                internal::check(c, slsgp_set_sweep_mode(c, (slsgp_sweep_mode) mode), "slsgp_set_sweep_mode");
                return 0;
            },
            1);
    }

    // ---- acquisition_func --------------------------------------------------------------------------------------------------
    double b200_acq_value(const void* r, int D, int acq, double ucb_beta, const double* x)
    {
        return guarded([&]() { return acquisition_func::CalcAcquisitionValue(*static_cast<const Regressor*>(r), vector(x, D), acq_type(acq), ucb_beta); }, kNaN);
    }
    int b200_acq_derivative(const void* r, int D, int acq, double ucb_beta, const double* x, double* out)
    {
        return guarded(
            [&]() { return store(acquisition_func::CalcAcquisitionValueDerivative(*static_cast<const Regressor*>(r), vector(x, D), acq_type(acq), ucb_beta), out), 0; }, 1);
    }
    int b200_acq_values(const void* r, int D, int M, int acq, double ucb_beta, const double* Xq, double* val, double* grad)
    {
        return guarded(
            [&]() {
                const auto* d = dynamic_cast<const DeviceRegressor*>(static_cast<const Regressor*>(r));
                if (!d) throw std::invalid_argument("not a device-backed regressor");
                MatrixXd g;
                store(acquisition_func::CalcAcquisitionValues(*d, matrix(Xq, D, M), acq_type(acq), ucb_beta, grad ? &g : nullptr), val);
                if (grad) store(g, grad);
                return 0;
            },
            1);
    }
    // Seconds per call of acquisition_func::CalcAcquisitionValues (values + gradients for M candidates held in an Eigen matrix,
    // results returned as Eigen objects: pageable memory both ways, allocations included - what a C++ caller of the drop-in pays),
    // mean over `reps` calls after one warm-up call. The candidate matrix is built once, outside the timed region.
    double b200_time_acq_values(const void* r, int D, int M, int acq, double ucb_beta, int reps)
    {
        return guarded(
            [&]() {
                const auto* d = dynamic_cast<const DeviceRegressor*>(static_cast<const Regressor*>(r));
                if (!d) throw std::invalid_argument("not a device-backed regressor");
                MatrixXd Xq = MatrixXd::Zero(D, M);
                uint64_t state = 0x9E3779B97F4A7C15ull;
                for (long i = 0; i < (long) D * M; ++i)
                {
                    state        = state * 6364136223846793005ull + 1442695040888963407ull;
                    Xq.data()[i] = (double) (state >> 11) * (1.0 / 9007199254740992.0);
                }
                MatrixXd grads;
                double   sink = 0.0;
                {
                    const VectorXd v = acquisition_func::CalcAcquisitionValues(*d, Xq, acq_type(acq), ucb_beta, &grads);
                    sink += v(0);
                }
                const auto t0 = std::chrono::steady_clock::now();
                for (int i = 0; i < reps; ++i)
                {
                    const VectorXd v = acquisition_func::CalcAcquisitionValues(*d, Xq, acq_type(acq), ucb_beta, &grads);
                    sink += v(M - 1) + grads(0, 0);
                }
                const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                return sink == 12345.678 ? -1.0 : dt / reps;
            },
            kNaN);
    }

    int b200_find_next_point(const void* r, int D, unsigned n_global, unsigned n_local, int acq, double ucb_beta, double* x_out)
    {
        return guarded(
            [&]() { return store(acquisition_func::FindNextPoint(*static_cast<const Regressor*>(r), n_global, n_local, acq_type(acq), ucb_beta), x_out), 0; }, 1);
    }
    int b200_find_next_points(const void* r, int D, unsigned n_points, unsigned n_global, unsigned n_local, int acq, double ucb_beta, double* X_out)
    {
        return guarded(
            [&]() {
                const auto pts = acquisition_func::FindNextPoints(*static_cast<const Regressor*>(r), n_points, n_global, n_local, acq_type(acq), ucb_beta);
                for (size_t i = 0; i < pts.size(); ++i) store(pts[i], X_out + i * (size_t) D);
                return 0;
            },
            1);
    }
    // ---- optimiser front-ends and driver-dependent entry points: the facade the test suite also compiles against the reference's classes -------
#define SLS_CAPI(name) b200_##name
#define SLS_CAPI_TRY \
    g_error.clear(); \
    try
#define SLS_CAPI_CATCH(value)       \
    catch (const std::exception& e) \
    {                               \
        g_error = e.what();         \
        return value;               \
    }
#include "loop_capi.inl"
    // incremental refit diagnostics (no reference counterpart): how many data points of the optimiser's current regressor kept
    // the factored model of the previous iteration
    // out: map_fit, search, slider (ms of the last SubmitFeedbackData) and the number of MAP objective evaluations
    void b200_sls_last_step_timings(void* h, double* out)
    {
        const auto& o = *static_cast<b200_SlsHandle*>(h)->opt;
        const auto  t = o.GetLastStepTimings();
        const auto  r = o.GetRegressor();
        out[0] = t.map_fit, out[1] = t.search, out[2] = t.slider, out[3] = r ? (double) r->GetNumMapEvaluations() : 0.0;
    }
    // The optimiser's current regressor against a deep copy of it (a copy rebuilds its device model from scratch out of X, y and the
    // hyper-parameters): largest difference of mu, sigma and their gradients over the M query points (D x M). With the incremental
    // refit the left side is an EXTENDED model, the right side a rebuilt one of the same data and the same goodness values.
    double b200_sls_model_vs_rebuilt_copy(void* h, int D, int M, const double* Xq)
    {
        return guarded(
            [&]() {
                const auto r = static_cast<b200_SlsHandle*>(h)->opt->GetRegressor();
                if (!r) return -1.0;
                const PreferenceRegressor copy(*r);
                double                    worst = 0.0;
                for (int m = 0; m < M; ++m)
                {
                    const VectorXd x = vector(Xq + (size_t) m * D, D);
                    worst            = std::max(worst, std::abs(r->PredictMu(x) - copy.PredictMu(x)));
                    worst            = std::max(worst, std::abs(r->PredictSigma(x) - copy.PredictSigma(x)));
                    const VectorXd dm = r->PredictMuDerivative(x) - copy.PredictMuDerivative(x), ds = r->PredictSigmaDerivative(x) - copy.PredictSigmaDerivative(x);
                    for (int d = 0; d < D; ++d) worst = std::max(worst, std::max(std::abs(dm(d)), std::abs(ds(d))));
                }
                return worst;
            },
            kNaN);
    }
    int b200_sls_num_points_kept(void* h)
    {
        const auto r = static_cast<b200_SlsHandle*>(h)->opt->GetRegressor();
        return r ? r->GetNumPointsKept() : 0;
    }
}
