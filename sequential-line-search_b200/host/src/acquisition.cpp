// acquisition_func::* (src/acquisition-function.cpp:170-298 of the reference) on libslsgp.
//   CalcAcquisitionValue{,Derivative}   one-candidate calls of the device sweep (slsgp_acq_batch)
//   CalcAcquisitionValues               the batched form the GPU path exists for
//   FindNextPoint                       FindGlobalSolution (:112-167): dense candidate sweep on the device instead of
//                                       NLopt DIRECT, then a bound-constrained quasi-Newton polish instead of LD_LBFGS
//   FindNextPoints                      Schonlau's batch criterion (:246-298): mu from the regressor, sigma from a
//                                       temporary GaussianProcessRegressor that also holds the points chosen so far
#include "device.hpp"
#include "nlopt_driver.hpp"
#include "optimizer.hpp"

#include <cmath>
#include <limits>

using Eigen::MatrixXd;
using Eigen::VectorXd;

namespace sequential_line_search
{
    using internal::check;

    namespace
    {
        // Scalar Expected Improvement / GP-UCB for a regressor that is NOT device-backed (a foreign subclass of
        // Regressor): mathtoolbox acquisition-functions.cpp:8-78 through the Predict* virtuals.
        double pdf(double z) { return (1.0 / std::sqrt(2.0 * 3.14159265358979323846)) * std::exp(-0.5 * z * z); }
        double cdf(double z) { return 0.5 * (1.0 + std::erf(z / std::sqrt(2.0))); }

        double generic_value(const Regressor& r, const VectorXd& x, AcquisitionFuncType type, double beta)
        {
            const double mu = r.PredictMu(x), sigma = r.PredictSigma(x);
            if (type == AcquisitionFuncType::GaussianProcessUpperConfidenceBound) return mu + beta * sigma;
            const double diff = mu - r.PredictMu(r.PredictMaximumPointFromData()), z = diff / sigma;
            const double ei   = diff * cdf(z) + sigma * pdf(z);
            return (sigma < 1e-16 || std::isnan(ei)) ? 0.0 : ei;
        }
        VectorXd generic_derivative(const Regressor& r, const VectorXd& x, AcquisitionFuncType type, double beta)
        {
            const double   mu = r.PredictMu(x), sigma = r.PredictSigma(x);
            const VectorXd dmu = r.PredictMuDerivative(x), dsg = r.PredictSigmaDerivative(x);
            VectorXd       g = VectorXd::Zero(x.size());
            if (type == AcquisitionFuncType::GaussianProcessUpperConfidenceBound)
            {
                for (int i = 0; i < (int) x.size(); ++i) g(i) = dmu(i) + beta * dsg(i);
                return g;
            }
            const double diff = mu - r.PredictMu(r.PredictMaximumPointFromData()), z = diff / sigma;
            bool         bad  = sigma < 1e-16;
            for (int i = 0; i < (int) x.size(); ++i)
            {
                const double dz = (dmu(i) - z * dsg(i)) / sigma;
                g(i)            = dmu(i) * cdf(z) + diff * dz * pdf(z) + dsg(i) * pdf(z) + sigma * dz * (-z * pdf(z));
                bad |= std::isnan(g(i));
            }
            return bad ? VectorXd::Zero(x.size()) : g;
        }

        const DeviceRegressor* device_of(const Regressor& r)
        {
            const auto* d = dynamic_cast<const DeviceRegressor*>(&r);
            return (d && d->HasModel()) ? d : nullptr;
        }

        void device_acq(const DeviceRegressor& r, const double* Xq, long M, AcquisitionFuncType type, double beta, double* val, double* grad)
        {
            slsgp_ctx* const            c = r.Device();
            std::lock_guard<std::mutex> lock(r.DeviceMutex());
            check(c, slsgp_acq_batch(c, internal::to_abi(type), beta, Xq, M, val, grad), "slsgp_acq_batch");
        }

        // maximise `value_and_gradient` over [0, 1]^D from x0 within `max_evals` evaluations
        VectorXd polish(const std::function<double(const VectorXd&, VectorXd&)>& value_and_gradient, const VectorXd& x0, unsigned max_evals)
        {
            const size_t              n = (size_t) x0.size();
            const std::vector<double> lo(n, 0.0), hi(n, 1.0);
            std::vector<double>       start(x0.data(), x0.data() + n);
            const internal::Objective neg = [&](const std::vector<double>& x, std::vector<double>& g) {
                VectorXd xe = VectorXd::Zero((long) n), ge = VectorXd::Zero((long) n);
                for (size_t i = 0; i < n; ++i) xe((long) i) = x[i];
                const double v = value_and_gradient(xe, ge);
                if (!std::isfinite(v)) return std::numeric_limits<double>::infinity();
                for (size_t i = 0; i < n; ++i) g[i] = std::isfinite(ge((long) i)) ? -ge((long) i) : 0.0;
                return -v;
            };
            const internal::MinimizeResult r = internal::minimize_bounded(neg, start, lo, hi, std::max(2u, max_evals), 1e-7, 1e-12);
            VectorXd                       out = VectorXd::Zero((long) n);
            for (size_t i = 0; i < n; ++i) out((long) i) = r.x[i];
            return out;
        }

        // FindGlobalSolution exactly as the reference's default build runs it (src/acquisition-function.cpp:112-167): a starting
        // point from Eigen's Random() (libc rand()), GN_DIRECT with `n_global` evaluations, then LD_LBFGS with `n_local` from
        // DIRECT's answer, both maximising over [0, 1]^D. SearchDriver::Reference only.
        VectorXd find_global_solution_reference(const internal::NloptObjective& objective, unsigned D, unsigned n_global, unsigned n_local)
        {
            const VectorXd upper = VectorXd::Constant(D, 1.0), lower = VectorXd::Constant(D, 0.0);
            const VectorXd x_ini = 0.5 * (VectorXd::Random(D) + VectorXd::Ones(D));
            const VectorXd x_global = internal::nlopt_solve(x_ini, upper, lower, objective, internal::NloptAlgorithm::GN_DIRECT, true, (int) n_global);
            return internal::nlopt_solve(x_global, upper, lower, objective, internal::NloptAlgorithm::LD_LBFGS, true, (int) n_local);
        }

        // Sets the sweep arithmetic of a regressor's context for the lifetime of the guard and puts the previous mode back on
        // every exit path (exceptions included), so a failed search never leaves a user's regressor answering in 1e-3-class
        // arithmetic. The regressor must not serve other threads while a search runs on it.
        class SweepModeGuard
        {
        public:
            SweepModeGuard(const DeviceRegressor& r, slsgp_sweep_mode mode) : m_r(r), m_ctx(r.Device())
            {
                std::lock_guard<std::mutex> lock(m_r.DeviceMutex());
                check(m_ctx, slsgp_get_sweep_mode(m_ctx, &m_previous), "slsgp_get_sweep_mode");
                check(m_ctx, slsgp_set_sweep_mode(m_ctx, mode), "slsgp_set_sweep_mode");
            }
            ~SweepModeGuard()
            {
                std::lock_guard<std::mutex> lock(m_r.DeviceMutex());
                slsgp_set_sweep_mode(m_ctx, m_previous);
            }
            SweepModeGuard(const SweepModeGuard&)            = delete;
            SweepModeGuard& operator=(const SweepModeGuard&) = delete;

        private:
            const DeviceRegressor& m_r;
            slsgp_ctx*             m_ctx;
            slsgp_sweep_mode       m_previous = SLSGP_SWEEP_FP64;
        };

        uint64_t search_seed(const Regressor& r) { return 0x9E3779B97F4A7C15ull ^ ((uint64_t) r.GetLargeX().cols() << 20) ^ (uint64_t) r.GetNumDims(); }
    } // namespace

    namespace acquisition_func
    {
        double CalcAcquisitionValue(const Regressor& regressor, const VectorXd& x, const AcquisitionFuncType func_type, const double hyperparam)
        {
            if (regressor.GetSmallY().rows() == 0) return 0.0; // :176-179
            if (const DeviceRegressor* d = device_of(regressor))
            {
                double v = 0.0;
                device_acq(*d, x.data(), 1, func_type, hyperparam, &v, nullptr);
                return v;
            }
            return generic_value(regressor, x, func_type, hyperparam);
        }

        VectorXd CalcAcquisitionValueDerivative(const Regressor& regressor, const VectorXd& x, const AcquisitionFuncType func_type,
                                                const double hyperparam)
        {
            if (regressor.GetSmallY().rows() == 0) return VectorXd::Zero(x.size()); // :206-209
            if (const DeviceRegressor* d = device_of(regressor))
            {
                VectorXd g = VectorXd::Zero(x.size());
                device_acq(*d, x.data(), 1, func_type, hyperparam, nullptr, g.data());
                return g;
            }
            return generic_derivative(regressor, x, func_type, hyperparam);
        }

        VectorXd CalcAcquisitionValues(const DeviceRegressor& regressor, const MatrixXd& Xq, const AcquisitionFuncType func_type,
                                       const double hyperparam, MatrixXd* derivatives)
        {
            const long M = Xq.cols();
            if (regressor.GetSmallY().rows() == 0 || M == 0) // the reference's "no data" answers (:176-179, :206-209)
            {
                if (derivatives) *derivatives = MatrixXd::Zero(Xq.rows(), M);
                return VectorXd::Zero(M);
            }
            if (!regressor.HasModel()) throw std::logic_error("the regressor holds no model");
            // results are written by the device path: no zero fill (136 MB of page touches per 2^20 candidates at D = 16), and a
            // gradient matrix of the right shape handed in by the caller is reused as it is
            VectorXd val(M);
            if (derivatives) derivatives->resize(Xq.rows(), M);
            device_acq(regressor, Xq.data(), M, func_type, hyperparam, val.data(), derivatives ? derivatives->data() : nullptr);
            return val;
        }

        VectorXd FindNextPoint(const Regressor& regressor, const unsigned num_global_search_iters, const unsigned num_local_search_iters,
                               const AcquisitionFuncType func_type, const double hyperparam)
        {
            const unsigned D = regressor.GetNumDims();
            if (internal::use_nlopt_for_search())
            {
                const DeviceRegressor*         d         = device_of(regressor);
                const internal::NloptObjective objective = [&](const std::vector<double>& x, std::vector<double>& grad) {
                    const VectorXd xe = Eigen::Map<const VectorXd>(x.data(), (long) x.size());
                    if (d && regressor.GetSmallY().rows() != 0) // value and gradient from ONE one-candidate sweep
                    {
                        double v = 0.0;
                        device_acq(*d, xe.data(), 1, func_type, hyperparam, &v, grad.empty() ? nullptr : grad.data());
                        return v;
                    }
                    if (!grad.empty())
                    {
                        const VectorXd g = CalcAcquisitionValueDerivative(regressor, xe, func_type, hyperparam);
                        for (size_t i = 0; i < grad.size(); ++i) grad[i] = g((long) i);
                    }
                    return CalcAcquisitionValue(regressor, xe, func_type, hyperparam);
                };
                return find_global_solution_reference(objective, D, num_global_search_iters, num_local_search_iters);
            }
            if (regressor.GetSmallY().rows() == 0 || D == 0) return VectorXd::Constant(D, 0.5); // flat objective: the box centre
            const long     count = (long) std::max(1u, num_global_search_iters) * kCandidatesPerGlobalIter;
            VectorXd       x0    = VectorXd::Zero(D);
            if (const DeviceRegressor* d = device_of(regressor))
            {
                // the tensor-core sweep pays off for large candidate counts; the ascent and the polish are FP64
                const bool tensor = count >= 32768 && D <= (regressor.GetKernelType() == KernelType::ArdSquaredExponentialKernel ? 67u : 66u);
                const SweepModeGuard        mode(*d, tensor ? SLSGP_SWEEP_TENSOR : SLSGP_SWEEP_FP64);
                slsgp_ctx* const            c = d->Device();
                std::lock_guard<std::mutex> lock(d->DeviceMutex());
                // global sweep + batched multi-start ascent, all on the device; the winner is polished below
                double    v0       = 0.0;
                const int n_starts = (int) std::min<long>(1024, std::max<long>(1, count / 64));
                check(c,
                      slsgp_acq_maximize(c, internal::to_abi(func_type), hyperparam, search_seed(regressor), 0, count, n_starts,
                                         (int) std::min(num_local_search_iters, 200u), x0.data(), &v0, nullptr, nullptr),
                      "slsgp_acq_maximize");
            }
            else
            {
                // foreign regressor: the same candidate sequence, evaluated through the virtual interface
                double best = -std::numeric_limits<double>::infinity();
                for (long i = 0; i < std::min<long>(count, 4096); ++i)
                {
                    VectorXd x = VectorXd::Zero(D);
                    for (unsigned k = 0; k < D; ++k) x(k) = std::fmod(0.5 + (i + 1) * 0.6180339887498949 * (k + 1), 1.0);
                    const double v = generic_value(regressor, x, func_type, hyperparam);
                    if (v > best) best = v, x0 = x;
                }
            }
            const DeviceRegressor* dev = device_of(regressor);
            return polish(
                [&](const VectorXd& x, VectorXd& g) {
                    if (dev) // value and gradient from ONE sweep call
                    {
                        double v = 0.0;
                        g        = VectorXd::Zero(x.size());
                        device_acq(*dev, x.data(), 1, func_type, hyperparam, &v, g.data());
                        return v;
                    }
                    g = CalcAcquisitionValueDerivative(regressor, x, func_type, hyperparam);
                    return CalcAcquisitionValue(regressor, x, func_type, hyperparam);
                },
                x0, num_local_search_iters);
        }

        std::vector<VectorXd> FindNextPoints(const Regressor& regressor, const unsigned num_points, const unsigned num_global_search_iters,
                                             const unsigned num_local_search_iters, const AcquisitionFuncType func_type, const double hyperparam)
        {
            const unsigned        D = regressor.GetNumDims();
            std::vector<VectorXd> points;
            if (num_points == 0) return points;
            const DeviceRegressor* orig = device_of(regressor);
            if (!orig) throw std::invalid_argument("FindNextPoints needs a device-backed regressor with data");
            slsgp_ctx* const orig_ctx = orig->Device(); // (a fresh copy builds its device model here, outside any lock)

            const VectorXd theta = regressor.GetKernelHyperparams();
            const double   noise = regressor.GetNoiseHyperparam();
            // As in the reference (:259-260) the temporary regressor is built with GaussianProcessRegressor's DEFAULT
            // kernel type (Matern 5/2) whatever kernel `regressor` uses; kept for parity.
            // It lives on the device only and grows by a bordered O(N^2) update per pending point
            // where the reference constructs, and inverts, a new regressor per option.
            std::unique_ptr<GaussianProcessRegressor> temp(new GaussianProcessRegressor(regressor.GetLargeX(), regressor.GetSmallY(), theta, noise,
                                                                                        KernelType::ArdMatern52Kernel,
                                                                                        GaussianProcessRegressor::DeviceOnly()));

            double f_best = 0.0;
            {
                std::lock_guard<std::mutex> lock(orig->DeviceMutex());
                check(orig_ctx, slsgp_get_f_best(orig_ctx, &f_best, nullptr), "slsgp_get_f_best");
            }
            // mu / dmu from the original model, sigma / dsigma from the temporary one, formulas on the device
            const auto pair_acq = [&](const MatrixXd& Xq, VectorXd& val, MatrixXd* grad) {
                const long M = Xq.cols();
                VectorXd   mu, sigma;
                MatrixXd   dmu, dsigma;
                orig->PredictBatch(Xq, &mu, nullptr, grad ? &dmu : nullptr, nullptr);
                temp->PredictBatch(Xq, nullptr, &sigma, nullptr, grad ? &dsigma : nullptr);
                val = VectorXd::Zero(M);
                if (grad) *grad = MatrixXd::Zero(D, M);
                std::lock_guard<std::mutex> lock(orig->DeviceMutex());
                check(orig_ctx,
                      slsgp_acq_from_posterior(orig_ctx, internal::to_abi(func_type), hyperparam, f_best, (int) D, M, mu.data(), sigma.data(),
                                               grad ? dmu.data() : nullptr, grad ? dsigma.data() : nullptr, val.data(), grad ? grad->data() : nullptr),
                      "slsgp_acq_from_posterior");
            };

            if (internal::use_nlopt_for_search())
            {
                // the reference's loop (:264-296): DIRECT + L-BFGS on Schonlau's criterion, one-candidate device calls per evaluation
                for (unsigned i = 0; i < num_points; ++i)
                {
                    const internal::NloptObjective objective = [&](const std::vector<double>& x, std::vector<double>& grad) {
                        MatrixXd X1 = MatrixXd::Zero(D, 1), G;
                        for (unsigned k = 0; k < D; ++k) X1(k, 0) = x[k];
                        VectorXd v;
                        pair_acq(X1, v, grad.empty() ? nullptr : &G);
                        for (size_t k = 0; k < grad.size(); ++k) grad[k] = G((long) k, 0);
                        return v(0);
                    };
                    const VectorXd x_star = find_global_solution_reference(objective, D, num_global_search_iters, num_local_search_iters);
                    points.push_back(x_star);
                    if (points.size() != num_points) temp->AppendPoint(x_star, temp->PredictMu(x_star));
                }
                return points;
            }

            // global stage: the counter-based candidate sequence, mu from the original model and sigma from the temporary one, all
            // on the device (slsgp_pair_acq_argmax); tensor sweep for large counts
            const long count  = std::min<long>((long) std::max(1u, num_global_search_iters) * kCandidatesPerGlobalIter, 1L << 21);
            const bool tensor = count >= 32768 && D <= 66;
            for (unsigned i = 0; i < num_points; ++i)
            {
                VectorXd       x_best = VectorXd::Constant(D, 0.5);
                const uint64_t seed   = search_seed(regressor) + 0x51ED27ull * (i + 1);
                {
                    const SweepModeGuard mode_orig(*orig, tensor ? SLSGP_SWEEP_TENSOR : SLSGP_SWEEP_FP64), mode_temp(*temp, tensor ? SLSGP_SWEEP_TENSOR : SLSGP_SWEEP_FP64);
                    slsgp_ctx* const     temp_ctx = temp->Device();
                    std::lock_guard<std::mutex> lock_orig(orig->DeviceMutex());
                    std::lock_guard<std::mutex> lock_temp(temp->DeviceMutex());
                    double v_best = 0.0;
                    check(orig_ctx, slsgp_pair_acq_argmax(orig_ctx, temp_ctx, internal::to_abi(func_type), hyperparam, seed, 0, count, x_best.data(), &v_best, nullptr),
                          "slsgp_pair_acq_argmax");
                } // the sweep modes are restored here: the polish below is IEEE double
                const VectorXd x_star = polish(
                    [&](const VectorXd& x, VectorXd& g) {
                        MatrixXd X1 = MatrixXd::Zero(D, 1), G;
                        X1.col(0)   = x;
                        VectorXd v;
                        pair_acq(X1, v, &G);
                        g = G.col(0);
                        return v(0);
                    },
                    x_best, num_local_search_iters);
                points.push_back(x_star);

                if (points.size() != num_points)
                {
                    // the pending point joins the temporary model with its predicted value (which never influences sigma)
                    temp->AppendPoint(x_star, temp->PredictMu(x_star));
                }
            }
            return points;
        }
    } // namespace acquisition_func
} // namespace sequential_line_search
