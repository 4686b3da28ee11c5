// SequentialLineSearchOptimizer / PreferentialBayesianOptimizer and their helpers (reference: src/sequential-line-search.cpp,
// src/preferential-bayesian-optimizer.cpp, src/preference-data-manager.cpp, src/slider.cpp). Pure host bookkeeping around
// the two device-backed steps: the PreferenceRegressor MAP fit and acquisition_func::FindNextPoint(s).
#include "device.hpp"
#include "nlopt_driver.hpp"

#include <sequential-line-search/optimizers.hpp>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

using Eigen::MatrixXd;
using Eigen::VectorXd;

namespace sequential_line_search
{
    namespace
    {
        VectorXd random_point(const int n) // uniform in [0, 1]^n from Eigen's Random(), as the reference's generators
        {
            const VectorXd r = VectorXd::Random(n);
            VectorXd       x = VectorXd::Zero(n);
            for (int i = 0; i < n; ++i) x(i) = 0.5 * (r(i) + 1.0);
            return x;
        }

        // ---- slider enlargement (src/slider.cpp:20-128) -------------------------------------------------------
        const double kBoxLo = 1e-16, kBoxHi = 1.0 - 1e-16; // the reference's crop() interval

        double   clamp_to_box(double v) { return v > kBoxLo ? (v < kBoxHi ? v : kBoxHi) : kBoxLo; }
        VectorXd clamp_to_box(const VectorXd& x)
        {
            VectorXd y = VectorXd::Zero(x.size());
            for (int i = 0; i < (int) x.size(); ++i) y(i) = clamp_to_box(x(i));
            return y;
        }
        // The largest t in [0, t_max] with c + t * dir inside the box: what the reference obtains from COBYLA
        // (maximise -(t - scale)^2 subject to "no coordinate is cropped").
        double largest_feasible_step(const VectorXd& c, const VectorXd& dir, double t_max)
        {
            double t = t_max;
            for (int i = 0; i < (int) c.size(); ++i)
            {
                if (dir(i) > 0.0) t = std::min(t, (kBoxHi - c(i)) / dir(i));
                if (dir(i) < 0.0) t = std::min(t, (kBoxLo - c(i)) / dir(i));
            }
            return std::max(t, 0.0);
        }
        std::pair<VectorXd, VectorXd> enlarge(const VectorXd& x_1, const VectorXd& x_2, double scale, double minimum_length)
        {
            const VectorXd a = clamp_to_box(x_1), b = clamp_to_box(x_2);
            const VectorXd c = 0.5 * (a + b);
            const VectorXd r = a - c;
            VectorXd       minus_r = VectorXd::Zero(r.size());
            for (int i = 0; i < (int) r.size(); ++i) minus_r(i) = -r(i);
            const double   t_1 = largest_feasible_step(c, r, scale), t_2 = largest_feasible_step(c, minus_r, scale);
            const VectorXd e_1 = clamp_to_box(c + t_1 * r), e_2 = clamp_to_box(c - t_2 * r);
            const double   length = (e_1 - e_2).norm();
            if (length < minimum_length) // a very short slider is stretched to the minimum length (:112-126)
            {
                const double k = minimum_length / length;
                if (std::abs(t_1 - t_2) < 1e-10) return {c + k * t_1 * r, c - k * t_2 * r};
                if (t_1 > t_2) return {c + 2.0 * k * t_1 * r, c - t_2 * r};
                return {c + t_1 * r, c - 2.0 * k * t_2 * r};
            }
            return {e_1, e_2};
        }

        // SearchDriver::Reference: the enlargement as the reference solves it (src/slider.cpp:73-122): two 1-D COBYLA runs from
        // t = 1 over [0, scale] on -(t - scale)^2 subject to "c +- t r needs no cropping" (tolerance 1e-10). The reference passes
        // 1000 in the position of `is_maximization` (so: maximise, default budget of 1000 evaluations); kept.
        std::pair<VectorXd, VectorXd> enlarge_reference(const VectorXd& x_1, const VectorXd& x_2, double scale, double minimum_length)
        {
            const VectorXd c = 0.5 * (clamp_to_box(x_1) + clamp_to_box(x_2));
            const VectorXd r = clamp_to_box(x_1) - c;
            const internal::NloptObjective objective = [&](const std::vector<double>& x, std::vector<double>&) { return -(x[0] - scale) * (x[0] - scale); };
            const auto constraint_of = [&](double sign) {
                return internal::NloptObjective([&c, &r, sign](const std::vector<double>& x, std::vector<double>&) {
                    const VectorXd y   = c + (sign * x[0]) * r;
                    double         sum = 0.0;
                    for (int i = 0; i < (int) y.size(); ++i) sum += (clamp_to_box(y(i)) - y(i)) * (clamp_to_box(y(i)) - y(i));
                    return sum;
                });
            };
            const internal::NloptObjective constraint_p = constraint_of(+1.0), constraint_n = constraint_of(-1.0);
            const VectorXd x0 = VectorXd::Constant(1, 1.0), upper = VectorXd::Constant(1, scale), lower = VectorXd::Constant(1, 0.0);
            const double   t_1 = internal::nlopt_solve(x0, upper, lower, objective, internal::NloptAlgorithm::LN_COBYLA, true, 1000, &constraint_p)(0);
            const double   t_2 = internal::nlopt_solve(x0, upper, lower, objective, internal::NloptAlgorithm::LN_COBYLA, true, 1000, &constraint_n)(0);
            const VectorXd e_1 = clamp_to_box(c + t_1 * r), e_2 = clamp_to_box(c - t_2 * r);
            const double   length = (e_1 - e_2).norm();
            if (length < minimum_length)
            {
                const double k = minimum_length / length;
                if (std::abs(t_1 - t_2) < 1e-10) return {c + k * t_1 * r, c - k * t_2 * r};
                if (t_1 > t_2) return {c + 2.0 * k * t_1 * r, c - t_2 * r};
                return {c + t_1 * r, c - 2.0 * k * t_2 * r};
            }
            return {e_1, e_2};
        }

        // ---- merging of (nearly) coincident data points (src/preference-data-manager.cpp:14-86) ------------
        bool merge_first_close_pair(double eps_squared, MatrixXd& X, std::vector<Preference>& D)
        {
            const int M = (int) X.cols(), dims = (int) X.rows();
            for (int i = 0; i < M; ++i)
                for (int j = i + 1; j < M; ++j)
                {
                    if (!((X.col(i) - X.col(j)).squaredNorm() < eps_squared)) continue;
                    // survivors keep their order; the merged point (the midpoint) becomes the last column
                    std::vector<unsigned> to_new((size_t) M);
                    unsigned              next = 0;
                    for (int old = 0; old < M; ++old) to_new[(size_t) old] = (old == i || old == j) ? (unsigned) (M - 2) : next++;
                    MatrixXd merged = MatrixXd::Zero(dims, M - 1);
                    for (int old = 0; old < M; ++old)
                        if (old != i && old != j) merged.col(to_new[(size_t) old]) = X.col(old);
                    merged.col(M - 2) = 0.5 * (X.col(i) + X.col(j));
                    X                 = merged;
                    for (Preference& p : D)
                        for (unsigned& index : p) index = to_new[index];
                    return true;
                }
            return false;
        }

        // state of the previous regressor as the starting point of the next MAP fit
        std::unique_ptr<MapWarmStart> warm_start_of(const std::shared_ptr<PreferenceRegressor>& r)
        {
            static const bool disabled = std::getenv("SLS_B200_NO_WARM_START") != nullptr; // A/B switch for timing studies
            if (disabled || !r || r->GetSmallY().size() == 0) return nullptr;
            std::unique_ptr<MapWarmStart> w(new MapWarmStart);
            w->X = r->GetLargeX(), w->y = r->GetSmallY(), w->kernel_hyperparams = r->GetKernelHyperparams(), w->noise_hyperparam = r->GetNoiseHyperparam();
            // incremental refit (driver.hpp): the old regressor is about to be replaced; its factored model moves on
            if (GetIncrementalRefit() && !r->m_use_map_hyperparams) w->device = r->HandOverDevice();
            return w;
        }
    } // namespace

    // ------------------------------------------------------------------------------------------------------------
    Slider::Slider(const VectorXd& end_0_in, const VectorXd& end_1_in, const bool enlarge_it, const double scale, const double minimum_length)
        : end_0(end_0_in), end_1(end_1_in), original_end_0(end_0_in), original_end_1(end_1_in)
    {
        if (!enlarge_it) return;
        const auto ends = internal::use_nlopt_for_search() ? enlarge_reference(original_end_0, original_end_1, scale, minimum_length)
                                                            : enlarge(original_end_0, original_end_1, scale, minimum_length);
        end_0 = ends.first, end_1 = ends.second;
    }

    void PreferenceDataManager::AddNewPoints(const VectorXd& x_preferable, const std::vector<VectorXd>& xs_other, const bool merge_close_points,
                                             const double epsilon)
    {
        const int  d = (int) x_preferable.size(), n_old = (int) m_X.cols(), n_new = (int) xs_other.size() + 1;
        const bool first_batch = m_X.rows() == 0;
        MatrixXd   X = MatrixXd::Zero(d, n_old + n_new);
        for (int j = 0; j < n_old; ++j) X.col(j) = m_X.col(j);
        X.col(n_old) = x_preferable;
        for (int i = 0; i + 1 < n_new; ++i) X.col(n_old + 1 + i) = xs_other[(size_t) i];
        m_X = X;
        std::vector<unsigned> tuple((size_t) n_new);
        for (int i = 0; i < n_new; ++i) tuple[(size_t) i] = (unsigned) (n_old + i);
        m_D.push_back(Preference(tuple));
        if (first_batch || !merge_close_points) return; // the reference does not merge inside the very first batch
        while (merge_first_close_pair(epsilon * epsilon, m_X, m_D)) {}
    }

    std::pair<VectorXd, VectorXd> GenerateRandomSliderEnds(const int num_dims) { return {random_point(num_dims), random_point(num_dims)}; }

    std::pair<VectorXd, VectorXd> GenerateCenteredFixedLengthRandomSliderEnds(const int num_dims)
    {
        const VectorXd centre = VectorXd::Constant(num_dims, 0.5), dir = 0.5 * VectorXd::Random(num_dims);
        return {centre + dir, centre - dir};
    }

    std::vector<VectorXd> GenerateRandomPoints(const int num_dims, const int num_options)
    {
        std::vector<VectorXd> options;
        for (int i = 0; i < num_options; ++i) options.push_back(random_point(num_dims));
        return options;
    }

    // ------------------------------------------------------------------------------------------------------------
    // SequentialLineSearchOptimizer
    // ------------------------------------------------------------------------------------------------------------
    SequentialLineSearchOptimizer::SequentialLineSearchOptimizer(const int num_dims, const bool use_slider_enlargement, const bool use_map_hyperparams,
                                                                 const KernelType kernel_type, const AcquisitionFuncType acquisition_func_type,
                                                                 const std::function<std::pair<VectorXd, VectorXd>(const int)>& initial_query_generator,
                                                                 const CurrentBestSelectionStrategy current_best_selection_strategy)
        : m_use_slider_enlargement(use_slider_enlargement),
          m_use_map_hyperparams(use_map_hyperparams),
          m_current_best_selection_strategy(current_best_selection_strategy),
          m_kernel_signal_var(0.500),
          m_kernel_length_scale(0.500),
          m_noise_level(0.005),
          m_kernel_hyperparams_prior_var(0.250),
          m_btl_scale(0.010),
          m_kernel_type(kernel_type),
          m_acquisition_func_type(acquisition_func_type),
          m_gaussian_process_upper_confidence_bound_hyperparam(1.0)
    {
        const auto ends = initial_query_generator(num_dims);
        m_data          = std::make_shared<PreferenceDataManager>();
        m_slider        = std::make_shared<Slider>(ends.first, ends.second, false); // the first slider is never enlarged
    }

    void SequentialLineSearchOptimizer::SetHyperparams(const double kernel_signal_var, const double kernel_length_scale, const double noise_level,
                                                       const double kernel_hyperparams_prior_var, const double btl_scale)
    {
        m_kernel_signal_var = kernel_signal_var, m_kernel_length_scale = kernel_length_scale, m_noise_level = noise_level;
        m_kernel_hyperparams_prior_var = kernel_hyperparams_prior_var, m_btl_scale = btl_scale;
    }

    void SequentialLineSearchOptimizer::SubmitFeedbackData(const double slider_position)
    {
        // the reference's effort heuristic (src/sequential-line-search.cpp:67-80): 100 / 50 D / 10 D
        const int num_dims = (int) GetMaximizer().size();
        SubmitFeedbackData(slider_position, 100, 50 * num_dims, 10 * num_dims);
    }

    void SequentialLineSearchOptimizer::SubmitFeedbackData(const double slider_position, const int num_map_estimation_iters,
                                                           const int num_global_search_iters, const int num_local_search_iters)
    {
        const VectorXd x_chosen = CalcPointFromSliderPosition(slider_position);
        m_data->AddNewPoints(x_chosen, {m_slider->original_end_0, m_slider->original_end_1}, true);

        const auto now = [] { return std::chrono::steady_clock::now(); };
        const auto ms  = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        const auto t0 = now();
        const std::unique_ptr<MapWarmStart> warm = warm_start_of(m_regressor);
        m_regressor = std::make_shared<PreferenceRegressor>(m_data->GetX(), m_data->GetD(), m_use_map_hyperparams, m_kernel_signal_var,
                                                            m_kernel_length_scale, m_noise_level, m_kernel_hyperparams_prior_var, m_btl_scale,
                                                            (unsigned) num_map_estimation_iters, m_kernel_type, warm.get());
        const auto t1 = now();

        const VectorXd x_plus = m_current_best_selection_strategy == CurrentBestSelectionStrategy::LargestExpectValue ? m_regressor->FindArgMax() : x_chosen;
        const VectorXd x_acquisition =
            acquisition_func::FindNextPoint(*m_regressor, (unsigned) num_global_search_iters, (unsigned) num_local_search_iters, m_acquisition_func_type,
                                            m_gaussian_process_upper_confidence_bound_hyperparam);
        const auto t2 = now();
        m_slider = std::make_shared<Slider>(x_plus, x_acquisition, m_use_slider_enlargement);
        m_last_timings.map_fit = ms(t0, t1), m_last_timings.search = ms(t1, t2), m_last_timings.slider = ms(t2, now());
    }

    std::pair<VectorXd, VectorXd> SequentialLineSearchOptimizer::GetSliderEnds() const { return {m_slider->end_0, m_slider->end_1}; }
    VectorXd SequentialLineSearchOptimizer::CalcPointFromSliderPosition(const double slider_position) const { return m_slider->GetValue(slider_position); }
    VectorXd SequentialLineSearchOptimizer::GetMaximizer() const { return m_slider->original_end_0; }

    double SequentialLineSearchOptimizer::GetPreferenceValueMean(const VectorXd& point) const { return m_regressor ? m_regressor->PredictMu(point) : 0.0; }
    double SequentialLineSearchOptimizer::GetPreferenceValueStdev(const VectorXd& point) const { return m_regressor ? m_regressor->PredictSigma(point) : 0.0; }
    double SequentialLineSearchOptimizer::GetAcquisitionFuncValue(const VectorXd& point) const
    {
        return m_regressor ? acquisition_func::CalcAcquisitionValue(*m_regressor, point, m_acquisition_func_type, m_gaussian_process_upper_confidence_bound_hyperparam)
                           : 0.0;
    }
    const MatrixXd& SequentialLineSearchOptimizer::GetRawDataPoints() const { return m_data->GetX(); }
    void            SequentialLineSearchOptimizer::DampData(const std::string& directory_path) const
    {
        if (m_regressor) m_regressor->DampData(directory_path);
    }

    // ------------------------------------------------------------------------------------------------------------
    // PreferentialBayesianOptimizer
    // ------------------------------------------------------------------------------------------------------------
    PreferentialBayesianOptimizer::PreferentialBayesianOptimizer(const int num_dims, const bool use_map_hyperparams, const KernelType kernel_type,
                                                                 const AcquisitionFuncType acquisition_func_type,
                                                                 const InitialQueryGenerator& initial_query_generator,
                                                                 const CurrentBestSelectionStrategy current_best_selection_strategy, const int num_options)
        : m_use_map_hyperparams(use_map_hyperparams),
          m_num_options(num_options),
          m_current_best_selection_strategy(current_best_selection_strategy),
          m_kernel_signal_var(0.500),
          m_kernel_length_scale(0.500),
          m_noise_level(0.005),
          m_kernel_hyperparams_prior_var(0.250),
          m_btl_scale(0.010),
          m_kernel_type(kernel_type),
          m_acquisition_func_type(acquisition_func_type),
          m_gaussian_process_upper_confidence_bound_hyperparam(1.0)
    {
        m_data            = std::make_shared<PreferenceDataManager>();
        m_current_options = initial_query_generator(num_dims, num_options);
        if ((int) m_current_options.size() != m_num_options) throw std::invalid_argument("the initial query generator must return num_options points");
    }

    void PreferentialBayesianOptimizer::SetHyperparams(const double kernel_signal_var, const double kernel_length_scale, const double noise_level,
                                                       const double kernel_hyperparams_prior_var, const double btl_scale)
    {
        m_kernel_signal_var = kernel_signal_var, m_kernel_length_scale = kernel_length_scale, m_noise_level = noise_level;
        m_kernel_hyperparams_prior_var = kernel_hyperparams_prior_var, m_btl_scale = btl_scale;
    }

    void PreferentialBayesianOptimizer::SubmitFeedbackData(const int option_index, const int num_map_estimation_iters)
    {
        if (option_index < 0 || option_index >= (int) m_current_options.size()) throw std::out_of_range("option_index");
        std::vector<VectorXd> others = m_current_options;
        others.erase(others.begin() + option_index);
        m_data->AddNewPoints(m_current_options[(size_t) option_index], others, true);
        PerformMapEstimation(num_map_estimation_iters);
    }

    void PreferentialBayesianOptimizer::SubmitCustomFeedbackData(const VectorXd& chosen_option, const std::vector<VectorXd>& other_options,
                                                                 const int num_map_estimation_iters)
    {
        m_data->AddNewPoints(chosen_option, other_options, true);
        PerformMapEstimation(num_map_estimation_iters);
    }

    void PreferentialBayesianOptimizer::DetermineNextQuery(const int num_global_search_iters_in, const int num_local_search_iters_in)
    {
        if (!m_regressor) throw std::logic_error("DetermineNextQuery needs feedback data first");
        // effort heuristic of the reference (src/preferential-bayesian-optimizer.cpp:95-103): 50 D^2 / 10 D
        const int num_dims = (int) GetMaximizer().size();
        const int n_global = num_global_search_iters_in > 0 ? num_global_search_iters_in : 50 * num_dims * num_dims;
        const int n_local  = num_local_search_iters_in > 0 ? num_local_search_iters_in : 10 * num_dims;

        const VectorXd x_plus =
            m_current_best_selection_strategy == CurrentBestSelectionStrategy::LargestExpectValue ? m_regressor->FindArgMax() : VectorXd(m_data->GetLastSelectedDataPoint());
        const std::vector<VectorXd> next = acquisition_func::FindNextPoints(*m_regressor, (unsigned) (m_num_options - 1), (unsigned) n_global, (unsigned) n_local,
                                                                            m_acquisition_func_type, m_gaussian_process_upper_confidence_bound_hyperparam);
        m_current_options[0] = x_plus; // the first option is always the current best
        for (int i = 1; i < m_num_options; ++i) m_current_options[(size_t) i] = next[(size_t) i - 1];
    }

    VectorXd PreferentialBayesianOptimizer::GetMaximizer() const { return m_current_options[0]; }
    double   PreferentialBayesianOptimizer::GetPreferenceValueMean(const VectorXd& point) const { return m_regressor ? m_regressor->PredictMu(point) : 0.0; }
    double   PreferentialBayesianOptimizer::GetPreferenceValueStdev(const VectorXd& point) const { return m_regressor ? m_regressor->PredictSigma(point) : 0.0; }
    double   PreferentialBayesianOptimizer::GetAcquisitionFuncValue(const VectorXd& point) const
    {
        return m_regressor ? acquisition_func::CalcAcquisitionValue(*m_regressor, point, m_acquisition_func_type, m_gaussian_process_upper_confidence_bound_hyperparam)
                           : 0.0;
    }
    const MatrixXd& PreferentialBayesianOptimizer::GetRawDataPoints() const { return m_data->GetX(); }
    void            PreferentialBayesianOptimizer::DampData(const std::string& directory_path) const
    {
        if (m_regressor) m_regressor->DampData(directory_path);
    }

    void PreferentialBayesianOptimizer::PerformMapEstimation(const int num_map_estimation_iters_in)
    {
        // default budget of the reference (:187-195): 10 (D + N)
        const int iters = num_map_estimation_iters_in > 0 ? num_map_estimation_iters_in : 10 * ((int) GetMaximizer().size() + m_data->GetNumDataPoints());
        const std::unique_ptr<MapWarmStart> warm = warm_start_of(m_regressor);
        m_regressor     = std::make_shared<PreferenceRegressor>(m_data->GetX(), m_data->GetD(), m_use_map_hyperparams, m_kernel_signal_var, m_kernel_length_scale,
                                                                m_noise_level, m_kernel_hyperparams_prior_var, m_btl_scale, (unsigned) iters, m_kernel_type, warm.get());
    }
} // namespace sequential_line_search
