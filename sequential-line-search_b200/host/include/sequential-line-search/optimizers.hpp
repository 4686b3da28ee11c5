// The two public optimisers of the reference and their helpers, on the device-backed regressors of regressors.hpp:
//   SequentialLineSearchOptimizer      include/sequential-line-search/sequential-line-search.hpp:17-125
//   PreferentialBayesianOptimizer      include/sequential-line-search/preferential-bayesian-optimizer.hpp:21-156
//   Slider                             slider.hpp:8-37          PreferenceDataManager   preference-data-manager.hpp:10-47
//   CurrentBestSelectionStrategy       current-best-selection-strategy.hpp:7-12
// Same constructor arguments, defaults, method names and bookkeeping; the MAP fit and the acquisition search inside
// SubmitFeedbackData / DetermineNextQuery run on the GPU. The slider enlargement, which the reference solves with two
// COBYLA runs (src/slider.cpp:74-128), is solved in closed form here (the largest step that keeps the end inside the box).
#ifndef SEQUENTIAL_LINE_SEARCH_B200_OPTIMIZERS_HPP
#define SEQUENTIAL_LINE_SEARCH_B200_OPTIMIZERS_HPP

#include <Eigen/Core>
#include <functional>
#include <memory>
#include <sequential-line-search/acquisition.hpp>
#include <sequential-line-search/regressors.hpp>
#include <string>
#include <utility>
#include <vector>

namespace sequential_line_search
{
    enum class CurrentBestSelectionStrategy
    {
        LargestExpectValue, // x^+ : the data point with the largest estimated goodness
        LastSelection,      // x^chosen : the point picked in the last query
    };

    class Slider
    {
    public:
        Slider(const Eigen::VectorXd& end_0, const Eigen::VectorXd& end_1, const bool enlarge, const double scale = 1.25,
               const double minimum_length = 0.25);

        Eigen::VectorXd GetValue(const double t) const { return (1.0 - t) * end_0 + t * end_1; }

        Eigen::VectorXd end_0, end_1;                   // what the user drags between
        Eigen::VectorXd original_end_0, original_end_1; // x^+ and x^EI before the enlargement
    };

    class PreferenceDataManager
    {
    public:
        // Appends [x_preferable, xs_other...] as new columns and one tuple with x_preferable first; then merges points
        // closer than epsilon (the pair is replaced by its midpoint, which becomes the LAST column).
        void AddNewPoints(const Eigen::VectorXd& x_preferable, const std::vector<Eigen::VectorXd>& xs_other,
                          const bool merge_close_points = true, const double epsilon = 1e-04);

        const Eigen::VectorXd GetLastSelectedDataPoint() const { return m_X.col(m_D.back()[0]); }
        int                   GetNumDataPoints() const { return m_X.cols(); }
        const Eigen::MatrixXd&         GetX() const { return m_X; }
        const std::vector<Preference>& GetD() const { return m_D; }

    private:
        Eigen::MatrixXd         m_X;
        std::vector<Preference> m_D;
    };

    std::pair<Eigen::VectorXd, Eigen::VectorXd> GenerateRandomSliderEnds(const int num_dims);
    std::pair<Eigen::VectorXd, Eigen::VectorXd> GenerateCenteredFixedLengthRandomSliderEnds(const int num_dims);

    class SequentialLineSearchOptimizer
    {
    public:
        SequentialLineSearchOptimizer(const int num_dims, const bool use_slider_enlargement = true, const bool use_map_hyperparams = true,
                                      const KernelType          kernel_type           = KernelType::ArdMatern52Kernel,
                                      const AcquisitionFuncType acquisition_func_type = AcquisitionFuncType::ExpectedImprovement,
                                      const std::function<std::pair<Eigen::VectorXd, Eigen::VectorXd>(const int)>& initial_query_generator =
                                          GenerateRandomSliderEnds,
                                      const CurrentBestSelectionStrategy current_best_selection_strategy =
                                          CurrentBestSelectionStrategy::LargestExpectValue);

        void SetHyperparams(const double kernel_signal_var = 0.500, const double kernel_length_scale = 0.500, const double noise_level = 0.005,
                            const double kernel_hyperparams_prior_var = 0.250, const double btl_scale = 0.010);

        // The user picked `slider_position` in [0, 1]: add (chosen > x^+, x^EI), refit, search, build the next slider.
        void SubmitFeedbackData(const double slider_position);
        void SubmitFeedbackData(const double slider_position, const int num_map_estimation_iters, const int num_global_search_iters,
                                const int num_local_search_iters);

        std::pair<Eigen::VectorXd, Eigen::VectorXd> GetSliderEnds() const;
        Eigen::VectorXd                             CalcPointFromSliderPosition(const double slider_position) const;
        Eigen::VectorXd                             GetMaximizer() const;

        double GetPreferenceValueMean(const Eigen::VectorXd& point) const;
        double GetPreferenceValueStdev(const Eigen::VectorXd& point) const;
        double GetAcquisitionFuncValue(const Eigen::VectorXd& point) const;

        const Eigen::MatrixXd& GetRawDataPoints() const;
        void                   DampData(const std::string& directory_path) const;

        void SetGaussianProcessUpperConfidenceBoundHyperparam(const double hyperparam) { m_gaussian_process_upper_confidence_bound_hyperparam = hyperparam; }

        // addition: the regressor of the last Submit (null before the first), for batched queries on the device
        std::shared_ptr<const PreferenceRegressor> GetRegressor() const { return m_regressor; }
        // addition: where the last SubmitFeedbackData spent its time, in milliseconds of host wall clock
        struct StepTimings
        {
            double map_fit = 0.0, search = 0.0, slider = 0.0;
        };
        StepTimings GetLastStepTimings() const { return m_last_timings; }

    private:
        StepTimings                            m_last_timings;
        const bool                             m_use_slider_enlargement;
        const bool                             m_use_map_hyperparams;
        const CurrentBestSelectionStrategy     m_current_best_selection_strategy;
        std::shared_ptr<PreferenceRegressor>   m_regressor;
        std::shared_ptr<Slider>                m_slider;
        std::shared_ptr<PreferenceDataManager> m_data;
        double                                 m_kernel_signal_var, m_kernel_length_scale, m_noise_level, m_kernel_hyperparams_prior_var, m_btl_scale;
        const KernelType                       m_kernel_type;
        const AcquisitionFuncType              m_acquisition_func_type;
        double                                 m_gaussian_process_upper_confidence_bound_hyperparam;
    };

    using InitialQueryGenerator = std::function<std::vector<Eigen::VectorXd>(const int, const int)>;
    std::vector<Eigen::VectorXd> GenerateRandomPoints(const int num_dims, const int num_options);

    class PreferentialBayesianOptimizer
    {
    public:
        PreferentialBayesianOptimizer(const int num_dims, const bool use_map_hyperparams = true,
                                      const KernelType                   kernel_type             = KernelType::ArdMatern52Kernel,
                                      const AcquisitionFuncType          acquisition_func_type   = AcquisitionFuncType::ExpectedImprovement,
                                      const InitialQueryGenerator&       initial_query_generator = GenerateRandomPoints,
                                      const CurrentBestSelectionStrategy current_best_selection_strategy =
                                          CurrentBestSelectionStrategy::LargestExpectValue,
                                      const int num_options = 2);

        void SetHyperparams(const double kernel_signal_var = 0.500, const double kernel_length_scale = 0.500, const double noise_level = 0.005,
                            const double kernel_hyperparams_prior_var = 0.250, const double btl_scale = 0.010);

        void SubmitFeedbackData(const int option_index, const int num_map_estimation_iters = 0);
        void SubmitCustomFeedbackData(const Eigen::VectorXd& chosen_option, const std::vector<Eigen::VectorXd>& other_options,
                                      const int num_map_estimation_iters = 0);
        void DetermineNextQuery(const int num_global_search_iters = 0, const int num_local_search_iters = 0);

        const std::vector<Eigen::VectorXd>& GetCurrentOptions() const { return m_current_options; }
        Eigen::VectorXd                     GetMaximizer() const;

        double GetPreferenceValueMean(const Eigen::VectorXd& point) const;
        double GetPreferenceValueStdev(const Eigen::VectorXd& point) const;
        double GetAcquisitionFuncValue(const Eigen::VectorXd& point) const;

        const Eigen::MatrixXd& GetRawDataPoints() const;
        void                   DampData(const std::string& directory_path) const;

        void SetGaussianProcessUpperConfidenceBoundHyperparam(const double hyperparam) { m_gaussian_process_upper_confidence_bound_hyperparam = hyperparam; }

        std::shared_ptr<const PreferenceRegressor> GetRegressor() const { return m_regressor; }

    private:
        const bool                             m_use_map_hyperparams;
        const int                              m_num_options;
        const CurrentBestSelectionStrategy     m_current_best_selection_strategy;
        std::shared_ptr<PreferenceRegressor>   m_regressor;
        std::shared_ptr<PreferenceDataManager> m_data;
        std::vector<Eigen::VectorXd>           m_current_options;
        double                                 m_kernel_signal_var, m_kernel_length_scale, m_noise_level, m_kernel_hyperparams_prior_var, m_btl_scale;
        const KernelType                       m_kernel_type;
        const AcquisitionFuncType              m_acquisition_func_type;
        double                                 m_gaussian_process_upper_confidence_bound_hyperparam;

        void PerformMapEstimation(const int num_map_estimation_iters);
    };
} // namespace sequential_line_search

#endif // SEQUENTIAL_LINE_SEARCH_B200_OPTIMIZERS_HPP
