// acquisition_func::* of the reference (include/sequential-line-search/acquisition-function.hpp:11-79) on libslsgp.
// Single-point calls keep their signatures; the batched forms and the sweep-based global search are what the GPU
// path is for. FindNextPoint replaces NLopt's DIRECT + L-BFGS (src/acquisition-function.cpp:112-167) by a dense
// candidate sweep on the device followed by a bound-constrained quasi-Newton polish of the winner.
#ifndef SEQUENTIAL_LINE_SEARCH_B200_ACQUISITION_HPP
#define SEQUENTIAL_LINE_SEARCH_B200_ACQUISITION_HPP

#include <Eigen/Core>
#include <sequential-line-search/regressors.hpp>
#include <vector>

namespace sequential_line_search
{
    enum class AcquisitionFuncType
    {
        ExpectedImprovement,
        GaussianProcessUpperConfidenceBound,
    };

    namespace acquisition_func
    {
        double CalcAcquisitionValue(const Regressor&          regressor,
                                    const Eigen::VectorXd&    x,
                                    const AcquisitionFuncType func_type,
                                    const double              gaussian_process_upper_confidence_bound_hyperparam = 1.0);

        Eigen::VectorXd CalcAcquisitionValueDerivative(const Regressor&          regressor,
                                                       const Eigen::VectorXd&    x,
                                                       const AcquisitionFuncType func_type,
                                                       const double gaussian_process_upper_confidence_bound_hyperparam = 1.0);

        // Batched form (addition): values and, if `derivatives` is not null, gradients (D x M) for every column of Xq.
        Eigen::VectorXd CalcAcquisitionValues(const DeviceRegressor&    regressor,
                                              const Eigen::MatrixXd&    Xq,
                                              const AcquisitionFuncType func_type,
                                              const double              gaussian_process_upper_confidence_bound_hyperparam = 1.0,
                                              Eigen::MatrixXd*          derivatives                                        = nullptr);

        Eigen::VectorXd FindNextPoint(const Regressor&          regressor,
                                      const unsigned            num_global_search_iters = 100,
                                      const unsigned            num_local_search_iters  = 50,
                                      const AcquisitionFuncType func_type = AcquisitionFuncType::ExpectedImprovement,
                                      const double gaussian_process_upper_confidence_bound_hyperparam = 1.0);

        std::vector<Eigen::VectorXd> FindNextPoints(const Regressor&          regressor,
                                                    const unsigned            num_points,
                                                    const unsigned            num_global_search_iters = 100,
                                                    const unsigned            num_local_search_iters  = 50,
                                                    const AcquisitionFuncType func_type = AcquisitionFuncType::ExpectedImprovement,
                                                    const double gaussian_process_upper_confidence_bound_hyperparam = 1.0);

        // Candidates the global stage evaluates per unit of `num_global_search_iters` (the reference spends that many
        // objective evaluations in DIRECT; a sweep evaluates this many candidates in the time of a few of them).
        constexpr unsigned kCandidatesPerGlobalIter = 1024;
    } // namespace acquisition_func
} // namespace sequential_line_search

#endif // SEQUENTIAL_LINE_SEARCH_B200_ACQUISITION_HPP
