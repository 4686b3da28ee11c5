// Internal: the NLopt-driven searches of the Reference / Hybrid drivers (sequential-line-search/driver.hpp).
#pragma once

#include <Eigen/Core>
#include <functional>
#include <sequential-line-search/driver.hpp>
#include <vector>

namespace sequential_line_search
{
    namespace internal
    {
        // The NLopt callback signature without the void*: grad is EMPTY for a derivative-free algorithm, else it has x's size.
        using NloptObjective = std::function<double(const std::vector<double>& x, std::vector<double>& grad)>;

        enum class NloptAlgorithm
        {
            GN_DIRECT,
            LD_LBFGS,
            LD_TNEWTON,
            LN_COBYLA,
        };

        // nloptutil::solve(x_initial, upper, lower, objective[, {}, {inequality}], algorithm, data, is_maximization, max_evaluations)
        // of the reference's helper (external/nlopt-util/include/nlopt-util.hpp:45-225) with its defaults (ftol_rel = xtol_rel =
        // 1e-6, constraint tolerance 1e-10), the std::function reached through the void* data. `inequality` may be null.
        // Throws std::runtime_error when the library was built without NLopt.
        Eigen::VectorXd nlopt_solve(const Eigen::VectorXd& x_initial, const Eigen::VectorXd& upper, const Eigen::VectorXd& lower,
                                    const NloptObjective& objective, NloptAlgorithm algorithm, bool is_maximization, int max_evaluations,
                                    const NloptObjective* inequality = nullptr);

        inline bool use_nlopt_for_map() { return GetSearchDriver() != SearchDriver::Native; }
        inline bool use_nlopt_for_search() { return GetSearchDriver() == SearchDriver::Reference; }
    } // namespace internal
} // namespace sequential_line_search
