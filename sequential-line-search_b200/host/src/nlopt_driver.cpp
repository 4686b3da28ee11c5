// SearchDriver state and the bridge to NLopt (sequential-line-search/driver.hpp, nlopt_driver.hpp).
// With SLS_B200_USE_NLOPT the reference's own helper header nlopt-util.hpp (external/nlopt-util, a header-only dependency of the
// reference; found through NLOPT_UTIL_INC at build time, never copied) is included unmodified and linked against libnlopt.a
// (third_party/nlopt/Makefile), so the solver set-up is the reference's by construction.
#include "nlopt_driver.hpp"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#ifdef SLS_B200_USE_NLOPT
#include <nlopt-util.hpp>
#endif

namespace sequential_line_search
{
    namespace
    {
        SearchDriver initial_driver()
        {
#ifdef SLS_B200_USE_NLOPT
            const SearchDriver fallback = SearchDriver::Hybrid;
#else
            const SearchDriver fallback = SearchDriver::Native;
#endif
            const char* env = std::getenv("SLS_B200_DRIVER");
            if (!env) return fallback;
            const std::string v(env);
            if (v == "native") return SearchDriver::Native;
#ifdef SLS_B200_USE_NLOPT
            if (v == "hybrid") return SearchDriver::Hybrid;
            if (v == "reference") return SearchDriver::Reference;
#endif
            return fallback;
        }
        std::atomic<int>& driver_state()
        {
            static std::atomic<int> state((int) initial_driver());
            return state;
        }
    } // namespace

    namespace
    {
        std::atomic<int>& incremental_state()
        {
            static std::atomic<int> state(std::getenv("SLS_B200_INCREMENTAL") && std::atoi(std::getenv("SLS_B200_INCREMENTAL")) != 0 ? 1 : 0);
            return state;
        }
    } // namespace
    void SetIncrementalRefit(bool on) { incremental_state().store(on ? 1 : 0); }
    bool GetIncrementalRefit() { return incremental_state().load() != 0; }

    bool IsNloptAvailable()
    {
#ifdef SLS_B200_USE_NLOPT
        return true;
#else
        return false;
#endif
    }

    void SetSearchDriver(SearchDriver mode)
    {
        if (mode != SearchDriver::Native && !IsNloptAvailable())
            throw std::runtime_error("this build of libsls_b200_host has no NLopt: only SearchDriver::Native is available");
        driver_state().store((int) mode);
    }

    SearchDriver GetSearchDriver() { return (SearchDriver) driver_state().load(); }

    namespace internal
    {
#ifdef SLS_B200_USE_NLOPT
        namespace
        {
            struct Bridge
            {
                const NloptObjective* objective;
                const NloptObjective* inequality;
            };
            double objective_trampoline(const std::vector<double>& x, std::vector<double>& grad, void* data)
            {
                return (*static_cast<Bridge*>(data)->objective)(x, grad);
            }
            double inequality_trampoline(const std::vector<double>& x, std::vector<double>& grad, void* data)
            {
                return (*static_cast<Bridge*>(data)->inequality)(x, grad);
            }
            nlopt::algorithm to_nlopt(NloptAlgorithm a)
            {
                switch (a)
                {
                    case NloptAlgorithm::GN_DIRECT: return nlopt::GN_DIRECT;
                    case NloptAlgorithm::LD_LBFGS: return nlopt::LD_LBFGS;
                    case NloptAlgorithm::LD_TNEWTON: return nlopt::LD_TNEWTON;
                    default: return nlopt::LN_COBYLA;
                }
            }
        } // namespace

        Eigen::VectorXd nlopt_solve(const Eigen::VectorXd& x_initial, const Eigen::VectorXd& upper, const Eigen::VectorXd& lower,
                                    const NloptObjective& objective, NloptAlgorithm algorithm, bool is_maximization, int max_evaluations,
                                    const NloptObjective* inequality)
        {
            Bridge bridge{&objective, inequality};
            if (inequality)
                return nloptutil::solve(x_initial, upper, lower, objective_trampoline, {}, {inequality_trampoline}, to_nlopt(algorithm), &bridge,
                                        is_maximization, max_evaluations);
            return nloptutil::solve(x_initial, upper, lower, objective_trampoline, to_nlopt(algorithm), &bridge, is_maximization, max_evaluations);
        }
#else
        Eigen::VectorXd nlopt_solve(const Eigen::VectorXd&, const Eigen::VectorXd&, const Eigen::VectorXd&, const NloptObjective&, NloptAlgorithm, bool,
                                    int, const NloptObjective*)
        {
            throw std::runtime_error("libsls_b200_host was built without NLopt");
        }
#endif
    } // namespace internal
} // namespace sequential_line_search
