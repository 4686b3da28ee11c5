// Which optimiser drives the two searches of the library (addition; the reference always uses NLopt):
//   the MAP fits          PreferenceRegressor / GaussianProcessRegressor::PerformMapEstimation
//                         (reference: src/preference-regressor.cpp:332-403, src/gaussian-process-regressor.cpp:274-299)
//   the acquisition search acquisition_func::FindNextPoint / FindNextPoints, and the slider enlargement
//                         (reference: src/acquisition-function.cpp:112-167, 246-298; src/slider.cpp:73-122)
// In every mode the objective functions (Gram, Cholesky, posterior, EI/UCB, MAP objective and gradients) run on the GPU.
//
//   Reference  NLopt runs both searches exactly as the reference sets them up (same algorithms, budgets, tolerances, bounds,
//              starting points and std::rand() stream, through the unmodified nloptutil::solve): LD_TNEWTON for the preference
//              MAP fit, GN_DIRECT + LD_TNEWTON for the GPR fit, GN_DIRECT + LD_LBFGS for the acquisition search, two LN_COBYLA
//              solves for the slider enlargement. The Submit -> next-slider step then reproduces the reference's.
//   Hybrid     the MAP fits as in Reference (they are at most a few hundred objective evaluations and define the model the
//              user sees); the acquisition search by the device-resident maximiser (a sweep over 10^5..10^7 candidates plus a
//              batched multi-start ascent: the GPU-native replacement of DIRECT + L-BFGS); slider enlargement in closed form.
//   Native     no NLopt anywhere: the MAP fits by the host layer's own quasi-Newton driver in whitened coordinates.
// Default: Hybrid when the library was built with NLopt (SLS_B200_USE_NLOPT), Native otherwise. The environment variable
// SLS_B200_DRIVER = reference | hybrid | native selects the initial value; SetSearchDriver overrides it process-wide.
#ifndef SEQUENTIAL_LINE_SEARCH_B200_DRIVER_HPP
#define SEQUENTIAL_LINE_SEARCH_B200_DRIVER_HPP

#include <vector>

namespace sequential_line_search
{
    enum class SearchDriver
    {
        Native,
        Hybrid,
        Reference,
    };

    // The GPUs new regressors are built on (addition). With more than one index every regressor owns a multi-GPU group
    // (slsgp_ctx_create_multi): the model is fitted on the first device and acquisition_func::FindNextPoint(s) / the batched
    // queries split their candidates over all of them. Initial value: the environment variable SLS_B200_DEVICES ("0,1,2,3"),
    // else SLS_B200_DEVICE (one index), else device 0.
    void             SetDevices(const std::vector<int>& device_ids);
    std::vector<int> GetDevices();

    // Incremental refit across iterations (addition). With fixed hyper-parameters (use_map_hyperparams ==
    // false) the optimisers hand the device model of the previous regressor to the next one, which grows K, its Cholesky factor and
    // K^-1 by the columns AddNewPoints appended (slsgp_set_data_extend: O(N^2) per new point) instead of rebuilding them. Off by
    // default: at the sizes the optimisers reach (N <= 600) a from-scratch Gram + factor + inverse costs 0.25-0.6 ms of a 30-60 ms
    // iteration, three bordered updates 0.33 ms (profiles/r02n_incremental_refit_study.txt), and the extended factor differs from a
    // rebuilt one in the last bits. Initial value: environment variable SLS_B200_INCREMENTAL=1.
    void SetIncrementalRefit(bool on);
    bool GetIncrementalRefit();

    bool         IsNloptAvailable();                // was the library built with NLopt?
    void         SetSearchDriver(SearchDriver mode); // throws std::runtime_error for Hybrid / Reference without NLopt
    SearchDriver GetSearchDriver();
} // namespace sequential_line_search

#endif
