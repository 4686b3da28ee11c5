// TEST INFRASTRUCTURE (oracle/): extern "C" handles onto the reference's optimiser front-ends and driver-dependent entry
// points, for the library variant that is linked against the REAL NLopt (third_party/nlopt/_build/libnlopt.a, built from
// the reference's own external/nlopt submodule): oracle/_ref/libsls_ref_loop.so. It is the reference's unmodified code that
// runs the whole loop here - PreferenceRegressor's LD_TNEWTON MAP fit (src/preference-regressor.cpp:332-403),
// FindGlobalSolution's GN_DIRECT + LD_LBFGS (src/acquisition-function.cpp:112-167), the slider's two COBYLA solves
// (src/slider.cpp:73-122), GaussianProcessRegressor's DIRECT + TNEWTON fit (src/gaussian-process-regressor.cpp:274-299) -
// with include/eigen-lite standing in for Eigen. The facade itself is the text shared with the product's host layer
// (sequential-line-search_b200/host/src/loop_capi.inl), compiled here with the ref_ prefix.
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <sequential-line-search/acquisition-function.hpp>
#include <sequential-line-search/gaussian-process-regressor.hpp>
#include <sequential-line-search/preference-data-manager.hpp>
#include <sequential-line-search/preference-regressor.hpp>
#include <sequential-line-search/preferential-bayesian-optimizer.hpp>
#include <sequential-line-search/sequential-line-search.hpp>
#include <sequential-line-search/slider.hpp>
#include <stdexcept>
#include <string>
#include <vector>

using Eigen::MatrixXd;
using Eigen::VectorXd;
using namespace sequential_line_search;

namespace
{
    thread_local std::string g_error;

    KernelType          to_kernel(int kt) { return kt == 0 ? KernelType::ArdSquaredExponentialKernel : KernelType::ArdMatern52Kernel; }
    AcquisitionFuncType to_acq(int t) { return t == 0 ? AcquisitionFuncType::ExpectedImprovement : AcquisitionFuncType::GaussianProcessUpperConfidenceBound; }
    MatrixXd            to_mat(const double* p, int rows, int cols)
    {
        MatrixXd m(rows, cols);
        if (rows && cols) std::memcpy(m.data(), p, sizeof(double) * size_t(rows) * size_t(cols));
        return m;
    }
    VectorXd to_vec(const double* p, int n)
    {
        VectorXd v(n);
        if (n) std::memcpy(v.data(), p, sizeof(double) * size_t(n));
        return v;
    }
    void put(const VectorXd& v, double* out) { std::memcpy(out, v.data(), sizeof(double) * size_t(v.size())); }
    void put(const MatrixXd& m, double* out) { std::memcpy(out, m.data(), sizeof(double) * size_t(m.rows()) * size_t(m.cols())); }
} // namespace

#define SLS_CAPI(name) ref_##name
#define SLS_CAPI_TRY try
#define SLS_CAPI_CATCH(value)       \
    catch (const std::exception& e) \
    {                               \
        g_error = e.what();         \
        return value;               \
    }

extern "C"
{
    const char* ref_loop_last_error() { return g_error.c_str(); }
#include "../sequential-line-search_b200/host/src/loop_capi.inl"
}
