// pySequentialLineSearch: the Python module of the reference (python/pySequentialLineSearch.cpp:12-152) over the B200 host
// layer. Same module name, enums, classes, method names, keyword arguments and defaults. Vectors cross the boundary as
// 1-D numpy float64 arrays, the data matrix as a 2-D (num_dims x num_points) array, exactly what pybind11's Eigen casters
// produce in the reference; the conversions are written out by hand so the module builds with or without a full Eigen.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <sequential-line-search/optimizers.hpp>

namespace py  = pybind11;
namespace sls = sequential_line_search;
using namespace py::literals;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using Array = py::array_t<double, py::array::forcecast>;

namespace
{
    VectorXd to_vector(const Array& a)
    {
        if (a.ndim() != 1) throw py::value_error("expected a 1-D float array");
        VectorXd v = VectorXd::Zero(a.shape(0));
        auto     r = a.unchecked<1>();
        for (py::ssize_t i = 0; i < a.shape(0); ++i) v(i) = r(i);
        return v;
    }
    py::array_t<double> to_numpy(const VectorXd& v)
    {
        py::array_t<double> a(v.size());
        auto                w = a.mutable_unchecked<1>();
        for (long i = 0; i < (long) v.size(); ++i) w(i) = v(i);
        return a;
    }
    py::array_t<double> to_numpy(const MatrixXd& m)
    {
        py::array_t<double> a({(py::ssize_t) m.rows(), (py::ssize_t) m.cols()});
        auto                w = a.mutable_unchecked<2>();
        for (long i = 0; i < (long) m.rows(); ++i)
            for (long j = 0; j < (long) m.cols(); ++j) w(i, j) = m(i, j);
        return a;
    }
    std::vector<VectorXd> to_vectors(const std::vector<Array>& list)
    {
        std::vector<VectorXd> out;
        for (const Array& a : list) out.push_back(to_vector(a));
        return out;
    }
    py::list to_list(const std::vector<VectorXd>& vs)
    {
        py::list l;
        for (const VectorXd& v : vs) l.append(to_numpy(v));
        return l;
    }

    // Python callables as the initial query generators (pybind11/functional.h in the reference); None = library default
    std::function<std::pair<VectorXd, VectorXd>(const int)> slider_generator(const py::object& f)
    {
        if (f.is_none()) return sls::GenerateRandomSliderEnds;
        return [f](const int num_dims) {
            const py::tuple t = f(num_dims).cast<py::tuple>();
            if (t.size() != 2) throw py::value_error("initial_query_generator must return two points");
            return std::make_pair(to_vector(t[0].cast<Array>()), to_vector(t[1].cast<Array>()));
        };
    }
    sls::InitialQueryGenerator options_generator(const py::object& f)
    {
        if (f.is_none()) return sls::GenerateRandomPoints;
        return [f](const int num_dims, const int num_options) { return to_vectors(f(num_dims, num_options).cast<std::vector<Array>>()); };
    }
} // namespace

PYBIND11_MODULE(pySequentialLineSearch, m)
{
    m.doc() = "sequential-line-search on B200: SequentialLineSearchOptimizer / PreferentialBayesianOptimizer served by libslsgp";

    py::enum_<sls::CurrentBestSelectionStrategy>(m, "CurrentBestSelectionStrategy", py::arithmetic())
        .value("LargestExpectValue", sls::CurrentBestSelectionStrategy::LargestExpectValue)
        .value("LastSelection", sls::CurrentBestSelectionStrategy::LastSelection);
    py::enum_<sls::AcquisitionFuncType>(m, "AcquisitionFuncType", py::arithmetic())
        .value("ExpectedImprovement", sls::AcquisitionFuncType::ExpectedImprovement)
        .value("GaussianProcessUpperConfidenceBound", sls::AcquisitionFuncType::GaussianProcessUpperConfidenceBound);
    py::enum_<sls::KernelType>(m, "KernelType", py::arithmetic())
        .value("ArdSquaredExponentialKernel", sls::KernelType::ArdSquaredExponentialKernel)
        .value("ArdMatern52Kernel", sls::KernelType::ArdMatern52Kernel);

    m.def("generate_random_slider_ends", [](int n) { const auto e = sls::GenerateRandomSliderEnds(n); return py::make_tuple(to_numpy(e.first), to_numpy(e.second)); });
    m.def("generate_centered_fixed_length_random_slider_ends",
          [](int n) { const auto e = sls::GenerateCenteredFixedLengthRandomSliderEnds(n); return py::make_tuple(to_numpy(e.first), to_numpy(e.second)); });

    using SLS = sls::SequentialLineSearchOptimizer;
    py::class_<SLS>(m, "SequentialLineSearchOptimizer")
        .def(py::init([](int num_dims, bool use_slider_enlargement, bool use_map_hyperparams, sls::KernelType kernel_type, sls::AcquisitionFuncType acq,
                         const py::object& generator, sls::CurrentBestSelectionStrategy strategy) {
                 return new SLS(num_dims, use_slider_enlargement, use_map_hyperparams, kernel_type, acq, slider_generator(generator), strategy);
             }),
             "num_dims"_a, "use_slider_enlargement"_a = true, "use_map_hyperparams"_a = true, "kernel_type"_a = sls::KernelType::ArdMatern52Kernel,
             "acquisition_func_type"_a = sls::AcquisitionFuncType::ExpectedImprovement, "initial_query_generator"_a = py::none(),
             "current_best_selection_strategy"_a = sls::CurrentBestSelectionStrategy::LargestExpectValue)
        .def("set_hyperparams", &SLS::SetHyperparams, "kernel_signal_var"_a = 0.500, "kernel_length_scale"_a = 0.500, "noise_level"_a = 0.005,
             "kernel_hyperparams_prior_var"_a = 0.250, "btl_scale"_a = 0.010)
        .def("submit_feedback_data", static_cast<void (SLS::*)(const double)>(&SLS::SubmitFeedbackData), "slider_position"_a)
        .def("submit_feedback_data", static_cast<void (SLS::*)(const double, const int, const int, const int)>(&SLS::SubmitFeedbackData),
             "slider_position"_a, "num_map_estimation_iters"_a, "num_global_search_iters"_a, "num_local_search_iters"_a)
        .def("get_slider_ends", [](const SLS& o) { const auto e = o.GetSliderEnds(); return py::make_tuple(to_numpy(e.first), to_numpy(e.second)); })
        .def("calc_point_from_slider_position", [](const SLS& o, double t) { return to_numpy(o.CalcPointFromSliderPosition(t)); }, "slider_position"_a)
        .def("get_maximizer", [](const SLS& o) { return to_numpy(o.GetMaximizer()); })
        .def("get_preference_value_mean", [](const SLS& o, const Array& p) { return o.GetPreferenceValueMean(to_vector(p)); }, "point"_a)
        .def("get_preference_value_stdev", [](const SLS& o, const Array& p) { return o.GetPreferenceValueStdev(to_vector(p)); }, "point"_a)
        .def("get_acquisition_func_value", [](const SLS& o, const Array& p) { return o.GetAcquisitionFuncValue(to_vector(p)); }, "point"_a)
        .def("get_raw_data_points", [](const SLS& o) { return to_numpy(o.GetRawDataPoints()); })
        .def("damp_data", &SLS::DampData, "directory_path"_a)
        .def("set_gaussian_process_upper_confidence_bound_hyperparam", &SLS::SetGaussianProcessUpperConfidenceBoundHyperparam, "hyperparam"_a);

    using PBO = sls::PreferentialBayesianOptimizer;
    py::class_<PBO>(m, "PreferentialBayesianOptimizer")
        .def(py::init([](int num_dims, bool use_map_hyperparams, sls::KernelType kernel_type, sls::AcquisitionFuncType acq, const py::object& generator,
                         sls::CurrentBestSelectionStrategy strategy, int num_options) {
                 return new PBO(num_dims, use_map_hyperparams, kernel_type, acq, options_generator(generator), strategy, num_options);
             }),
             "num_dims"_a, "use_map_hyperparams"_a = true, "kernel_type"_a = sls::KernelType::ArdMatern52Kernel,
             "acquisition_func_type"_a = sls::AcquisitionFuncType::ExpectedImprovement, "initial_query_generator"_a = py::none(),
             "current_best_selection_strategy"_a = sls::CurrentBestSelectionStrategy::LargestExpectValue, "num_options"_a = 2)
        .def("set_hyperparams", &PBO::SetHyperparams, "kernel_signal_var"_a = 0.500, "kernel_length_scale"_a = 0.500, "noise_level"_a = 0.005,
             "kernel_hyperparams_prior_var"_a = 0.250, "btl_scale"_a = 0.010)
        .def("submit_feedback_data", &PBO::SubmitFeedbackData, "option_index"_a, "num_map_estimation_iters"_a = 0)
        .def("submit_custom_feedback_data",
             [](PBO& o, const Array& chosen, const std::vector<Array>& others, int iters) { o.SubmitCustomFeedbackData(to_vector(chosen), to_vectors(others), iters); },
             "chosen_option"_a, "other_options"_a, "num_map_estimation_iters"_a = 0)
        .def("determine_next_query", &PBO::DetermineNextQuery, "num_global_search_iters"_a = 0, "num_local_search_iters"_a = 0)
        .def("get_current_options", [](const PBO& o) { return to_list(o.GetCurrentOptions()); })
        .def("get_maximizer", [](const PBO& o) { return to_numpy(o.GetMaximizer()); })
        .def("get_preference_value_mean", [](const PBO& o, const Array& p) { return o.GetPreferenceValueMean(to_vector(p)); }, "point"_a)
        .def("get_preference_value_stdev", [](const PBO& o, const Array& p) { return o.GetPreferenceValueStdev(to_vector(p)); }, "point"_a)
        .def("get_acquisition_func_value", [](const PBO& o, const Array& p) { return o.GetAcquisitionFuncValue(to_vector(p)); }, "point"_a)
        .def("get_raw_data_points", [](const PBO& o) { return to_numpy(o.GetRawDataPoints()); })
        .def("damp_data", &PBO::DampData, "directory_path"_a)
        .def("set_gaussian_process_upper_confidence_bound_hyperparam", &PBO::SetGaussianProcessUpperConfidenceBoundHyperparam, "hyperparam"_a);
}
