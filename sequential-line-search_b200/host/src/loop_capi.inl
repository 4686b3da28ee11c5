// extern "C" handles onto the optimiser front-ends and the driver-dependent entry points (MAP fits, FindNextPoint(s)).
// Written ONLY against the public interface both class hierarchies share (include/sequential-line-search/*.hpp of the
// reference and of this host layer), so the very same text is compiled twice:
//   * into libsls_b200_host.so by capi.cpp                       with  SLS_CAPI(name) = b200_##name
//   * by the test suite's CPU checker (the reference's own sources) with        SLS_CAPI(name) = ref_##name
// and the step-level parity tests drive both sides call for call. The including file provides the includes, `using`
// declarations and the helpers to_kernel(int), to_acq(int), to_mat(p, rows, cols), to_vec(p, n), put(vector | matrix, out)
// and SLS_CAPI_TRY { .. } SLS_CAPI_CATCH(value_on_error) (exceptions must not cross the C boundary).

struct SLS_CAPI(SlsHandle)
{
    std::unique_ptr<SequentialLineSearchOptimizer> opt;
    int                                            dims;
};
struct SLS_CAPI(PboHandle)
{
    std::unique_ptr<PreferentialBayesianOptimizer> opt;
    int                                            dims, options;
};

// libc's rand() feeds Eigen's Random() on both sides (initial queries, the starting point of FindGlobalSolution)
void SLS_CAPI(srand)(unsigned seed) { std::srand(seed); }

// ---- SequentialLineSearchOptimizer (include/sequential-line-search/sequential-line-search.hpp:23-125) -------------------
// init_ends: 2 D doubles (both slider ends) or null for the class's default generator (GenerateRandomSliderEnds).
void* SLS_CAPI(sls_create)(int D, int enlarge, int use_map, int kt, int acq, int strategy, const double* init_ends)
{
    SLS_CAPI_TRY
    {
            auto* h = new SLS_CAPI(SlsHandle);
            h->dims = D;
            std::function<std::pair<VectorXd, VectorXd>(const int)> gen = GenerateRandomSliderEnds;
            if (init_ends)
            {
                const VectorXd e0 = to_vec(init_ends, D), e1 = to_vec(init_ends + D, D);
                gen = [e0, e1](const int) { return std::pair<VectorXd, VectorXd>(e0, e1); };
            }
            h->opt.reset(new SequentialLineSearchOptimizer(D, enlarge != 0, use_map != 0, to_kernel(kt), to_acq(acq), gen,
                                                           strategy == 0 ? CurrentBestSelectionStrategy::LargestExpectValue
                                                                         : CurrentBestSelectionStrategy::LastSelection));
            return static_cast<void*>(h);
        }
    SLS_CAPI_CATCH(nullptr)
}
void SLS_CAPI(sls_destroy)(void* h) { delete static_cast<SLS_CAPI(SlsHandle)*>(h); }
void SLS_CAPI(sls_set_hyperparams)(void* h, double a, double r, double b, double prior_var, double btl_scale)
{
    static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->SetHyperparams(a, r, b, prior_var, btl_scale);
}
void SLS_CAPI(sls_set_ucb_hyperparam)(void* h, double beta)
{
    static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->SetGaussianProcessUpperConfidenceBoundHyperparam(beta);
}
// n_map < 0 selects the one-argument overload (the library's own effort heuristic)
int SLS_CAPI(sls_submit)(void* h, double position, int n_map, int n_global, int n_local)
{
    SLS_CAPI_TRY
    {
            auto& o = *static_cast<SLS_CAPI(SlsHandle)*>(h)->opt;
            if (n_map < 0)
                o.SubmitFeedbackData(position);
            else
                o.SubmitFeedbackData(position, n_map, n_global, n_local);
            return 0;
        }
    SLS_CAPI_CATCH(1)
}
void SLS_CAPI(sls_get_slider_ends)(void* h, double* end_0, double* end_1)
{
    const auto ends = static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->GetSliderEnds();
    put(ends.first, end_0);
    put(ends.second, end_1);
}
void SLS_CAPI(sls_get_maximizer)(void* h, double* out) { put(static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->GetMaximizer(), out); }
void SLS_CAPI(sls_calc_point)(void* h, double position, double* out)
{
    put(static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->CalcPointFromSliderPosition(position), out);
}
int SLS_CAPI(sls_num_points)(void* h) { return (int) static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->GetRawDataPoints().cols(); }
void SLS_CAPI(sls_get_raw_data_points)(void* h, double* X_out) { put(static_cast<SLS_CAPI(SlsHandle)*>(h)->opt->GetRawDataPoints(), X_out); }
// what: 0 GetPreferenceValueMean, 1 GetPreferenceValueStdev, 2 GetAcquisitionFuncValue
double SLS_CAPI(sls_query)(void* h, int what, const double* x)
{
    SLS_CAPI_TRY
    {
            auto&          hd = *static_cast<SLS_CAPI(SlsHandle)*>(h);
            const VectorXd xv = to_vec(x, hd.dims);
            return what == 0 ? hd.opt->GetPreferenceValueMean(xv) : what == 1 ? hd.opt->GetPreferenceValueStdev(xv) : hd.opt->GetAcquisitionFuncValue(xv);
        }
    SLS_CAPI_CATCH(std::numeric_limits<double>::quiet_NaN())
}

// ---- PreferentialBayesianOptimizer (include/sequential-line-search/preferential-bayesian-optimizer.hpp:33-156) -----------
// init_options: num_options x D doubles or null for GenerateRandomPoints.
void* SLS_CAPI(pbo_create)(int D, int use_map, int kt, int acq, int strategy, int num_options, const double* init_options)
{
    SLS_CAPI_TRY
    {
            auto* h    = new SLS_CAPI(PboHandle);
            h->dims    = D;
            h->options = num_options;
            InitialQueryGenerator gen = GenerateRandomPoints;
            if (init_options)
            {
                std::vector<VectorXd> pts;
                for (int i = 0; i < num_options; ++i) pts.push_back(to_vec(init_options + (size_t) i * D, D));
                gen = [pts](const int, const int) { return pts; };
            }
            h->opt.reset(new PreferentialBayesianOptimizer(D, use_map != 0, to_kernel(kt), to_acq(acq), gen,
                                                           strategy == 0 ? CurrentBestSelectionStrategy::LargestExpectValue
                                                                         : CurrentBestSelectionStrategy::LastSelection,
                                                           num_options));
            return static_cast<void*>(h);
        }
    SLS_CAPI_CATCH(nullptr)
}
void SLS_CAPI(pbo_destroy)(void* h) { delete static_cast<SLS_CAPI(PboHandle)*>(h); }
void SLS_CAPI(pbo_set_hyperparams)(void* h, double a, double r, double b, double prior_var, double btl_scale)
{
    static_cast<SLS_CAPI(PboHandle)*>(h)->opt->SetHyperparams(a, r, b, prior_var, btl_scale);
}
int SLS_CAPI(pbo_submit)(void* h, int option_index, int n_map)
{
    SLS_CAPI_TRY
    {
            static_cast<SLS_CAPI(PboHandle)*>(h)->opt->SubmitFeedbackData(option_index, n_map);
            return 0;
        }
    SLS_CAPI_CATCH(1)
}
int SLS_CAPI(pbo_determine_next_query)(void* h, int n_global, int n_local)
{
    SLS_CAPI_TRY
    {
            static_cast<SLS_CAPI(PboHandle)*>(h)->opt->DetermineNextQuery(n_global, n_local);
            return 0;
        }
    SLS_CAPI_CATCH(1)
}
void SLS_CAPI(pbo_get_current_options)(void* h, double* out /* num_options x D */)
{
    auto&       hd   = *static_cast<SLS_CAPI(PboHandle)*>(h);
    const auto& opts = hd.opt->GetCurrentOptions();
    for (size_t i = 0; i < opts.size(); ++i) put(opts[i], out + i * (size_t) hd.dims);
}
void SLS_CAPI(pbo_get_maximizer)(void* h, double* out) { put(static_cast<SLS_CAPI(PboHandle)*>(h)->opt->GetMaximizer(), out); }
int  SLS_CAPI(pbo_num_points)(void* h) { return (int) static_cast<SLS_CAPI(PboHandle)*>(h)->opt->GetRawDataPoints().cols(); }

// ---- regressors fitted by the library's own MAP driver ------------------------------------------------------------------------
// Tuples in CSR form (first member preferred). Returns a PreferenceRegressor*; release with pref_fit_destroy.
void* SLS_CAPI(pref_fit)(int kt, int D, int N, const double* X, int P, const unsigned* offsets, const unsigned* idx, int use_map, double a,
                         double r, double b, double prior_var, double btl_scale, unsigned num_iters)
{
    SLS_CAPI_TRY
    {
            std::vector<Preference> prefs;
            for (int t = 0; t < P; ++t) prefs.push_back(Preference(std::vector<unsigned>(idx + offsets[t], idx + offsets[t + 1])));
            return static_cast<void*>(new PreferenceRegressor(to_mat(X, D, N), prefs, use_map != 0, a, r, b, prior_var, btl_scale, num_iters, to_kernel(kt)));
        }
    SLS_CAPI_CATCH(nullptr)
}
void        SLS_CAPI(pref_fit_destroy)(void* h) { delete static_cast<PreferenceRegressor*>(h); }
const void* SLS_CAPI(pref_fit_regressor)(void* h) { return static_cast<const Regressor*>(static_cast<PreferenceRegressor*>(h)); }
void        SLS_CAPI(pref_fit_get_state)(void* h, double* y, double* theta, double* b)
{
    const PreferenceRegressor& r = *static_cast<PreferenceRegressor*>(h);
    if (y) put(r.GetSmallY(), y);
    if (theta) put(r.GetKernelHyperparams(), theta);
    if (b) *b = r.GetNoiseHyperparam();
}
void SLS_CAPI(pref_fit_find_arg_max)(void* h, double* x_out) { put(static_cast<PreferenceRegressor*>(h)->FindArgMax(), x_out); }

// GaussianProcessRegressor with hyper-parameters by MAP estimation; theta_out holds D + 1 values
void* SLS_CAPI(gpr_fit)(int kt, int D, int N, const double* X, const double* y, double* theta_out, double* b_out)
{
    SLS_CAPI_TRY
    {
            auto* r = new GaussianProcessRegressor(to_mat(X, D, N), to_vec(y, N), to_kernel(kt));
            if (theta_out) put(r->GetKernelHyperparams(), theta_out);
            if (b_out) *b_out = r->GetNoiseHyperparam();
            return static_cast<void*>(r);
        }
    SLS_CAPI_CATCH(nullptr)
}
void* SLS_CAPI(gpr_given)(int kt, int D, int N, const double* X, const double* y, const double* theta, double b)
{
    SLS_CAPI_TRY
    { return static_cast<void*>(new GaussianProcessRegressor(to_mat(X, D, N), to_vec(y, N), to_vec(theta, D + 1), b, to_kernel(kt))); }
    SLS_CAPI_CATCH(nullptr)
}
void        SLS_CAPI(gpr_fit_destroy)(void* h) { delete static_cast<GaussianProcessRegressor*>(h); }
const void* SLS_CAPI(gpr_fit_regressor)(void* h) { return static_cast<const Regressor*>(static_cast<GaussianProcessRegressor*>(h)); }

// ---- acquisition_func::FindNextPoint(s) and single-point queries through the Regressor interface ------------------------------
int SLS_CAPI(loop_find_next_point)(const void* r, int D, unsigned n_global, unsigned n_local, int acq, double ucb_beta, double* x_out)
{
    SLS_CAPI_TRY
    {
            put(acquisition_func::FindNextPoint(*static_cast<const Regressor*>(r), n_global, n_local, to_acq(acq), ucb_beta), x_out);
            return 0;
        }
    SLS_CAPI_CATCH(1)
}
int SLS_CAPI(loop_find_next_points)(const void* r, int D, unsigned n_points, unsigned n_global, unsigned n_local, int acq, double ucb_beta, double* X_out)
{
    SLS_CAPI_TRY
    {
            const auto pts = acquisition_func::FindNextPoints(*static_cast<const Regressor*>(r), n_points, n_global, n_local, to_acq(acq), ucb_beta);
            for (size_t i = 0; i < pts.size(); ++i) put(pts[i], X_out + i * (size_t) D);
            return 0;
        }
    SLS_CAPI_CATCH(1)
}
double SLS_CAPI(loop_acq_value)(const void* r, int D, int acq, double ucb_beta, const double* x)
{
    SLS_CAPI_TRY
    { return acquisition_func::CalcAcquisitionValue(*static_cast<const Regressor*>(r), to_vec(x, D), to_acq(acq), ucb_beta); }
    SLS_CAPI_CATCH(std::numeric_limits<double>::quiet_NaN())
}
int SLS_CAPI(loop_acq_derivative)(const void* r, int D, int acq, double ucb_beta, const double* x, double* out)
{
    SLS_CAPI_TRY
    {
            put(acquisition_func::CalcAcquisitionValueDerivative(*static_cast<const Regressor*>(r), to_vec(x, D), to_acq(acq), ucb_beta), out);
            return 0;
        }
    SLS_CAPI_CATCH(1)
}
double SLS_CAPI(loop_predict)(const void* r, int D, int what /* 0 mu, 1 sigma */, const double* x)
{
    SLS_CAPI_TRY
    {
            const Regressor& reg = *static_cast<const Regressor*>(r);
            return what == 0 ? reg.PredictMu(to_vec(x, D)) : reg.PredictSigma(to_vec(x, D));
        }
    SLS_CAPI_CATCH(std::numeric_limits<double>::quiet_NaN())
}
// Slider with the library's enlargement (src/slider.cpp:73-142)
void SLS_CAPI(loop_slider)(int D, const double* end_0, const double* end_1, int enlarge, double* out_0, double* out_1)
{
    const Slider s(to_vec(end_0, D), to_vec(end_1, D), enlarge != 0);
    put(s.end_0, out_0);
    put(s.end_1, out_1);
}
