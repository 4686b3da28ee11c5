// Host-side C++ mirror of the reference's regressor interface for the GP hot path, served by libslsgp (sm_100a).
//
// Same namespace, class names, constructor arguments, public members and method semantics as
//   include/sequential-line-search/kernel-type.hpp:8-22, preference.hpp:9-19, regressor.hpp:10-72,
//   gaussian-process-regressor.hpp:9-52, preference-regressor.hpp:13-88
// of yuki-koyama/sequential-line-search, so code written against those headers compiles against this one. What is
// different is where the numbers come from: every regressor owns a libslsgp context (include/slsgp.h) holding X, the
// Gram matrix, its Cholesky factor, K^-1 and alpha on the GPU; Predict* are one-candidate calls of the batched sweep,
// and PredictBatch / acquisition_func::CalcAcquisitionValues expose the batched form directly. There is no CPU
// fallback: constructing a regressor without a usable CUDA device throws std::runtime_error.
//
// Deviations from the reference surface:
//   * PreferenceRegressor::m_K_llt is an Eigen::LLT computed on the host (O(N^3)); it is built when the header is compiled
//     against the real Eigen (not the bundled include/eigen-lite subset) or with SLS_B200_HOST_LLT defined, and omitted
//     otherwise. m_L, the lower Cholesky factor read back from the device, is always there.
//   * Copies of a device-backed regressor are deep: the copy refits its own device model on first use.
#ifndef SEQUENTIAL_LINE_SEARCH_B200_REGRESSORS_HPP
#define SEQUENTIAL_LINE_SEARCH_B200_REGRESSORS_HPP

#include <Eigen/Core>
#if !defined(SLS_B200_HOST_LLT) && !defined(EIGEN_LITE_CORE)
#define SLS_B200_HOST_LLT 1 // the real Eigen is present: keep the reference's public member m_K_llt
#endif
#ifdef SLS_B200_HOST_LLT
#include <Eigen/Cholesky>
#endif
#include <memory>
#include <mutex>
#include <string>
#include <vector>

struct slsgp_ctx;

namespace sequential_line_search
{
    enum class KernelType
    {
        ArdSquaredExponentialKernel,
        ArdMatern52Kernel,
    };

    // k(x_a, x_b; theta), d k / d theta and d k / d x_a as plain function pointers, as the reference hands them out
    using Kernel                   = double (*)(const Eigen::VectorXd&, const Eigen::VectorXd&, const Eigen::VectorXd&);
    using KernelThetaDerivative    = Eigen::VectorXd (*)(const Eigen::VectorXd&, const Eigen::VectorXd&, const Eigen::VectorXd&);
    using KernelFirstArgDerivative = Eigen::VectorXd (*)(const Eigen::VectorXd&, const Eigen::VectorXd&, const Eigen::VectorXd&);

    // One preference tuple: indices into the columns of X, the first one is the chosen (preferred) point.
    struct Preference : public std::vector<unsigned>
    {
        Preference(unsigned i, unsigned j) : std::vector<unsigned>{i, j} {}
        Preference(unsigned i, unsigned j, unsigned k) : std::vector<unsigned>{i, j, k} {}
        Preference(const std::vector<unsigned>& indices) : std::vector<unsigned>(indices) {}
    };

    class Regressor
    {
    public:
        explicit Regressor(const KernelType kernel_type);
        virtual ~Regressor() {}

        unsigned GetNumDims() const { return GetLargeX().rows(); }

        virtual double          PredictMu(const Eigen::VectorXd& x) const              = 0;
        virtual double          PredictSigma(const Eigen::VectorXd& x) const           = 0;
        virtual Eigen::VectorXd PredictMuDerivative(const Eigen::VectorXd& x) const    = 0;
        virtual Eigen::VectorXd PredictSigmaDerivative(const Eigen::VectorXd& x) const = 0;

        virtual const Eigen::VectorXd& GetKernelHyperparams() const = 0;
        virtual double                 GetNoiseHyperparam() const   = 0;
        virtual const Eigen::MatrixXd& GetLargeX() const            = 0;
        virtual const Eigen::VectorXd& GetSmallY() const            = 0;

        // The data point with the largest posterior mean (one GEMV on the device for the regressors below; N calls of
        // PredictMu for a foreign subclass, as in the reference).
        Eigen::VectorXd PredictMaximumPointFromData() const;

        Kernel                   GetKernel() const { return m_kernel; }
        KernelThetaDerivative    GetKernelThetaDerivative() const { return m_kernel_theta_derivative; }
        KernelFirstArgDerivative GetKernelFirstArgDerivative() const { return m_kernel_first_arg_derivative; }
        KernelType               GetKernelType() const { return m_kernel_type; } // addition: the reference drops the enum

    protected:
        Kernel                   m_kernel;
        KernelThetaDerivative    m_kernel_theta_derivative;
        KernelFirstArgDerivative m_kernel_first_arg_derivative;
        KernelType               m_kernel_type;
    };

    // Common part of the two device-backed regressors: the libslsgp context and the batched entry points.
    class DeviceRegressor : public Regressor
    {
    public:
        explicit DeviceRegressor(const KernelType kernel_type);
        // Value semantics as in the reference (it copy-assigns its temporary regressor, src/acquisition-function.cpp:294): a copy
        // shares NOTHING with its source. The device model is not duplicated eagerly; the copy rebuilds it from its own host
        // state (X, y, hyper-parameters) the first time it is asked for a prediction. Moves take the model along.
        DeviceRegressor(const DeviceRegressor& other);
        DeviceRegressor& operator=(const DeviceRegressor& other);
        DeviceRegressor(DeviceRegressor&& other) noexcept            = default;
        DeviceRegressor& operator=(DeviceRegressor&& other) noexcept = default;

        double          PredictMu(const Eigen::VectorXd& x) const override;
        double          PredictSigma(const Eigen::VectorXd& x) const override;
        Eigen::VectorXd PredictMuDerivative(const Eigen::VectorXd& x) const override;
        Eigen::VectorXd PredictSigmaDerivative(const Eigen::VectorXd& x) const override;

        // Posterior of every column of Xq (D x M) in one sweep; any output pointer may be null.
        void PredictBatch(const Eigen::MatrixXd& Xq,
                          Eigen::VectorXd*       mu,
                          Eigen::VectorXd*       sigma,
                          Eigen::MatrixXd*       mu_derivative    = nullptr,
                          Eigen::MatrixXd*       sigma_derivative = nullptr) const;

        // Gives this regressor's device context (and the model on it) away, for the incremental refit of the next iteration
        // (MapWarmStart::device). The regressor stays valid: like a fresh copy it rebuilds its model if it is used again.
        std::shared_ptr<slsgp_ctx> HandOverDevice();

        bool       HasModel() const { return m_fitted || m_refit_pending; }
        slsgp_ctx* Device() const; // the context holding this regressor's model (rebuilt first if this is a fresh copy)
        std::mutex& DeviceMutex() const { return *m_mutex; } // calls on one context are serialised

    protected:
        // set_data + gram + factor + inverse + alpha for (X, y, theta, noise); fills K / Kinv / L when asked
        void FitOnDevice(const Eigen::MatrixXd& X,
                         const Eigen::VectorXd& y,
                         const Eigen::VectorXd& kernel_hyperparams,
                         double                 noise,
                         Eigen::MatrixXd*       K_out,
                         Eigen::MatrixXd*       Kinv_out,
                         Eigen::MatrixXd*       L_out);
        void EnsureDevice();
        // Rebuild the device model from the host state of the derived class (used by copies).
        virtual void RefitOnDevice() = 0;

        std::shared_ptr<slsgp_ctx>  m_device;
        std::shared_ptr<std::mutex> m_mutex;
        bool                        m_fitted = false, m_data_on_device = false;
        mutable bool                m_refit_pending = false; // a copy that has not rebuilt its model yet
    };

    class GaussianProcessRegressor : public DeviceRegressor
    {
    public:
        // hyper-parameters by MAP estimation (log marginal likelihood + log-normal priors)
        GaussianProcessRegressor(const Eigen::MatrixXd& X,
                                 const Eigen::VectorXd& y,
                                 const KernelType       kernel_type = KernelType::ArdMatern52Kernel);
        // hyper-parameters given
        GaussianProcessRegressor(const Eigen::MatrixXd& X,
                                 const Eigen::VectorXd& y,
                                 const Eigen::VectorXd& kernel_hyperparams,
                                 double                 noise_hyperparam,
                                 const KernelType       kernel_type = KernelType::ArdMatern52Kernel);

        // Additions. DeviceOnly: the same model without the host copies m_K_y / m_K_y_inv (they stay empty), for
        // regressors that only serve predictions, like the temporary one inside FindNextPoints.
        struct DeviceOnly
        {
        };
        GaussianProcessRegressor(const Eigen::MatrixXd& X,
                                 const Eigen::VectorXd& y,
                                 const Eigen::VectorXd& kernel_hyperparams,
                                 double                 noise_hyperparam,
                                 const KernelType       kernel_type,
                                 DeviceOnly);
        // Grows the model by one observation in O(N^2) (slsgp_append_point: bordered update of the factor and the inverse)
        // instead of constructing a new regressor; hyper-parameters stay as they are.
        void AppendPoint(const Eigen::VectorXd& x, double y);

        Eigen::MatrixXd m_K_y;
        Eigen::MatrixXd m_K_y_inv;

        const Eigen::MatrixXd& GetLargeX() const override { return m_X; }
        const Eigen::VectorXd& GetSmallY() const override { return m_y; }
        const Eigen::VectorXd& GetKernelHyperparams() const override { return m_kernel_hyperparams; }
        double                 GetNoiseHyperparam() const override { return m_noise_hyperparam; }

    private:
        void PerformMapEstimation();
        void RefitOnDevice() override;

        Eigen::MatrixXd m_X;
        Eigen::VectorXd m_y;
        Eigen::VectorXd m_kernel_hyperparams;
        double          m_noise_hyperparam = 0.0;
    };

    // Optional starting point of a PreferenceRegressor's MAP fit (addition): the state of the regressor of the previous
    // iteration. Goodness values are carried over to the data points that are still present (matched by coordinates);
    // the optimum searched for is the same, it is just reached in fewer objective evaluations.
    struct MapWarmStart
    {
        Eigen::MatrixXd X;                    // D x N_prev
        Eigen::VectorXd y;                    // N_prev
        Eigen::VectorXd kernel_hyperparams;   // D + 1 (used when hyper-parameters are estimated)
        double          noise_hyperparam = 0.0;
        // the previous regressor's device context (DeviceRegressor::HandOverDevice), or null: with fixed hyper-parameters the
        // new regressor adopts it and extends the factored model by the new data points (slsgp_set_data_extend)
        std::shared_ptr<slsgp_ctx> device;
    };

    class PreferenceRegressor : public DeviceRegressor
    {
    public:
        PreferenceRegressor(const Eigen::MatrixXd&         X,
                            const std::vector<Preference>& D,
                            const bool                     use_map_hyperparams          = false,
                            const double                   default_kernel_signal_var    = 0.500,
                            const double                   default_kernel_length_scale  = 0.500,
                            const double                   default_noise_level          = 0.005,
                            const double                   kernel_hyperparams_prior_var = 0.250,
                            const double                   btl_scale                    = 0.010,
                            const unsigned                 num_map_estimation_iters     = 100,
                            const KernelType               kernel_type = KernelType::ArdMatern52Kernel,
                            const MapWarmStart*            warm_start  = nullptr);

        const bool m_use_map_hyperparams;

        Eigen::VectorXd FindArgMax() const; // the data point with the largest goodness value y_i

        Eigen::MatrixXd         m_X;
        std::vector<Preference> m_D;
        double                  m_noise_hyperparam = 0.0;
        Eigen::VectorXd         m_kernel_hyperparams;
        Eigen::MatrixXd         m_K;
        Eigen::MatrixXd         m_L; // lower Cholesky factor of m_K (from the device)
#ifdef SLS_B200_HOST_LLT
        Eigen::LLT<Eigen::MatrixXd> m_K_llt;
#endif

        void DampData(const std::string& dir_path, const std::string& prefix = "") const; // X.csv and D.csv

        const Eigen::MatrixXd& GetLargeX() const override { return m_X; }
        const Eigen::VectorXd& GetSmallY() const override { return m_y; }
        const Eigen::VectorXd& GetKernelHyperparams() const override { return m_kernel_hyperparams; }
        double                 GetNoiseHyperparam() const override { return m_noise_hyperparam; }

        const double m_default_kernel_signal_var;
        const double m_default_kernel_length_scale;
        const double m_default_noise_level;
        const double m_kernel_hyperparams_prior_var;
        const double m_btl_scale;

        // MAP objective F(y[, a, b, r]) and its gradient at an arbitrary point (what NLopt's callback evaluates in the
        // reference, src/preference-regressor.cpp:129-259); exposed for hosts that drive their own optimiser.
        double EvaluateMapObjective(const Eigen::VectorXd& x, Eigen::VectorXd* gradient) const;
        // Diagnostics of the last MAP run
        unsigned GetNumMapEvaluations() const { return m_num_map_evaluations; }
        // Incremental refit: the number of leading data points whose factored model was taken over from the previous regressor
        int GetNumPointsKept() const { return m_num_points_kept; }

    private:
        Eigen::VectorXd m_y;
        unsigned        m_num_map_evaluations = 0;

        int             m_num_points_kept = 0;

        void PerformMapEstimation(const unsigned num_iters, const MapWarmStart* warm_start);
        void RefitOnDevice() override;
    };

    // K_y = K_f + noise I and K_f for one of the two library kernels, built by the device Gram kernel
    // (regressor.hpp:51-59 of the reference). `kernel` must be a pointer obtained from Regressor::GetKernel().
    Eigen::MatrixXd CalcLargeKY(const Eigen::MatrixXd& X,
                                const Eigen::VectorXd& kernel_hyperparameters,
                                const double           noise_level,
                                const Kernel           kernel);
    Eigen::MatrixXd CalcLargeKF(const Eigen::MatrixXd& X, const Eigen::VectorXd& kernel_hyperparameters, const Kernel kernel);

    // The remaining L1 free functions of the reference (regressor.hpp:42-70), for programs that call them directly; the
    // regressors themselves never do (the sweep and MAP kernels fuse these arrays away). Device kernels for the library's two
    // kernels, the reference's own per-pair loop for a foreign function pointer.
    Eigen::VectorXd CalcSmallK(const Eigen::VectorXd& x,
                               const Eigen::MatrixXd& X,
                               const Eigen::VectorXd& kernel_hyperparameters,
                               const Kernel           kernel);
    Eigen::MatrixXd CalcSmallKSmallXDerivative(const Eigen::VectorXd&         x,
                                               const Eigen::MatrixXd&         X,
                                               const Eigen::VectorXd&         kernel_hyperparameters,
                                               const KernelFirstArgDerivative kernel_first_arg_derivative);
    std::vector<Eigen::MatrixXd> CalcLargeKYThetaDerivative(const Eigen::MatrixXd&      X,
                                                            const Eigen::VectorXd&      kernel_hyperparameters,
                                                            const KernelThetaDerivative kernel_theta_derivative);
    Eigen::MatrixXd CalcLargeKYNoiseLevelDerivative(const Eigen::MatrixXd& X,
                                                    const Eigen::VectorXd& kernel_hyperparameters,
                                                    const double           noise_level);

    // Frees the idle device contexts the library keeps for re-use by the next regressor (addition).
    void ReleaseDeviceResources();
} // namespace sequential_line_search

#endif // SEQUENTIAL_LINE_SEARCH_B200_REGRESSORS_HPP
