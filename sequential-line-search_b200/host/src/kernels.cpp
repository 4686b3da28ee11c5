// The two library kernels as host scalar functions. They exist because the reference's Regressor hands out raw
// function pointers (GetKernel() & co, include/sequential-line-search/regressor.hpp:30-32) and callers may evaluate a
// single pair with them; nothing on the device path calls these (the Gram, sweep and MAP kernels fuse the same
// formulas). Formulas: external/mathtoolbox/src/kernel-functions.cpp:7-20 (SE), :22-50 (SE theta derivative),
// :81-93 (SE x derivative, 2x the analytic one -- reproduced), :95-112, :114-142, :179-212 (Matern 5/2; its x
// derivative is defined as 0 when sqrt(5) r < 1e-30).
#include "device.hpp"

#include <cmath>

namespace
{
    using Eigen::VectorXd;

    // r^2 = sum ((a_i - b_i) / l_i)^2 with theta = (signal variance, l_1 .. l_D)
    double scaled_sq_dist(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        double r2 = 0.0;
        for (int i = 0; i < (int) a.size(); ++i)
        {
            const double t = (a(i) - b(i)) / theta(i + 1);
            r2 += t * t;
        }
        return r2;
    }

    double se_kernel(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        return theta(0) * std::exp(-0.5 * scaled_sq_dist(a, b, theta));
    }
    VectorXd se_theta_derivative(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        const double e = std::exp(-0.5 * scaled_sq_dist(a, b, theta)), k = theta(0) * e;
        VectorXd     out(theta.size());
        out(0) = e;
        for (int i = 0; i < (int) a.size(); ++i)
        {
            const double d = a(i) - b(i), l = theta(i + 1);
            out(i + 1)     = k * d * d / (l * l * l);
        }
        return out;
    }
    VectorXd se_first_arg_derivative(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        const double k = se_kernel(a, b, theta);
        VectorXd     out(a.size());
        for (int i = 0; i < (int) a.size(); ++i) out(i) = -2.0 * k * (a(i) - b(i)) / (theta(i + 1) * theta(i + 1));
        return out;
    }

    double matern52_kernel(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        const double r2 = scaled_sq_dist(a, b, theta), s = std::sqrt(5.0 * r2);
        return theta(0) * (1.0 + s + (5.0 / 3.0) * r2) * std::exp(-s);
    }
    VectorXd matern52_theta_derivative(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        const double r2 = scaled_sq_dist(a, b, theta), s = std::sqrt(5.0 * r2), e = std::exp(-s);
        VectorXd     out(theta.size());
        out(0) = (1.0 + s + (5.0 / 3.0) * r2) * e;
        for (int i = 0; i < (int) a.size(); ++i)
        {
            const double d = a(i) - b(i), l = theta(i + 1);
            out(i + 1)     = (5.0 / 3.0) * theta(0) * e * (1.0 + s) * d * d / (l * l * l);
        }
        return out;
    }
    VectorXd matern52_first_arg_derivative(const VectorXd& a, const VectorXd& b, const VectorXd& theta)
    {
        const double r2 = scaled_sq_dist(a, b, theta), s = std::sqrt(5.0 * r2), e = std::exp(-s);
        VectorXd     out = VectorXd::Zero(a.size());
        if (s < 1e-30) return out;
        const double scale = 1.0 + s + (5.0 / 3.0) * r2, ds_dr2 = 0.5 * std::sqrt(5.0 / r2);
        for (int i = 0; i < (int) a.size(); ++i)
        {
            const double dr2 = 2.0 * (a(i) - b(i)) / (theta(i + 1) * theta(i + 1)); // d r^2 / d a_i
            const double ds  = ds_dr2 * dr2;
            out(i)           = theta(0) * ((ds + (5.0 / 3.0) * dr2) * e - scale * ds * e);
        }
        return out;
    }
} // namespace

namespace sequential_line_search
{
    namespace internal
    {
        Kernel kernel_of(KernelType t) { return t == KernelType::ArdSquaredExponentialKernel ? se_kernel : matern52_kernel; }
        KernelThetaDerivative kernel_theta_derivative_of(KernelType t)
        {
            return t == KernelType::ArdSquaredExponentialKernel ? se_theta_derivative : matern52_theta_derivative;
        }
        KernelFirstArgDerivative kernel_first_arg_derivative_of(KernelType t)
        {
            return t == KernelType::ArdSquaredExponentialKernel ? se_first_arg_derivative : matern52_first_arg_derivative;
        }
        bool kernel_type_of(Kernel k, KernelType* out)
        {
            if (k == se_kernel) return *out = KernelType::ArdSquaredExponentialKernel, true;
            if (k == matern52_kernel) return *out = KernelType::ArdMatern52Kernel, true;
            return false;
        }
    } // namespace internal
} // namespace sequential_line_search
