// Internal helpers of the host layer: libslsgp context ownership, status -> exception, enum translation.
#pragma once

#include "../../../include/slsgp.h"

#include <Eigen/Core>
#include <memory>
#include <sequential-line-search/acquisition.hpp>
#include <sequential-line-search/regressors.hpp>
#include <stdexcept>
#include <string>
#include <vector>

namespace sequential_line_search
{
    namespace internal
    {
        // Any non-OK status of the C ABI becomes a std::runtime_error carrying libslsgp's message. (The reference only
        // asserts; a failed Cholesky or a NaN input is silently propagated there.)
        inline void check(slsgp_ctx* ctx, slsgp_status s, const char* what)
        {
            if (s == SLSGP_OK) return;
            std::string msg = std::string("libslsgp: ") + what + ": " + slsgp_status_string(s);
            if (ctx && slsgp_last_error(ctx) && *slsgp_last_error(ctx)) msg += std::string(" (") + slsgp_last_error(ctx) + ")";
            throw std::runtime_error(msg);
        }

        // New context on the device named by SLS_B200_DEVICE (default 0). Throws when there is no usable GPU.
        std::shared_ptr<slsgp_ctx> make_device();
        void                       drain_device_pool(); // destroy the idle pooled contexts
        void                       set_device_list(const std::vector<int>& ids);
        std::vector<int>           get_device_list();

        inline slsgp_kernel_type to_abi(KernelType t)
        {
            return t == KernelType::ArdSquaredExponentialKernel ? SLSGP_KERNEL_ARD_SQUARED_EXP : SLSGP_KERNEL_ARD_MATERN52;
        }
        inline slsgp_acq_type to_abi(AcquisitionFuncType t)
        {
            return t == AcquisitionFuncType::ExpectedImprovement ? SLSGP_ACQ_EXPECTED_IMPROVEMENT : SLSGP_ACQ_GP_UCB;
        }

        // Function-pointer <-> enum mapping for the two library kernels (Regressor keeps both).
        Kernel                   kernel_of(KernelType t);
        KernelThetaDerivative    kernel_theta_derivative_of(KernelType t);
        KernelFirstArgDerivative kernel_first_arg_derivative_of(KernelType t);
        bool                     kernel_type_of(Kernel k, KernelType* out);
    } // namespace internal
} // namespace sequential_line_search
