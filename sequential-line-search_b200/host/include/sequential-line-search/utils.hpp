// sequential_line_search::utils, as declared by the reference's include/sequential-line-search/utils.hpp:10-60
// (random vectors, the Bradley-Terry-Luce model, CSV export). Host-side helpers; the device evaluates the BTL terms of
// the MAP objective itself (csrc/map.cuh).
#ifndef SEQUENTIAL_LINE_SEARCH_B200_UTILS_HPP
#define SEQUENTIAL_LINE_SEARCH_B200_UTILS_HPP

#include <Eigen/Core>
#include <cmath>
#include <string>

namespace sequential_line_search
{
    namespace utils
    {
        // a point drawn uniformly from [0, 1]^n (src/utils.cpp:8-11)
        Eigen::VectorXd GenerateRandomVector(unsigned n);

        // Bradley-Terry-Luce likelihood that option 0 is chosen among f(0), ..., f(m - 1):
        //   p = exp(f_0 / s) / sum_i exp(f_i / s)        (utils.hpp:25-29)
        // written through the differences f_i - f_0, which cannot overflow for the winning option.
        inline double CalcBtl(const Eigen::VectorXd& f, double scale = 1.0)
        {
            double denom = 1.0;
            for (long i = 1; i < f.rows(); ++i) denom += std::exp((f(i) - f(0)) / scale);
            return 1.0 / denom;
        }

        // gradient of p with respect to f (utils.hpp:31-52): with e_i = exp((f_i - f_0) / s) and p = 1 / (1 + sum e_i),
        //   dp/df_0 = p^2 sum_i e_i / s,   dp/df_i = -p^2 e_i / s  (i >= 1)
        inline Eigen::VectorXd CalcBtlDerivative(const Eigen::VectorXd& f, double scale = 1.0)
        {
            const long      m = f.rows();
            Eigen::VectorXd d = Eigen::VectorXd::Zero(m);
            double          total = 0.0;
            for (long i = 1; i < m; ++i)
            {
                d(i) = std::exp((f(i) - f(0)) / scale);
                total += d(i);
            }
            const double p = 1.0 / (1.0 + total), w = p * p / scale;
            for (long i = 1; i < m; ++i) d(i) = -w * d(i);
            if (m > 0) d(0) = w * total;
            return d;
        }

        // comma-separated rows, stream precision, no alignment, no trailing newline (src/utils.cpp:13-18)
        void ExportMatrixToCsv(const std::string& file_path, const Eigen::MatrixXd& X);
    } // namespace utils
} // namespace sequential_line_search

#endif
