// Host-side helpers under sequential_line_search::utils, with the names and signatures a program written against the
// reference expects (reference: include/sequential-line-search/utils.hpp:10-60). Implemented in host/src/utils.cpp; the device
// evaluates the Bradley-Terry-Luce terms of the MAP objective itself (csrc/map.cuh), these are for callers and demos.
#ifndef SEQUENTIAL_LINE_SEARCH_B200_UTILS_HPP
#define SEQUENTIAL_LINE_SEARCH_B200_UTILS_HPP

#include <Eigen/Core>
#include <string>

namespace sequential_line_search
{
    namespace utils
    {
        // A point drawn uniformly from the unit cube [0, 1]^n.
        Eigen::VectorXd GenerateRandomVector(unsigned n);

        // Probability, under the Bradley-Terry-Luce model with temperature `scale`, that the FIRST of the options with
        // goodness values f(0), f(1), ... is the one chosen: softmax(f / scale)[0]. Evaluated through the differences
        // f(i) - f(0), so a clear winner cannot overflow.
        double CalcBtl(const Eigen::VectorXd& f, double scale = 1.0);

        // Gradient of CalcBtl with respect to f.
        Eigen::VectorXd CalcBtlDerivative(const Eigen::VectorXd& f, double scale = 1.0);

        // Writes X as comma-separated rows (default stream precision, no padding, no trailing newline).
        void ExportMatrixToCsv(const std::string& file_path, const Eigen::MatrixXd& X);
    } // namespace utils
} // namespace sequential_line_search

#endif
