// sequential_line_search::utils: the two non-inline helpers (reference: src/utils.cpp:8-18).
#include <cmath>
#include <fstream>
#include <sequential-line-search/utils.hpp>

namespace sequential_line_search
{
    namespace utils
    {
        Eigen::VectorXd GenerateRandomVector(unsigned n)
        {
            // Eigen's Random() is uniform in [-1, 1]; the reference maps it to [0, 1] the same way
            Eigen::VectorXd v = Eigen::VectorXd::Random(n);
            for (unsigned i = 0; i < n; ++i) v(i) = 0.5 * (v(i) + 1.0);
            return v;
        }

        // p = 1 / (1 + sum_{i >= 1} e_i) with e_i = exp((f_i - f_0) / s)          (reference: utils.hpp:25-29)
        double CalcBtl(const Eigen::VectorXd& f, double scale)
        {
            double denom = 1.0;
            for (long i = 1; i < f.rows(); ++i) denom += std::exp((f(i) - f(0)) / scale);
            return 1.0 / denom;
        }

        // dp/df_0 = p^2 sum_i e_i / s,   dp/df_i = -p^2 e_i / s  (i >= 1)           (reference: utils.hpp:31-52)
        Eigen::VectorXd CalcBtlDerivative(const Eigen::VectorXd& f, double scale)
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

        void ExportMatrixToCsv(const std::string& file_path, const Eigen::MatrixXd& X)
        {
            std::ofstream out(file_path);
            for (long i = 0; i < X.rows(); ++i)
            {
                if (i > 0) out << '\n';
                for (long j = 0; j < X.cols(); ++j)
                {
                    if (j > 0) out << ',';
                    out << X(i, j);
                }
            }
        }
    } // namespace utils
} // namespace sequential_line_search
