// sequential_line_search::utils: the two non-inline helpers (reference: src/utils.cpp:8-18).
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
