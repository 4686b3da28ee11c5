// Bound-constrained limited-memory quasi-Newton minimiser used by the host layer where the reference calls NLopt
// (LD_TNEWTON for the MAP fits, LD_LBFGS for the acquisition polish; external/nlopt-util/include/nlopt-util.hpp:45-196).
// Projected L-BFGS: two-loop recursion on the free variables, Armijo backtracking along the projected path.
// The objective runs on the GPU (one libslsgp call per evaluation); this driver is plain host logic.
#pragma once

#include <algorithm>
#include <cmath>
#include <deque>
#include <functional>
#include <limits>
#include <vector>

namespace sequential_line_search
{
    namespace internal
    {
        struct MinimizeResult
        {
            std::vector<double> x;
            double              f         = 0.0;
            unsigned            evals     = 0;
            bool                converged = false;
        };

        // fun(x, grad) returns the value to MINIMISE and fills grad (same length as x). It may return +inf (grad
        // ignored) for points it cannot evaluate; the line search backs off from those.
        using Objective = std::function<double(const std::vector<double>&, std::vector<double>&)>;

        inline MinimizeResult minimize_bounded(const Objective& fun, std::vector<double> x, const std::vector<double>& lo,
                                               const std::vector<double>& hi, unsigned max_evals, double gtol = 1e-8,
                                               double ftol = 1e-15, size_t history = 12)
        {
            const size_t n = x.size();
            for (size_t i = 0; i < n; ++i) x[i] = std::min(std::max(x[i], lo[i]), hi[i]);
            MinimizeResult      res;
            std::vector<double> g(n), xn(n), gn(n), d(n), pg(n);
            double              f = fun(x, g);
            res.evals             = 1;
            std::deque<std::vector<double>> S, Y;
            std::deque<double>              R;
            int                             flat = 0;
            const auto is_active = [&](size_t i) { return (x[i] <= lo[i] && g[i] > 0.0) || (x[i] >= hi[i] && g[i] < 0.0); };
            while (std::isfinite(f))
            {
                double pg_inf = 0.0;
                for (size_t i = 0; i < n; ++i)
                {
                    pg[i]  = is_active(i) ? 0.0 : g[i];
                    pg_inf = std::max(pg_inf, std::fabs(pg[i]));
                }
                if (pg_inf <= gtol)
                {
                    res.converged = true;
                    break;
                }
                if (res.evals >= max_evals) break;

                // d = -H pg (two-loop recursion), restricted to the free variables
                d = pg;
                std::vector<double> a(S.size());
                for (size_t k = S.size(); k-- > 0;)
                {
                    double s = 0.0;
                    for (size_t i = 0; i < n; ++i) s += S[k][i] * d[i];
                    a[k] = R[k] * s;
                    for (size_t i = 0; i < n; ++i) d[i] -= a[k] * Y[k][i];
                }
                if (!S.empty())
                {
                    double yy = 0.0;
                    for (size_t i = 0; i < n; ++i) yy += Y.back()[i] * Y.back()[i];
                    const double gamma = 1.0 / (R.back() * yy);
                    for (size_t i = 0; i < n; ++i) d[i] *= gamma;
                }
                for (size_t k = 0; k < S.size(); ++k)
                {
                    double s = 0.0;
                    for (size_t i = 0; i < n; ++i) s += Y[k][i] * d[i];
                    const double b = R[k] * s;
                    for (size_t i = 0; i < n; ++i) d[i] += (a[k] - b) * S[k][i];
                }
                double slope = 0.0;
                for (size_t i = 0; i < n; ++i)
                {
                    d[i]  = is_active(i) ? 0.0 : -d[i];
                    slope += d[i] * g[i];
                }
                if (!(slope < 0.0)) // not a descent direction: steepest descent on the free variables
                {
                    slope = 0.0;
                    for (size_t i = 0; i < n; ++i) d[i] = -pg[i], slope += d[i] * g[i];
                    S.clear(), Y.clear(), R.clear();
                }

                double t = S.empty() ? std::min(1.0, 1.0 / pg_inf) : 1.0;
                double fn = std::numeric_limits<double>::infinity();
                bool   ok = false;
                for (int ls = 0; ls < 40 && res.evals < max_evals + 8; ++ls, t *= 0.5)
                {
                    double decrease = 0.0, moved = 0.0;
                    for (size_t i = 0; i < n; ++i)
                    {
                        xn[i] = std::min(std::max(x[i] + t * d[i], lo[i]), hi[i]);
                        decrease += g[i] * (xn[i] - x[i]);
                        moved = std::max(moved, std::fabs(xn[i] - x[i]));
                    }
                    if (moved == 0.0) break;
                    fn = fun(xn, gn);
                    ++res.evals;
                    if (std::isfinite(fn) && fn <= f + 1e-4 * decrease)
                    {
                        ok = true;
                        break;
                    }
                }
                if (!ok)
                {
                    if (S.empty()) break; // steepest descent failed too: stationary to working precision
                    S.clear(), Y.clear(), R.clear();
                    continue;
                }
                std::vector<double> s(n), y(n);
                double              sy = 0.0, ss = 0.0, yy = 0.0;
                for (size_t i = 0; i < n; ++i)
                {
                    s[i] = xn[i] - x[i], y[i] = gn[i] - g[i];
                    sy += s[i] * y[i], ss += s[i] * s[i], yy += y[i] * y[i];
                }
                if (sy > 1e-10 * std::sqrt(ss * yy))
                {
                    S.push_back(s), Y.push_back(y), R.push_back(1.0 / sy);
                    if (S.size() > history) S.pop_front(), Y.pop_front(), R.pop_front();
                }
                flat = (f - fn <= ftol * std::max(1.0, std::fabs(f))) ? flat + 1 : 0;
                x.swap(xn), g.swap(gn), f = fn;
                if (flat >= 5) // the objective no longer changes in its last digits: stationary to working precision
                {
                    res.converged = true;
                    break;
                }
            }
            res.x = x, res.f = f;
            return res;
        }
    } // namespace internal
} // namespace sequential_line_search
