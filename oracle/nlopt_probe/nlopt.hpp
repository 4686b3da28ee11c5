// TEST INFRASTRUCTURE (oracle/): a *probe* stand-in for <nlopt.hpp>, used only when compiling the reference's
// unchanged sources into oracle/_ref/libsls_ref_probe.so.
//
// The reference's MAP objectives live in anonymous namespaces (src/preference-regressor.cpp:129,
// src/gaussian-process-regressor.cpp:141) and are reachable only through the NLopt callback that
// nloptutil::solve() registers (external/nlopt-util/include/nlopt-util.hpp:45-196). This header implements the
// slice of the nlopt C++ API that nlopt-util touches, but `opt::optimize` does not optimise: it hands the
// registered callback to a test-controlled hook, which can (a) evaluate the reference objective + gradient at
// arbitrary points and (b) force the "solution" NLopt returns, so that a reference regressor can be put into
// any state (y, theta, b). Nothing in the product links this.
#ifndef SLSGP_ORACLE_NLOPT_PROBE_HPP
#define SLSGP_ORACLE_NLOPT_PROBE_HPP

#include <functional>
#include <stdexcept>
#include <vector>

namespace nlopt
{
    enum algorithm
    {
        GN_DIRECT  = 0,
        LD_LBFGS   = 11,
        LD_TNEWTON = 15,
        LN_COBYLA  = 25,
    };

    typedef double (*vfunc)(const std::vector<double>& x, std::vector<double>& grad, void* data);

    class roundoff_limited : public std::runtime_error
    {
    public:
        roundoff_limited() : std::runtime_error("nlopt roundoff-limited") {}
    };
    class forced_stop : public std::runtime_error
    {
    public:
        forced_stop() : std::runtime_error("nlopt forced stop") {}
    };

    class opt;

    namespace probe
    {
        struct Hook
        {
            // Called from opt::optimize. May evaluate `f` anywhere and may overwrite x (the returned solution).
            std::function<void(const opt& solver, vfunc f, void* data, std::vector<double>& x, double& fval)> fn;
        };
        inline Hook& hook()
        {
            static Hook h;
            return h;
        }
    } // namespace probe

    class opt
    {
    public:
        opt(algorithm a, unsigned n) : m_alg(a), m_n(n), m_maxeval(0), m_f(nullptr), m_data(nullptr), m_max(false) {}

        void set_upper_bounds(const std::vector<double>& u) { m_ub = u; }
        void set_lower_bounds(const std::vector<double>& l) { m_lb = l; }
        void set_maxeval(int n) { m_maxeval = n; }
        void set_ftol_rel(double) {}
        void set_xtol_rel(double) {}
        void set_min_objective(vfunc f, void* d) { m_f = f, m_data = d, m_max = false; }
        void set_max_objective(vfunc f, void* d) { m_f = f, m_data = d, m_max = true; }
        void add_equality_constraint(vfunc, void*, double) {}
        void add_inequality_constraint(vfunc, void*, double) {}
        void get_initial_step(const std::vector<double>&, std::vector<double>& dx) const
        {
            for (auto& d : dx) d = 1.0;
        }
        void set_initial_step(const std::vector<double>&) {}
        int  get_numevals() const { return 0; }

        void optimize(std::vector<double>& x, double& fval)
        {
            if (probe::hook().fn)
            {
                probe::hook().fn(*this, m_f, m_data, x, fval);
            }
            else
            {
                std::vector<double> no_grad;
                fval = m_f(x, no_grad, m_data);
            }
        }

        algorithm                  get_algorithm() const { return m_alg; }
        unsigned                   get_dimension() const { return m_n; }
        int                        get_maxeval() const { return m_maxeval; }
        bool                       is_maximization() const { return m_max; }
        const std::vector<double>& upper() const { return m_ub; }
        const std::vector<double>& lower() const { return m_lb; }

    private:
        algorithm           m_alg;
        unsigned            m_n;
        int                 m_maxeval;
        vfunc               m_f;
        void*               m_data;
        bool                m_max;
        std::vector<double> m_ub, m_lb;
    };
} // namespace nlopt

#endif
