// dual_mma.hpp -- the m-dimensional dual optimiser (host side, m is tiny).
//
// In the reference the dual problem  max_{y >= 0} g(y)  is handed to a second nlopt object
// whose algorithm is again NLOPT_LD_MMA (src/api/optimize.c:818-826, deprecated.c:27), i.e. the
// same mma_minimize runs one level down with n' = m variables and m' = 0 constraints, and ITS
// dual problem has dimension 0, which nlopt_optimize_ short-circuits to a single evaluation
// (optimize.c:536-539) -- the closed-form MMA step.  north_star keeps this level on the host.
//
// This file states that two-level recursion directly for the m' = 0 case instead of recursing:
//   DualMMA::solve  = mma.c:145-452 specialised to "no constraints, always feasible"
//   DualMMA::step   = mma.c:59-137 with m = 0 (the level-3 evaluation)
// The objective F(y) (= -val of the level-1 dual evaluation) and its gradient are supplied by
// the caller as a functor; each call is one launch of the n-dimensional dual kernel.
#pragma once

#include <chrono>
#include <cmath>
#include <vector>

namespace nb200 {

inline double wall_seconds()
{
    using clk = std::chrono::steady_clock;
    static const clk::time_point t0 = clk::now();
    return std::chrono::duration<double>(clk::now() - t0).count();
}

// nlopt_isinf (stop.c:219-228): also treats |x| >= 0.99 HUGE_VAL as infinite
inline bool nl_isinf(double x) { return std::fabs(x) >= HUGE_VAL * 0.99 || std::isinf(x); }

// relstop (stop.c:81-86)
inline bool rel_stop(double vold, double vnew, double reltol, double abstol)
{
    if (nl_isinf(vold)) return false;
    const double d = std::fabs(vnew - vold);
    return d < abstol || d < reltol * (std::fabs(vnew) + std::fabs(vold)) * 0.5 || (reltol > 0 && vnew == vold);
}

struct DualStop {
    double ftol_rel = 1e-14, ftol_abs = 0, xtol_rel = 0, xtol_abs = 0;   // optimize.c:822-825
    int maxeval = 100000;                                                // optimize.c:826
    double maxtime = 0;        // <= 0: unlimited (optimize.c:1104-1105 semantics already applied)
};

// return codes are nlopt_result values
enum { kRetSuccess = 1, kRetFtol = 3, kRetXtol = 4, kRetMaxeval = 5, kRetMaxtime = 6, kRetFailure = -1,
       kRetInvalid = -2 };

class DualMMA {
public:
    explicit DualMMA(unsigned m)
        : m_(m), sigma_(m), g_(m), g_cur_(m), ycur_(m), yprev_(m), yprevprev_(m) {}

    // Minimise F over the box [lo, hi]^m starting from (and returning in) y.
    // eval(y, grad) -> F(y), fills grad[m]; returns NaN-safe doubles; `ok` false aborts.
    template <class Eval>
    int solve(Eval &&eval, double *y, const double *lo, const double *hi, const DualStop &st,
              double *fmin_out, long *nevals_out)
    {
        const unsigned m = m_;
        const double start = wall_seconds();
        long nevals = 0;
        int ret = kRetSuccess;
        for (unsigned i = 0; i < m; ++i)                 // optimize.c:547-551
            if (lo[i] > hi[i] || y[i] < lo[i] || y[i] > hi[i]) return kRetInvalid;
        for (unsigned i = 0; i < m; ++i)                 // mma.c:202-210 (no initial step, sigma_min 0)
            sigma_[i] = (nl_isinf(hi[i]) || nl_isinf(lo[i])) ? 1.0 : 0.5 * (hi[i] - lo[i]);
        double rho = 1.0;                                // rho_init default
        bool ok = true;
        double fbase = eval(y, g_.data(), &ok);          // mma.c:218
        if (!ok) return kRetFailure;
        ++nevals;
        double fmin = fbase, fcur = fbase;
        for (unsigned i = 0; i < m; ++i) ycur_[i] = y[i];
        unsigned k = 0;
        auto timed_out = [&] { return st.maxtime > 0 && wall_seconds() - start >= st.maxtime; };
        auto evals_out = [&] { return st.maxeval > 0 && nevals >= st.maxeval; };

        for (;;) {                                       // outer, mma.c:255
            const double fprev = fcur;
            if (evals_out()) ret = kRetMaxeval;          // stopval is -inf: never reached
            else if (timed_out()) ret = kRetMaxtime;
            if (ret != kRetSuccess) break;
            if (++k > 1) yprevprev_ = yprev_;
            yprev_ = ycur_;
            for (;;) {                                   // inner, mma.c:267
                double gval, wval;
                step(y, lo, hi, fbase, rho, &gval, &wval);     // level-3 closed form -> ycur_
                fcur = eval(ycur_.data(), g_cur_.data(), &ok); // mma.c:297
                if (!ok) return kRetFailure;
                ++nevals;
                const bool inner_done = gval >= fcur;          // mma.c:304
                if (fcur < fmin) {                             // mma.c:334 with m' = 0: always "feasible"
                    fbase = fmin = fcur;
                    for (unsigned i = 0; i < m; ++i) { y[i] = ycur_[i]; g_[i] = g_cur_[i]; }
                }
                if (evals_out()) ret = kRetMaxeval;
                else if (timed_out()) ret = kRetMaxtime;
                if (ret != kRetSuccess) goto done;
                if (inner_done) break;
                if (fcur > gval) {                             // mma.c:403-404
                    const double a = 10 * rho, b = 1.1 * (rho + (fcur - gval) / wval);
                    rho = a < b ? a : b;
                }
            }
            if (rel_stop(fprev, fcur, st.ftol_rel, st.ftol_abs)) ret = kRetFtol;   // mma.c:418
            if (x_converged(st)) ret = kRetXtol;                                   // mma.c:420
            if (ret != kRetSuccess) break;
            rho = 0.1 * rho > 1e-5 ? 0.1 * rho : 1e-5;                             // mma.c:425
            if (k > 1)
                for (unsigned i = 0; i < m; ++i) {                                 // mma.c:431-442
                    const double osc = (ycur_[i] - yprev_[i]) * (yprev_[i] - yprevprev_[i]);
                    double s = sigma_[i] * (osc < 0 ? 0.7 : (osc > 0 ? 1.2 : 1));
                    if (!nl_isinf(hi[i]) && !nl_isinf(lo[i])) {
                        const double top = 10 * (hi[i] - lo[i]), bot = 0.01 * (hi[i] - lo[i]);
                        s = s < top ? s : top;
                        s = s > bot ? s : bot;
                    }
                    sigma_[i] = s > 0.0 ? s : 0.0;             // sigma_min = 0
                }
        }
    done:
        *fmin_out = fmin;
        *nevals_out = nevals;
        return ret;
    }

private:
    // One MMA dual evaluation with zero constraints on the m dual variables (mma.c:59-137, m = 0):
    // ycur_ <- argmin of the separable approximant around y; gval/wval as at mma.c:123-125.
    void step(const double *y, const double *lo, const double *hi, double fbase, double rho,
              double *gval, double *wval)
    {
        double gs = fbase, ws = 0;
        for (unsigned i = 0; i < m_; ++i) {
            const double s = sigma_[i];
            if (s == 0) { ycur_[i] = y[i]; continue; }
            double u = g_[i];
            const double v = std::fabs(g_[i]) * s + 0.5 * rho;
            const double s2 = s * s;
            u *= s2;
            const double r = u / (v * s);
            double dy = (u / v) / (-1 - std::sqrt(std::fabs(1 - r * r)));
            double yc = y[i] + dy;
            if (yc > hi[i]) yc = hi[i];
            else if (yc < lo[i]) yc = lo[i];
            if (yc > y[i] + 0.9 * s) yc = y[i] + 0.9 * s;
            else if (yc < y[i] - 0.9 * s) yc = y[i] - 0.9 * s;
            ycur_[i] = yc;
            dy = yc - y[i];
            const double dy2 = dy * dy, dinv = 1.0 / (s2 - dy2), c = s2 * dy;
            gs += (g_[i] * c + (std::fabs(g_[i]) * s + 0.5 * rho) * dy2) * dinv;
            ws += 0.5 * dy2 * dinv;
        }
        *gval = gs;
        *wval = ws;
    }

    // nlopt_stop_x on (ycur, yprev) with unit weights and a uniform xtol_abs (stop.c:98-108).
    // The dual object always carries an xtol_abs array (nlopt_set_xtol_abs1, optimize.c:825).
    bool x_converged(const DualStop &st) const
    {
        double dn = 0, xn = 0;
        for (unsigned i = 0; i < m_; ++i) dn += std::fabs(ycur_[i] - yprev_[i]);
        for (unsigned i = 0; i < m_; ++i) xn += std::fabs(ycur_[i]);
        if (dn < st.xtol_rel * xn) return true;
        for (unsigned i = 0; i < m_; ++i)
            if (std::fabs(ycur_[i] - yprev_[i]) >= st.xtol_abs) return false;
        return true;
    }

    unsigned m_;
    std::vector<double> sigma_, g_, g_cur_, ycur_, yprev_, yprevprev_;
};

}  // namespace nb200
