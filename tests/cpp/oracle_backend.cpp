// tests/cpp/oracle_backend.cpp -- TEST ONLY.  A CPU stand-in for DeviceBackend, built on the
// oracle's plain-C routines, so that the host-side logic of the product (nlopt_api.cpp +
// ccsa_driver.cpp + dual_mma.hpp) can be exercised by `pytest -m "not gpu"` on a machine without
// a GPU.  It is linked into tests/_build/libnlopt_hosttest.so, never into libnlopt_b200.so.
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../nlopt_b200/csrc/backend_factory.hpp"
#include "../../oracle/ccsa_port.h"

namespace nb200 {

namespace {
double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

class OracleBackend : public Backend {
public:
    explicit OracleBackend(const BackendConfig &c) : cfg(c), n_(c.n)
    {
        m_ = 0;
        for (const FuncSpec &f : c.constraints) m_ += f.m;
        x.assign(c.x0_host, c.x0_host + n_);
        lb.assign(c.lb, c.lb + n_);
        ub.assign(c.ub, c.ub + n_);
        xcur = x;
        xprev.assign(n_, 0.0);
        xprevprev.assign(n_, 0.0);
        sigma.assign(n_, 0.0);
        g.assign(n_, 0.0);
        gcur.assign(n_, 0.0);
        G.assign((size_t) m_ * n_, 0.0);
        Gcur.assign((size_t) m_ * n_, 0.0);
        zeros.assign(m_ ? m_ : 1, 0.0);
        gc_tmp.assign(m_ ? m_ : 1, 0.0);
    }
    unsigned n() const override { return n_; }
    unsigned m() const override { return m_; }
    unsigned num_constraint_objects() const override { return (unsigned) cfg.constraints.size(); }
    unsigned constraint_dim(unsigned ic) const override { return cfg.constraints[ic].m; }

    bool init_sigma(double sigma_min) override
    {
        port_sigma_init(n_, lb.data(), ub.data(), cfg.sigma_init, sigma_min, sigma.data());
        return true;
    }
    bool eval_objective(Slot slot, bool want_grad, double *value) override
    {
        if (!cfg.objective.f) { err = "host test backend needs a host objective"; return false; }
        const double *xs = (slot == kBase ? x : xcur).data();
        double *gs = want_grad ? (slot == kBase ? g : gcur).data() : nullptr;
        const double t0 = now_s();
        double L = cfg.objective.f(n_, xs, gs, cfg.objective.data);
        cb += now_s() - t0;
        *value = L;
        if (!cfg.penalty) return true;
        // the augmented-Lagrangian objective (PenaltySpec): plain host loops standing in for the device kernel
        const PenaltySpec &ps = *cfg.penalty;
        if (ps.nevals_p) ++*ps.nevals_p;
        if (ps.force_stop && *ps.force_stop) return true;
        unsigned maxdim = 1;
        for (const FuncSpec &fs : ps.eq) maxdim = fs.m > maxdim ? fs.m : maxdim;
        for (const FuncSpec &fs : ps.ineq) maxdim = fs.m > maxdim ? fs.m : maxdim;
        std::vector<double> vals(maxdim), rows(want_grad ? (size_t) maxdim * n_ : 1);
        unsigned ii = 0;
        for (int pass = 0; pass < 2; ++pass) {
            ii = 0;
            for (const FuncSpec &fs : (pass == 0 ? ps.eq : ps.ineq)) {
                if (fs.f) vals[0] = fs.f(n_, xs, want_grad ? rows.data() : nullptr, fs.data);
                else fs.mf(fs.m, vals.data(), n_, xs, want_grad ? rows.data() : nullptr, fs.data);
                if (ps.force_stop && *ps.force_stop) return true;
                for (unsigned k = 0; k < fs.m; ++k, ++ii) {
                    const double v = vals[k] + (pass == 0 ? ps.lambda[ii] : ps.mu[ii]) / ps.rho;
                    if (pass == 1 && !(v > 0)) continue;
                    L += 0.5 * ps.rho * v * v;
                    if (want_grad)
                        for (unsigned j = 0; j < n_; ++j) gs[j] += (ps.rho * v) * rows[(size_t) k * n_ + j];
                }
            }
        }
        *value = L;
        return true;
    }
    bool eval_constraint(Slot slot, unsigned ic, unsigned row0, bool want_grad, double *values) override
    {
        const FuncSpec &f = cfg.constraints[ic];
        double *gp = want_grad ? (slot == kBase ? G : Gcur).data() + (size_t) row0 * n_ : nullptr;
        const double *xp = (slot == kBase ? x : xcur).data();
        const double t0 = now_s();
        if (f.f) values[0] = f.f(n_, xp, gp, f.data);
        else if (f.mf) f.mf(f.m, values, n_, xp, gp, f.data);
        else { err = "device constraint in host test backend"; return false; }
        cb += now_s() - t0;
        return true;
    }
    bool dual_eval(const double *y, const DualScalars &sc, bool materialize, DualSums *out) override
    {
        // raw sums: run the oracle with zero constants (NaN kept so MMA's on/off rule applies)
        std::vector<double> c0(m_ ? m_ : 1, 0.0);
        for (unsigned i = 0; i < m_; ++i)
            if (cfg.variant == kMMA && std::isnan(sc.fcval[i])) c0[i] = sc.fcval[i];
        port_dual_in in;
        in.n = n_; in.m = m_;
        in.x = x.data(); in.lb = lb.data(); in.ub = ub.data(); in.sigma = sigma.data();
        in.grad_f = g.data(); in.grad_c = G.data();
        in.f0 = 0.0; in.rho = sc.rho; in.c0 = c0.data(); in.rhoc = sc.rhoc;
        std::vector<double> scratch;
        port_dual_out o;
        if (materialize) o.xcur = xcur.data();
        else { scratch.resize(n_); o.xcur = scratch.data(); }
        o.gc = gc_tmp.data();
        const double r = cfg.variant == kMMA ? port_dual_mma(&in, y, nullptr, &o) : port_dual_ccsaq(&in, y, nullptr, &o);
        out->val = -r;
        out->gval = o.g0;
        out->wval = o.w;
        for (unsigned i = 0; i < m_; ++i) out->gc[i] = o.gc[i];
        return true;
    }
    void accept_candidate() override
    {
        x = xcur;
        g = gcur;
        G = Gcur;
    }
    bool first_outer() override { xprev = xcur; return true; }
    bool end_outer(unsigned k, double sigma_min, double *dnorm, double *xnorm, bool *below) override
    {
        double d = 0, s = 0;
        bool all = true;
        for (unsigned j = 0; j < n_; ++j) {
            const double w = cfg.x_weights ? cfg.x_weights[j] : 1.0;
            d += w * std::fabs(xcur[j] - xprev[j]);
            s += w * std::fabs(xcur[j]);
            if (cfg.xtol_abs && std::fabs(xcur[j] - xprev[j]) >= cfg.xtol_abs[j]) all = false;
        }
        *dnorm = d; *xnorm = s; *below = all;
        if (k > 1)
            port_sigma_update(cfg.variant == kMMA ? PORT_MMA : PORT_CCSAQ, n_, xcur.data(), xprev.data(),
                              xprevprev.data(), lb.data(), ub.data(), sigma_min, sigma.data());
        xprevprev = xprev;
        xprev = xcur;
        return true;
    }
    bool fetch_x(double *out) override { std::memcpy(out, x.data(), n_ * sizeof(double)); return true; }
    const std::string &error() const override { return err; }
    double seconds_in_callbacks() const override { return cb; }

private:
    BackendConfig cfg;
    unsigned n_, m_;
    std::vector<double> x, lb, ub, xcur, xprev, xprevprev, sigma, g, gcur, G, Gcur, zeros, gc_tmp;
    std::string err;
    double cb = 0;
};
}  // namespace

Backend *make_backend(const BackendConfig &cfg, std::string *err)
{
    if (!cfg.x0_host) { if (err) *err = "host test backend needs a host start point"; return nullptr; }
    return new OracleBackend(cfg);
}

}  // namespace nb200
