// nlopt_mini.hpp -- a small header-only C++ wrapper over the C ABI, written for the drop-in test
// (tests/test_dropin_gpu.py).  It drives the library the way the reference's generated wrapper does
// (src/api/nlopt-in.hpp): std::vector arguments, exceptions for negative result codes (:93-104), functor
// data owned by the nlopt_opt through nlopt_set_munge (free on destroy, duplicate on copy: :122-146, :266),
// C++ callbacks that receive std::vector and whose exceptions become forced stops (:149-166).
#pragma once

#include <nlopt.h>

#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

namespace nlopt_mini {

typedef std::function<double(const std::vector<double> &x, std::vector<double> &grad)> vfunc;

class opt {
    struct slot {
        opt *owner;
        vfunc f;
        std::vector<double> x, g;
    };
    nlopt_opt o_;
    std::string pending_;          // what() of an exception thrown inside a callback

    static void *free_slot(void *p) { delete static_cast<slot *>(p); return nullptr; }
    static void *dup_slot(void *p) { return p ? new slot(*static_cast<slot *>(p)) : nullptr; }
    static double thunk(unsigned n, const double *x, double *grad, void *data)
    {
        slot *s = static_cast<slot *>(data);
        try {
            s->x.assign(x, x + n);
            s->g.assign(grad ? n : 0, 0.0);
            const double v = s->f(s->x, s->g);
            if (grad) for (unsigned i = 0; i < n; ++i) grad[i] = s->g[i];
            return v;
        } catch (const std::exception &e) {
            s->owner->pending_ = e.what();
            nlopt_force_stop(s->owner->o_);
            return HUGE_VAL;
        }
    }
    void check(nlopt_result r) const
    {
        if (r >= 0) return;
        const char *msg = nlopt_get_errmsg(o_);
        if (r == NLOPT_INVALID_ARGS) throw std::invalid_argument(msg ? msg : "nlopt invalid argument");
        if (r == NLOPT_OUT_OF_MEMORY) throw std::bad_alloc();
        throw std::runtime_error(msg ? msg : "nlopt failure");
    }

public:
    opt(nlopt_algorithm a, unsigned n) : o_(nlopt_create(a, n))
    {
        if (!o_) throw std::bad_alloc();
        nlopt_set_munge(o_, free_slot, dup_slot);
    }
    opt(const opt &other) : o_(nlopt_copy(other.o_))
    {
        if (!o_) throw std::bad_alloc();
    }
    ~opt() { nlopt_destroy(o_); }
    opt &operator=(const opt &) = delete;

    void set_min_objective(vfunc f) { check(nlopt_set_min_objective(o_, thunk, new slot{this, std::move(f), {}, {}})); }
    void add_inequality_constraint(vfunc f, double tol) { check(nlopt_add_inequality_constraint(o_, thunk, new slot{this, std::move(f), {}, {}}, tol)); }
    void remove_inequality_constraints() { check(nlopt_remove_inequality_constraints(o_)); }
    void set_lower_bounds(const std::vector<double> &lb) { check(nlopt_set_lower_bounds(o_, lb.data())); }
    void set_upper_bounds(const std::vector<double> &ub) { check(nlopt_set_upper_bounds(o_, ub.data())); }
    void set_xtol_rel(double t) { check(nlopt_set_xtol_rel(o_, t)); }
    void set_stopval(double v) { check(nlopt_set_stopval(o_, v)); }
    void set_maxeval(int n) { check(nlopt_set_maxeval(o_, n)); }
    void set_param(const char *name, double v) { check(nlopt_set_param(o_, name, v)); }
    double get_param(const char *name, double dflt) const { return nlopt_get_param(o_, name, dflt); }
    unsigned get_dimension() const { return nlopt_get_dimension(o_); }
    int get_numevals() const { return nlopt_get_numevals(o_); }
    const char *get_algorithm_name() const { return nlopt_algorithm_name(nlopt_get_algorithm(o_)); }

    nlopt_result optimize(std::vector<double> &x, double &minf)
    {
        if (x.size() != get_dimension()) throw std::invalid_argument("dimension mismatch");
        pending_.clear();
        nlopt_set_force_stop(o_, 0);
        const nlopt_result r = nlopt_optimize(o_, x.data(), &minf);
        if (r == NLOPT_FORCED_STOP && !pending_.empty()) throw std::runtime_error(pending_);
        check(r);
        return r;
    }
};

}  // namespace nlopt_mini
