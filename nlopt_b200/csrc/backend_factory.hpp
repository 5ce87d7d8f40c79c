// backend_factory.hpp -- how the API layer obtains the n-dimensional state holder.
// The product links device_backend.cu (CUDA, sm_100a).  Nothing else implements this in
// libnlopt_b200.so; if no CUDA device is usable make_backend fails with a message and
// nlopt_optimize returns NLOPT_FAILURE -- there is deliberately no CPU path.
#pragma once

#include <string>
#include <vector>

#include "../../include/nlopt_b200.h"
#include "backend.hpp"

namespace nb200 {

// one user function: exactly one of f / mf / df is set
struct FuncSpec {
    unsigned m = 1;
    nlopt_func f = nullptr;
    nlopt_mfunc mf = nullptr;
    nlopt_b200_dfunc df = nullptr;
    nlopt_b200_dfunc2 df2 = nullptr;         // asynchronous device callback (takes precedence over df) ...
    nlopt_b200_dfinish dfin = nullptr;       // ... and its host-side finish
    int halo = 0;
    nlopt_b200_sfunc sf = nullptr;           // sharded host callback
    void *data = nullptr;
};

// The augmented-Lagrangian objective of NLOPT_AUGLAG* (src/algs/auglag/auglag.c:25-65), evaluated by the backend:
//   L(x) = f(x) + rho/2 sum_k (h_k(x) + lambda_k/rho)^2 + rho/2 sum_k max(0, c_k(x) + mu_k/rho)^2
// and, where a gradient is wanted, grad L = grad f + sum_k coef_k grad(h_k | c_k) accumulated in the reference's
// order.  The caller owns the multipliers and changes them (and rho) between sub-optimisations.
struct PenaltySpec {
    std::vector<FuncSpec> eq, ineq;          // constraint objects folded into the objective
    double rho = 1.0;
    const double *lambda = nullptr;          // one per scalar equality constraint
    const double *mu = nullptr;              // one per scalar inequality constraint
    int *nevals_p = nullptr;                 // the outer object's evaluation counter (auglag.c:38)
    const int *force_stop = nullptr;         // the outer object's force-stop flag (auglag.c:39)
};

struct BackendConfig {
    Variant variant = kMMA;
    unsigned n = 0;                          // global problem size
    FuncSpec objective;
    const PenaltySpec *penalty = nullptr;    // non-null: `objective` is f of the augmented Lagrangian above
    std::vector<FuncSpec> constraints;       // inequality constraint objects, in registration order
    const double *lb = nullptr, *ub = nullptr;   // host, n entries
    bool lb_uniform = false, ub_uniform = false; // all entries equal lb[0] / ub[0]: fill on the device, no H2D
    const double *x0_host = nullptr;         // host start point (n entries) ...
    double *x_dev = nullptr;                 // ... or this rank's device shard (device mode, in/out)
    const double *sigma_init = nullptr;      // nlopt initial step (host) or null
    const double *x_weights = nullptr;       // host or null
    const double *xtol_abs = nullptr;        // host or null
    nlopt_b200_stats *stats = nullptr;       // h2d/d2h bytes, launches, kernel time
};

Backend *make_backend(const BackendConfig &cfg, std::string *err);

}  // namespace nb200
