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
    void *data = nullptr;
};

struct BackendConfig {
    Variant variant = kMMA;
    unsigned n = 0;                          // global problem size
    FuncSpec objective;
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
