// ccsa_driver.hpp -- host-side CCSA outer/inner loop (scalars only), see ccsa_driver.cpp.
#pragma once

#include <cmath>
#include <string>
#include <vector>

#include "backend.hpp"

namespace nb200 {

// nlopt_stopping (src/util/nlopt-util.h:79-91) minus the O(n) arrays, which live in the backend
struct StopCriteria {
    double minf_max = -HUGE_VAL;
    double ftol_rel = 0, ftol_abs = 0, xtol_rel = 0;
    bool has_xtol_abs = false;
    int maxeval = 0;
    double maxtime = 0;
    double start = -1;         // wall_seconds() when nlopt_optimize was entered (optimize.c:1001: stop.start is
                               // taken before any algorithm setup); < 0: the driver takes it itself
    const int *force_stop = nullptr;
    int *nevals_p = nullptr;
};

// algorithm parameters as read in src/api/optimize.c:798-826
struct CcsaParams {
    int inner_maxeval = 0;
    int verbosity = 0;
    double rho_init = 1.0;
    int inner_gradients = 1;
    int always_improve = 1;
    double sigma_min = 0.0;
    double dual_ftol_rel = 1e-14, dual_ftol_abs = 0, dual_xtol_rel = 0, dual_xtol_abs = 0;
    int dual_maxeval = 100000;
    int fused_solve = 1;     // library knob b200_fused_solve: run each dual solve as one persistent kernel
};

struct DriverStats {
    long long dual_evals = 0, dual_solves = 0, outer_iters = 0;
    double seconds_dual = 0, seconds_eval = 0, seconds_glue = 0;   // wall-clock breakdown of the loop
};

// Runs NLOPT_LD_MMA / NLOPT_LD_CCSAQ on the state held by `be`.  `tol` has one feasibility
// tolerance per scalar constraint (be.m() entries).  Returns an nlopt_result value; on
// return *minf is the best objective value and the backend's x the corresponding point.
int ccsa_minimize(Variant variant, Backend &be, const std::vector<double> &tol, double *minf,
                  const StopCriteria &stop, const CcsaParams &prm, DriverStats *stats, std::string *errmsg);

}  // namespace nb200
