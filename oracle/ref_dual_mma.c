/* oracle/ref_dual_mma.c -- TEST INFRASTRUCTURE (builds into oracle/_ref/libref_dual.so).
 *
 * Reaches the reference's *static* MMA dual function (src/algs/mma/mma.c:59-137)
 * without copying it: the reference translation unit is #included from where it
 * lies (the -I path in oracle/Makefile points at $(REF)/src/algs/mma) and a
 * flat C wrapper is exported next to it.  Nothing below is reference text.
 */
#include "mma.c"

__attribute__((visibility("default")))
double ref_mma_dual_eval(unsigned n, unsigned m, const double *y, double *grad,
                         const double *x, const double *lb, const double *ub,
                         const double *sigma, const double *dfdx, const double *dfcdx,
                         double fval, double rho, const double *fcval, const double *rhoc,
                         double *xcur, double *gcval, double *gval_wval)
{
    dual_data dd;
    double r;
    dd.count = 0; dd.n = n;
    dd.x = x; dd.lb = lb; dd.ub = ub; dd.sigma = sigma; dd.dfdx = dfdx; dd.dfcdx = dfcdx;
    dd.fval = fval; dd.rho = rho; dd.fcval = fcval; dd.rhoc = rhoc;
    dd.xcur = xcur; dd.gcval = gcval;
    r = dual_func(m, y, grad, &dd);
    gval_wval[0] = dd.gval;
    gval_wval[1] = dd.wval;
    return r;
}
