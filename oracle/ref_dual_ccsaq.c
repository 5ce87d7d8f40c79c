/* oracle/ref_dual_ccsaq.c -- TEST INFRASTRUCTURE (builds into oracle/_ref/libref_dual.so).
 *
 * Same include-trick as ref_dual_mma.c, for the CCSAQ dual function
 * (src/algs/mma/ccsa_quadratic.c:79-148).  Separate TU because both reference
 * files define `dual_data` / `dual_func`.
 */
#include "ccsa_quadratic.c"

__attribute__((visibility("default")))
double ref_ccsaq_dual_eval(unsigned n, unsigned m, const double *y, double *grad,
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
    dd.pre = NULL; dd.pre_data = NULL; dd.prec = NULL; dd.prec_data = NULL; dd.scratch = NULL;
    r = dual_func(m, y, grad, &dd);
    gval_wval[0] = dd.gval;
    gval_wval[1] = dd.wval;
    return r;
}
