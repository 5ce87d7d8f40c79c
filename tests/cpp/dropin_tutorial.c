/* dropin_tutorial.c -- a plain C user program written against <nlopt.h> only (the names, enum values and
 * signatures of the reference's src/api/nlopt.h).  The GPU test compiles it with `-lnlopt` against the
 * libnlopt.so.1 build of this repository and runs it on the B200: the tutorial problem of the reference's
 * documentation (doc/docs/NLopt_Tutorial.md; same problem as test/t_tutorial.cxx) for the algorithm id given
 * on the command line (24 = LD_MMA, 41 = LD_CCSAQ, 31 = LD_AUGLAG over the default MMA).
 * Exit status 0 iff |f* - sqrt(8/27)| < 1e-3, the pin of test/t_tutorial.cxx:76. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <nlopt.h>

typedef struct { double a, b; } cdata;
static int count = 0;

static double objective(unsigned n, const double *x, double *grad, void *data)
{
    (void) n; (void) data;
    ++count;
    if (grad) { grad[0] = 0.0; grad[1] = 0.5 / sqrt(x[1]); }
    return sqrt(x[1]);
}

static double constraint(unsigned n, const double *x, double *grad, void *data)
{
    const cdata *d = (const cdata *) data;
    const double t = d->a * x[0] + d->b;
    (void) n;
    if (grad) { grad[0] = 3 * d->a * t * t; grad[1] = -1.0; }
    return t * t * t - x[1];
}

int main(int argc, char **argv)
{
    const nlopt_algorithm alg = argc > 1 ? (nlopt_algorithm) atoi(argv[1]) : NLOPT_LD_MMA;
    const double exactmin = 0.544331053951817355154952;
    double lb[2] = {-HUGE_VAL, 1e-6}, x[2] = {1.234, 5.678}, minf = 0.0;
    cdata data[2] = {{2, 0}, {-1, 1}};
    int major, minor, bugfix;
    nlopt_result r;
    nlopt_opt opt = nlopt_create(alg, 2);
    if (!opt) { fprintf(stderr, "nlopt_create failed\n"); return 2; }
    nlopt_version(&major, &minor, &bugfix);
    nlopt_set_lower_bounds(opt, lb);
    nlopt_set_min_objective(opt, objective, NULL);
    nlopt_add_inequality_constraint(opt, constraint, &data[0], 1e-8);
    nlopt_add_inequality_constraint(opt, constraint, &data[1], 1e-8);
    nlopt_set_xtol_rel(opt, 1e-4);
    if (nlopt_set_param(opt, "inner_maxeval", 123) != NLOPT_SUCCESS || nlopt_get_param(opt, "inner_maxeval", 0) != 123) return 3;
    r = nlopt_optimize(opt, x, &minf);
    if (r < 0) {
        fprintf(stderr, "nlopt_optimize failed: %d (%s)\n", (int) r, nlopt_get_errmsg(opt) ? nlopt_get_errmsg(opt) : "");
        nlopt_destroy(opt);
        return 1;
    }
    printf("%s (library %d.%d.%d) found minimum at f(%g,%g) = %.10g after %d evaluations, result %d\n",
           nlopt_algorithm_name(alg), major, minor, bugfix, x[0], x[1], minf, count, (int) r);
    nlopt_destroy(opt);
    return fabs(minf - exactmin) < 1e-3 ? 0 : 1;
}
