/* oracle/ccsa_port.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Plain-C CPU restatement ("port") of the one NLopt path this repository
 * accelerates: the CCSA family NLOPT_LD_MMA / NLOPT_LD_CCSAQ
 * (reference: src/algs/mma/mma.c, src/algs/mma/ccsa_quadratic.c,
 * src/util/stop.c, src/api/optimize.c:795-834).  It exists only so that tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg can CHECK the CUDA
 * path; nothing under nlopt_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle_port.py checks this port
 *   (a) against the known-answer vectors of SURVEY.md Appendix B (measured from the
 *       reference build), and
 *   (b) bit-for-bit against the unmodified reference compiled into oracle/_ref/
 *       (same operation order, both built with -ffp-contract=off).
 */
#ifndef CCSA_PORT_H
#define CCSA_PORT_H

#ifdef __cplusplus
extern "C" {
#endif

/* same shape as the reference's nlopt_func (src/api/nlopt.h:60-62) */
typedef double (*port_func)(unsigned n, const double *x, double *grad, void *data);

enum { PORT_MMA = 0, PORT_CCSAQ = 1 };

/* result codes: numerically identical to nlopt_result (src/api/nlopt.h:162-176) */
enum {
    PORT_FAILURE = -1, PORT_INVALID_ARGS = -2, PORT_OUT_OF_MEMORY = -3,
    PORT_ROUNDOFF_LIMITED = -4, PORT_FORCED_STOP = -5,
    PORT_SUCCESS = 1, PORT_STOPVAL_REACHED = 2, PORT_FTOL_REACHED = 3,
    PORT_XTOL_REACHED = 4, PORT_MAXEVAL_REACHED = 5, PORT_MAXTIME_REACHED = 6
};

/* All the arrays one dual evaluation reads (reference: dual_data,
 * mma.c:46-55 / ccsa_quadratic.c:63-75).  grad_c is m-by-n, row i = d c_i / d x. */
typedef struct {
    unsigned n, m;
    const double *x, *lb, *ub, *sigma, *grad_f;
    const double *grad_c;          /* [m*n], row-major by constraint        */
    double f0, rho;                /* objective value at x, its penalty     */
    const double *c0, *rhoc;       /* [m] constraint values / penalties     */
} port_dual_in;

typedef struct {
    double *xcur;                  /* [n] minimiser x*(y)                   */
    double *gc;                    /* [m] approximants g_i(x*(y))           */
    double g0, w;                  /* objective approximant, w(x*(y))       */
} port_dual_out;

/* returns -val (the quantity the dual optimiser minimises); grad may be NULL */
double port_dual_mma(const port_dual_in *in, const double *y, double *grad, port_dual_out *out);
double port_dual_ccsaq(const port_dual_in *in, const double *y, double *grad, port_dual_out *out);

/* sigma (asymptote / trust radius) initialisation and per-outer-iteration update */
void port_sigma_init(unsigned n, const double *lb, const double *ub,
                     const double *sigma_init /* may be NULL */, double sigma_min, double *sigma);
void port_sigma_update(int variant, unsigned n, const double *xcur, const double *xprev,
                       const double *xprevprev, const double *lb, const double *ub,
                       double sigma_min, double *sigma);

/* stopping predicates (stop.c:81-108) */
int port_relstop(double vold, double vnew, double reltol, double abstol);
int port_stop_x(unsigned n, const double *x, const double *oldx, const double *w /* may be NULL */,
                double xtol_rel, const double *xtol_abs /* may be NULL */);
int port_isinf(double x);

typedef struct {
    /* stopping criteria of the user-level problem (nlopt_stopping, nlopt-util.h:79-91) */
    double stopval;                /* default -HUGE_VAL */
    double ftol_rel, ftol_abs, xtol_rel;
    const double *xtol_abs;        /* may be NULL */
    const double *x_weights;       /* may be NULL */
    int maxeval;                   /* <= 0: unlimited */
    double maxtime;                /* <= 0: unlimited */
    /* algorithm parameters (optimize.c:798-826) */
    int inner_maxeval;             /* 0 */
    double rho_init;               /* 1.0 */
    int inner_gradients;           /* 1 */
    int always_improve;            /* 1 */
    double sigma_min;              /* 0 */
    const double *sigma_init;      /* nlopt initial step; may be NULL */
    double dual_ftol_rel;          /* 1e-14 */
    double dual_ftol_abs;          /* 0 */
    double dual_xtol_rel;          /* 0 */
    double dual_xtol_abs;          /* 0 */
    int dual_maxeval;              /* 100000 */
    int *force_stop;               /* may be NULL */
} port_options;

typedef struct {
    int numevals;                  /* objective evaluations (nlopt_get_numevals)        */
    long dual_evals;               /* total level-1 dual evaluations                     */
    int inner_iters;               /* number of dual solves                               */
    int outer_iters;
    long dual_count_log[64];       /* dual evaluations of the first 64 dual solves       */
} port_stats;

void port_default_options(port_options *o);

/* The whole solver: outer/inner CCSA loop + the m-dimensional dual optimiser
 * (which is MMA again, one level down) -- mma.c:145-452, ccsa_quadratic.c:211-606
 * with pre == NULL.  m scalar inequality constraints fc[i](x) <= 0 with
 * feasibility tolerances tol[i].  x is in/out.  Returns an nlopt_result value. */
int port_ccsa_minimize(int variant, unsigned n, port_func f, void *f_data,
                       unsigned m, const port_func *fc, void *const *fc_data, const double *tol,
                       const double *lb, const double *ub, double *x, double *minf,
                       const port_options *opt, port_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
