/* oracle/ccsa_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT (see ccsa_port.h).
 *
 * Sequential plain-C restatement of the MMA / CCSAQ path of NLopt 2.11.0.
 * Every routine cites the reference lines it restates.  Operation order inside
 * each floating-point expression is kept identical to the reference so that,
 * built with -ffp-contract=off like the reference (CMakeLists.txt:281-284), the
 * port is bit-identical to oracle/_ref (checked in tests/test_oracle_port.py).
 *
 * Parity: PINNED against SURVEY.md Appendix B known answers and oracle/_ref.
 */
#include "ccsa_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#define RHO_FLOOR 1e-5 /* MMA_RHOMIN / CCSA_RHOMIN, mma.c:41, ccsa_quadratic.c:58 */

/* The n-term sums of the dual functions are accumulated in `acc_t`.  The default build uses double, like the
 * reference (bit-identical to oracle/_ref).  -DPORT_WIDE_SUMS (tools/noise_floor_c3.py only) accumulates in the x87
 * 80-bit format: the same per-variable terms, a nearly exact sum -- it measures how far the reference's own
 * sequential double summation moves a run at large n (the noise floor end-to-end comparisons are judged against). */
#ifdef PORT_WIDE_SUMS
typedef long double acc_t;
#define PORT_MAX_WIDE_M 64
#else
typedef double acc_t;
#endif

/* ------------------------------------------------------------------ */
/* small predicates (stop.c:219-228, :254-263)                         */

int port_isinf(double x)
{
    return fabs(x) >= HUGE_VAL * 0.99 || isinf(x);
}

static int is_nan(double x) { return isnan(x); }

static double now_seconds(void)
{
    /* timer.c:41-52: seconds since the first call */
    static __thread int inited = 0;
    static __thread struct timeval t0;
    struct timeval t;
    if (!inited) { inited = 1; gettimeofday(&t0, NULL); }
    gettimeofday(&t, NULL);
    return (t.tv_sec - t0.tv_sec) + 1.e-6 * (t.tv_usec - t0.tv_usec);
}

/* ------------------------------------------------------------------ */
/* dual evaluation, MMA flavour -- mma.c:59-137                        */

double port_dual_mma(const port_dual_in *in, const double *y, double *grad, port_dual_out *out)
{
    const unsigned n = in->n, m = in->m;
    unsigned i, j;
    acc_t val, gsum, wsum;
#ifdef PORT_WIDE_SUMS
    acc_t gcw[PORT_MAX_WIDE_M];
#define GC(i) gcw[i]
#else
#define GC(i) out->gc[i]
#endif

    /* mma.c:75-78: a NaN constraint value switches that constraint off */
    val = gsum = in->f0;
    wsum = 0;
    for (i = 0; i < m; ++i) {
        GC(i) = is_nan(in->c0[i]) ? 0 : in->c0[i];
        val += y[i] * (double) GC(i);
    }

    for (j = 0; j < n; ++j) {
        const double xj = in->x[j], sj = in->sigma[j], gj = in->grad_f[j];
        double u, v, s2, dx, xc, dx2, dinv, c;

        if (sj == 0) {              /* mma.c:96-99: fixed variable */
            out->xcur[j] = xj;
            continue;
        }
        /* mma.c:101-106 */
        u = gj;
        v = fabs(gj) * sj + 0.5 * in->rho;
        for (i = 0; i < m; ++i)
            if (!is_nan(in->c0[i])) {
                const double a = in->grad_c[(size_t) i * n + j];
                u += a * y[i];
                v += (fabs(a) * sj + 0.5 * in->rhoc[i]) * y[i];
            }
        /* mma.c:107-108: root of the stationarity quadratic, roundoff-safe form */
        s2 = sj * sj;
        u *= s2;
        {
            double r = u / (v * sj);
            dx = (u / v) / (-1 - sqrt(fabs(1 - r * r)));
        }
        /* mma.c:109-114: box clamp, then 0.9 sigma move limit */
        xc = xj + dx;
        if (xc > in->ub[j]) xc = in->ub[j];
        else if (xc < in->lb[j]) xc = in->lb[j];
        if (xc > xj + 0.9 * sj) xc = xj + 0.9 * sj;
        else if (xc < xj - 0.9 * sj) xc = xj - 0.9 * sj;
        out->xcur[j] = xc;
        dx = xc - xj;

        /* mma.c:117-129 */
        dx2 = dx * dx;
        dinv = 1.0 / (s2 - dx2);
        val += (u * dx + v * dx2) * dinv;
        c = s2 * dx;
        gsum += (gj * c + (fabs(gj) * sj + 0.5 * in->rho) * dx2) * dinv;
        wsum += 0.5 * dx2 * dinv;
        for (i = 0; i < m; ++i)
            if (!is_nan(in->c0[i])) {
                const double a = in->grad_c[(size_t) i * n + j];
                GC(i) += (a * c + (fabs(a) * sj + 0.5 * in->rhoc[i]) * dx2) * dinv;
            }
    }
    out->g0 = (double) gsum;
    out->w = (double) wsum;
#ifdef PORT_WIDE_SUMS
    for (i = 0; i < m; ++i) out->gc[i] = (double) gcw[i];
#endif
    /* mma.c:135-136: we maximise the dual, so hand back the negation */
    if (grad)
        for (i = 0; i < m; ++i) grad[i] = -out->gc[i];
    return -(double) val;
}

/* ------------------------------------------------------------------ */
/* dual evaluation, CCSAQ flavour -- ccsa_quadratic.c:79-148           */

double port_dual_ccsaq(const port_dual_in *in, const double *y, double *grad, port_dual_out *out)
{
    const unsigned n = in->n, m = in->m;
    unsigned i, j;
    acc_t val, gsum, wsum;
#ifdef PORT_WIDE_SUMS
    acc_t gcw[PORT_MAX_WIDE_M];
#endif

    /* ccsa_quadratic.c:95-98 (no NaN handling in this flavour) */
    val = gsum = in->f0;
    wsum = 0;
    for (i = 0; i < m; ++i) {
        GC(i) = in->c0[i];
        val += y[i] * (double) GC(i);
    }

    for (j = 0; j < n; ++j) {
        const double xj = in->x[j], sj = in->sigma[j], gj = in->grad_f[j];
        double u, v, s2, dx, xc, dx2, q;

        if (sj == 0) {              /* ccsa_quadratic.c:111-114 */
            out->xcur[j] = xj;
            continue;
        }
        /* ccsa_quadratic.c:116-122 */
        u = in->rho;
        v = gj;
        for (i = 0; i < m; ++i) {
            u += in->rhoc[i] * y[i];
            v += in->grad_c[(size_t) i * n + j] * y[i];
        }
        s2 = sj * sj;
        dx = -s2 * v / u;
        /* ccsa_quadratic.c:126-130: trust clamp then box clamp */
        if (fabs(dx) > sj) dx = copysign(sj, dx);
        xc = xj + dx;
        if (xc > in->ub[j]) xc = in->ub[j];
        else if (xc < in->lb[j]) xc = in->lb[j];
        out->xcur[j] = xc;
        dx = xc - xj;

        /* ccsa_quadratic.c:133-140 */
        dx2 = dx * dx;
        val += v * dx + 0.5 * u * dx2 / s2;
        q = 0.5 * dx2 / s2;
        gsum += gj * dx + in->rho * q;
        wsum += q;
        for (i = 0; i < m; ++i)
            GC(i) += in->grad_c[(size_t) i * n + j] * dx + in->rhoc[i] * q;
    }
    out->g0 = (double) gsum;
    out->w = (double) wsum;
#ifdef PORT_WIDE_SUMS
    for (i = 0; i < m; ++i) out->gc[i] = (double) gcw[i];
#endif
    if (grad)
        for (i = 0; i < m; ++i) grad[i] = -out->gc[i];
    return -(double) val;
}
#undef GC

/* ------------------------------------------------------------------ */
/* sigma handling -- mma.c:202-210 and :431-442, ccsa_quadratic.c:324-332, :577-590 */

void port_sigma_init(unsigned n, const double *lb, const double *ub,
                     const double *sigma_init, double sigma_min, double *sigma)
{
    unsigned j;
    for (j = 0; j < n; ++j) {
        double s;
        if (sigma_init && sigma_init[j] > 0) s = sigma_init[j];
        else if (port_isinf(ub[j]) || port_isinf(lb[j])) s = 1.0;
        else s = 0.5 * (ub[j] - lb[j]);
        sigma[j] = s > sigma_min ? s : sigma_min;
    }
}

void port_sigma_update(int variant, unsigned n, const double *xcur, const double *xprev,
                       const double *xprevprev, const double *lb, const double *ub,
                       double sigma_min, double *sigma)
{
    /* lower clamp factor: 0.01 in mma.c:439, 1e-8 in ccsa_quadratic.c:587 */
    const double kappa = variant == PORT_MMA ? 0.01 : 1e-8;
    unsigned j;
    for (j = 0; j < n; ++j) {
        double s = sigma[j];
        const double osc = (xcur[j] - xprev[j]) * (xprev[j] - xprevprev[j]);
        const double gam = osc < 0 ? 0.7 : (osc > 0 ? 1.2 : 1);
        s *= gam;
        if (!port_isinf(ub[j]) && !port_isinf(lb[j])) {
            const double hi = 10 * (ub[j] - lb[j]), lo = kappa * (ub[j] - lb[j]);
            s = s < hi ? s : hi;
            s = s > lo ? s : lo;
        }
        sigma[j] = s > sigma_min ? s : sigma_min;
    }
}

/* ------------------------------------------------------------------ */
/* stopping predicates -- stop.c:81-108                                */

int port_relstop(double vold, double vnew, double reltol, double abstol)
{
    if (port_isinf(vold)) return 0;
    return fabs(vnew - vold) < abstol
        || fabs(vnew - vold) < reltol * (fabs(vnew) + fabs(vold)) * 0.5
        || (reltol > 0 && vnew == vold);
}

int port_stop_x(unsigned n, const double *x, const double *oldx, const double *w,
                double xtol_rel, const double *xtol_abs)
{
    double dnorm = 0, xnorm = 0;
    unsigned i;
    if (w) {
        for (i = 0; i < n; ++i) dnorm += w[i] * fabs(x[i] - oldx[i]);
        for (i = 0; i < n; ++i) xnorm += w[i] * fabs(x[i]);
    } else {
        for (i = 0; i < n; ++i) dnorm += fabs(x[i] - oldx[i]);
        for (i = 0; i < n; ++i) xnorm += fabs(x[i]);
    }
    if (dnorm < xtol_rel * xnorm) return 1;
    if (!xtol_abs) return 0;
    for (i = 0; i < n; ++i)
        if (fabs(x[i] - oldx[i]) >= xtol_abs[i]) return 0;
    return 1;
}

/* ------------------------------------------------------------------ */
/* the solver                                                          */

void port_default_options(port_options *o)
{
    memset(o, 0, sizeof *o);
    o->stopval = -HUGE_VAL;
    o->rho_init = 1.0;
    o->inner_gradients = 1;
    o->always_improve = 1;
    o->dual_ftol_rel = 1e-14;
    o->dual_maxeval = 100000;
}

/* stop bundle of one level (nlopt_stopping, nlopt-util.h:79-91) */
typedef struct {
    unsigned n;
    double minf_max, ftol_rel, ftol_abs, xtol_rel;
    const double *xtol_abs, *x_weights;
    int nevals, maxeval;
    double maxtime, start;
    int *force_stop;
} stop_t;

static int forced(const stop_t *s) { return s->force_stop && *s->force_stop; }
static int evals_out(const stop_t *s) { return s->maxeval > 0 && s->nevals >= s->maxeval; }
static int time_out(const stop_t *s) { return s->maxtime > 0 && now_seconds() - s->start >= s->maxtime; }

/* closure turning one level's dual evaluation into the next level's objective */
typedef struct {
    int variant;
    port_dual_in in;
    port_dual_out out;
    long count;
} dual_closure;

static double dual_objective(unsigned m, const double *y, double *grad, void *p)
{
    dual_closure *dc = (dual_closure *) p;
    (void) m;
    dc->count++;
    return dc->variant == PORT_MMA ? port_dual_mma(&dc->in, y, grad, &dc->out)
                                   : port_dual_ccsaq(&dc->in, y, grad, &dc->out);
}

typedef struct {
    int inner_maxeval, inner_gradients, always_improve;
    double rho_init, sigma_min;
    const double *sigma_init;
    /* configuration of the dual optimiser one level down (optimize.c:822-826) */
    double dual_ftol_rel, dual_ftol_abs, dual_xtol_rel, dual_xtol_abs;
    int dual_maxeval;
} params_t;

static int ccsa_level(int variant, unsigned n, port_func f, void *f_data,
                      unsigned m, const port_func *fc, void *const *fc_data, const double *tol,
                      const double *lb, const double *ub, double *x, double *minf,
                      stop_t *stop, const params_t *prm, port_stats *stats);

/* what nlopt_optimize_ does around the algorithm (optimize.c:514-566): trivial n,
 * bound sanity, fresh eval counter and clock */
static int enter_level(int variant, unsigned n, port_func f, void *f_data,
                       unsigned m, const port_func *fc, void *const *fc_data, const double *tol,
                       const double *lb, const double *ub, double *x, double *minf,
                       stop_t *stop, const params_t *prm, port_stats *stats)
{
    unsigned i;
    if (n == 0) {                   /* optimize.c:536-539 */
        *minf = f(0, x, NULL, f_data);
        return PORT_SUCCESS;
    }
    *minf = HUGE_VAL;
    for (i = 0; i < n; ++i)         /* optimize.c:547-551 */
        if (lb[i] > ub[i] || x[i] < lb[i] || x[i] > ub[i]) return PORT_INVALID_ARGS;
    stop->n = n;
    stop->nevals = 0;
    stop->start = now_seconds();
    return ccsa_level(variant, n, f, f_data, m, fc, fc_data, tol, lb, ub, x, minf, stop, prm, stats);
}

/* mma.c:145-452 / ccsa_quadratic.c:211-606 (no preconditioner) */
static int ccsa_level(int variant, unsigned n, port_func f, void *f_data,
                      unsigned m, const port_func *fc, void *const *fc_data, const double *tol,
                      const double *lb, const double *ub, double *x, double *minf,
                      stop_t *stop, const params_t *prm, port_stats *stats)
{
    const int is_mma = variant == PORT_MMA;
    int ret = PORT_SUCCESS, feasible;
    unsigned i, k = 0;
    double rho, fcur, infeas;
    dual_closure dd;
    size_t nn = n, mm = m;
    double *work = (double *) malloc(sizeof(double) * (6 * nn + 2 * mm * nn + 8 * mm + 1));
    double *sigma, *g, *g_cur, *xcur, *xprev, *xprevprev, *G, *G_cur;
    double *c, *c_cur, *rhoc, *gc, *ylo, *yhi, *y, *ytol;
    if (!work) return PORT_OUT_OF_MEMORY;
    sigma = work; g = sigma + nn; g_cur = g + nn; xcur = g_cur + nn; xprev = xcur + nn;
    xprevprev = xprev + nn; c = xprevprev + nn; c_cur = c + mm; rhoc = c_cur + mm; gc = rhoc + mm;
    ylo = gc + mm; yhi = ylo + mm; y = yhi + mm; ytol = y + mm; G = ytol + mm; G_cur = G + mm * nn;

    dd.variant = variant;
    dd.in.n = n; dd.in.m = m;
    dd.in.x = x; dd.in.lb = lb; dd.in.ub = ub; dd.in.sigma = sigma;
    dd.in.grad_f = g; dd.in.grad_c = G; dd.in.c0 = c; dd.in.rhoc = rhoc;
    dd.out.xcur = xcur; dd.out.gc = gc;
    dd.count = 0;

    port_sigma_init(n, lb, ub, prm->sigma_init, prm->sigma_min, sigma);   /* mma.c:202-210 */
    rho = prm->rho_init;
    for (i = 0; i < m; ++i) {                                              /* mma.c:211-216 */
        rhoc[i] = prm->rho_init;
        ylo[i] = y[i] = 0.0;
        yhi[i] = HUGE_VAL;
        ytol[i] = prm->dual_xtol_abs;
    }

    /* mma.c:218-229: first evaluation of everything, with gradients */
    dd.in.f0 = fcur = *minf = f(n, x, g, f_data);
    stop->nevals++;
    memcpy(xcur, x, sizeof(double) * nn);
    if (forced(stop)) { ret = PORT_FORCED_STOP; goto done; }
    feasible = 1; infeas = 0;
    for (i = 0; i < m; ++i) {
        c[i] = fc[i](n, x, G + i * nn, fc_data[i]);
        if (forced(stop)) { ret = PORT_FORCED_STOP; goto done; }
    }
    for (i = 0; i < m; ++i) {                                              /* mma.c:230-233 */
        feasible = feasible && (c[i] <= 0 || (is_mma && is_nan(c[i])));
        if (c[i] > infeas) infeas = c[i];
    }
    if (!feasible)                                                         /* mma.c:245-246 */
        for (i = 0; i < m; ++i) yhi[i] = 1e40;

    for (;;) {                      /* outer iterations, mma.c:255 */
        int inner_nevals = 0;
        double fprev = fcur;
        if (forced(stop)) ret = PORT_FORCED_STOP;
        else if (evals_out(stop)) ret = PORT_MAXEVAL_REACHED;
        else if (time_out(stop)) ret = PORT_MAXTIME_REACHED;
        else if (feasible && *minf < stop->minf_max) ret = PORT_STOPVAL_REACHED;
        if (ret != PORT_SUCCESS) goto done;
        if (++k > 1) memcpy(xprevprev, xprev, sizeof(double) * nn);
        memcpy(xprev, xcur, sizeof(double) * nn);
        if (stats) stats->outer_iters++;

        for (;;) {                  /* inner iterations, mma.c:267 */
            double min_dual, infeas_cur;
            int feasible_cur, inner_done, new_infeasible = 0, reti;

            /* --- dual solve (mma.c:275-288) --- */
            dd.in.rho = rho;
            dd.count = 0;
            {
                /* dual_opt as configured in optimize.c:818-826 plus mma.c:248-253, run through
                 * nlopt_optimize_limited (optimize.c:1087-1113) with the time that is left */
                stop_t ds;
                params_t dp;
                double tleft = stop->maxtime - (now_seconds() - stop->start);
                memset(&ds, 0, sizeof ds);
                ds.minf_max = -HUGE_VAL;
                ds.ftol_rel = prm->dual_ftol_rel;
                ds.ftol_abs = prm->dual_ftol_abs;
                ds.xtol_rel = prm->dual_xtol_rel;
                ds.xtol_abs = ytol;
                ds.maxeval = prm->dual_maxeval;
                ds.maxtime = tleft;             /* dual_opt's own maxtime is 0, so the limit is taken as is */
                ds.force_stop = NULL;
                memset(&dp, 0, sizeof dp);
                dp.inner_gradients = 1; dp.always_improve = 1; dp.rho_init = 1.0;
                dp.dual_ftol_rel = 1e-14; dp.dual_maxeval = 100000;  /* never used: next level has n = 0 */
                reti = enter_level(PORT_MMA, m, dual_objective, &dd, 0, NULL, NULL, NULL,
                                   ylo, yhi, y, &min_dual, &ds, &dp, NULL);
            }
            if (reti < 0 || reti == PORT_MAXTIME_REACHED) { ret = reti; goto done; }
            dual_objective(m, y, NULL, &dd);    /* mma.c:288: final x*(y), g, w */
            if (stats) {
                if (stats->inner_iters < 64) stats->dual_count_log[stats->inner_iters] = dd.count;
                stats->inner_iters++;
                stats->dual_evals += dd.count;
            }

            /* --- candidate evaluation (mma.c:297-326) --- */
            fcur = f(n, xcur, prm->inner_gradients ? g_cur : NULL, f_data);
            stop->nevals++;
            ++inner_nevals;
            if (forced(stop)) { ret = PORT_FORCED_STOP; goto done; }
            feasible_cur = 1; infeas_cur = 0;
            inner_done = dd.out.g0 >= fcur;
            for (i = 0; i < m; ++i) {
                c_cur[i] = fc[i](n, xcur, prm->inner_gradients ? G_cur + i * nn : NULL, fc_data[i]);
                if (forced(stop)) { ret = PORT_FORCED_STOP; goto done; }
            }
            for (i = 0; i < m; ++i) {
                if (is_mma && is_nan(c_cur[i])) continue;           /* mma.c:315 */
                feasible_cur = feasible_cur && c_cur[i] <= tol[i];
                if (!is_mma || !is_nan(c[i]))
                    inner_done = inner_done && gc[i] >= c_cur[i];
                else if (c_cur[i] > 0)
                    new_infeasible = 1;                             /* mma.c:321-322 */
                if (c_cur[i] > infeas_cur) infeas_cur = c_cur[i];
            }
            inner_done = inner_done || (prm->inner_maxeval > 0 && inner_nevals == prm->inner_maxeval);

            /* --- acceptance (mma.c:334-392) --- */
            if (prm->always_improve
                    ? ((fcur < *minf && (inner_done || feasible_cur || !feasible))
                       || (!feasible && infeas_cur < infeas))
                    : inner_done) {
                if (!prm->inner_gradients) {    /* mma.c:339-370: gradients now, nevals untouched */
                    fcur = f(n, xcur, g_cur, f_data);
                    if (forced(stop)) { ret = PORT_FORCED_STOP; goto done; }
                    feasible_cur = 1; infeas_cur = 0; new_infeasible = 0;
                    if (is_mma) inner_done = dd.out.g0 >= fcur;     /* mma.c:346 only */
                    for (i = 0; i < m; ++i) {
                        c_cur[i] = fc[i](n, xcur, G_cur + i * nn, fc_data[i]);
                        if (forced(stop)) { ret = PORT_FORCED_STOP; goto done; }
                    }
                    for (i = 0; i < m; ++i) {
                        if (is_mma && is_nan(c_cur[i])) continue;
                        feasible_cur = feasible_cur && c_cur[i] <= tol[i];
                        if (is_mma && is_nan(c[i]) && c_cur[i] > 0) new_infeasible = 1;
                        if (c_cur[i] > infeas_cur) infeas_cur = c_cur[i];
                    }
                }
                dd.in.f0 = *minf = fcur;
                infeas = infeas_cur;
                memcpy(c, c_cur, sizeof(double) * mm);
                memcpy(x, xcur, sizeof(double) * nn);
                memcpy(g, g_cur, sizeof(double) * nn);
                memcpy(G, G_cur, sizeof(double) * nn * mm);
                if (infeas_cur == 0) {          /* mma.c:384-390 */
                    if (!feasible)
                        for (i = 0; i < m; ++i) yhi[i] = HUGE_VAL;
                    feasible = 1;
                } else if (new_infeasible)
                    feasible = 0;
            }
            if (forced(stop)) ret = PORT_FORCED_STOP;
            else if (evals_out(stop)) ret = PORT_MAXEVAL_REACHED;
            else if (time_out(stop)) ret = PORT_MAXTIME_REACHED;
            else if (feasible && *minf < stop->minf_max) ret = PORT_STOPVAL_REACHED;
            if (ret != PORT_SUCCESS) goto done;

            if (inner_done) break;

            /* --- raise the penalties where the approximant was not conservative (mma.c:403-410) --- */
            if (fcur > dd.out.g0) {
                double a = 10 * rho, b = 1.1 * (rho + (fcur - dd.out.g0) / dd.out.w);
                rho = a < b ? a : b;
            }
            for (i = 0; i < m; ++i)
                if (!(is_mma && is_nan(c_cur[i])) && c_cur[i] > gc[i]) {
                    double a = 10 * rhoc[i], b = 1.1 * (rhoc[i] + (c_cur[i] - gc[i]) / dd.out.w);
                    rhoc[i] = a < b ? a : b;
                }
        }

        /* mma.c:418-422: x test overrides f test */
        if (port_relstop(fprev, fcur, stop->ftol_rel, stop->ftol_abs)) ret = PORT_FTOL_REACHED;
        if (port_stop_x(n, xcur, xprev, stop->x_weights, stop->xtol_rel, stop->xtol_abs))
            ret = PORT_XTOL_REACHED;
        if (ret != PORT_SUCCESS) goto done;

        /* mma.c:425-446 */
        rho = 0.1 * rho > RHO_FLOOR ? 0.1 * rho : RHO_FLOOR;
        for (i = 0; i < m; ++i) rhoc[i] = 0.1 * rhoc[i] > RHO_FLOOR ? 0.1 * rhoc[i] : RHO_FLOOR;
        if (k > 1)
            port_sigma_update(variant, n, xcur, xprev, xprevprev, lb, ub, prm->sigma_min, sigma);
    }

done:
    free(work);
    return ret;
}

int port_ccsa_minimize(int variant, unsigned n, port_func f, void *f_data,
                       unsigned m, const port_func *fc, void *const *fc_data, const double *tol,
                       const double *lb, const double *ub, double *x, double *minf,
                       const port_options *opt, port_stats *stats)
{
    stop_t st;
    params_t prm;
    int ret;
    if (!f || !x || !minf || !opt) return PORT_INVALID_ARGS;
    /* optimize.c:807-814 */
    if (!(opt->rho_init > 0) && !port_isinf(opt->rho_init)) return PORT_INVALID_ARGS;
    if ((opt->inner_gradients != 0 && opt->inner_gradients != 1)
        || (opt->always_improve != 0 && opt->always_improve != 1) || opt->sigma_min < 0.0)
        return PORT_INVALID_ARGS;
    memset(&st, 0, sizeof st);
    st.minf_max = opt->stopval;
    st.ftol_rel = opt->ftol_rel; st.ftol_abs = opt->ftol_abs;
    st.xtol_rel = opt->xtol_rel; st.xtol_abs = opt->xtol_abs; st.x_weights = opt->x_weights;
    st.maxeval = opt->maxeval; st.maxtime = opt->maxtime;
    st.force_stop = opt->force_stop;
    prm.inner_maxeval = opt->inner_maxeval;
    prm.inner_gradients = opt->inner_gradients;
    prm.always_improve = opt->always_improve;
    prm.rho_init = opt->rho_init;
    prm.sigma_min = opt->sigma_min;
    prm.sigma_init = opt->sigma_init;
    prm.dual_ftol_rel = opt->dual_ftol_rel; prm.dual_ftol_abs = opt->dual_ftol_abs;
    prm.dual_xtol_rel = opt->dual_xtol_rel; prm.dual_xtol_abs = opt->dual_xtol_abs;
    prm.dual_maxeval = opt->dual_maxeval;
    if (stats) memset(stats, 0, sizeof *stats);
    ret = enter_level(variant, n, f, f_data, m, fc, fc_data, tol, lb, ub, x, minf, &st, &prm, stats);
    if (stats) stats->numevals = st.nevals;
    return ret;
}
