// problems.cu -> libnlopt_b200_problems.so : the benchmark / test problems of BASELINE.json as
// USER code of the library (it only uses the public headers).  Device-resident versions are
// written as __device__ functors through include/nlopt_b200_device.cuh; host versions are plain
// nlopt_func callbacks usable with any library exporting the NLopt ABI (ours or the reference).
//
//   chained Rosenbrock  (formula of reference test/testfuncs.c:124-139)   -- config 3 objective
//   dense linear inequality  c(x) = w.x - b  with a caller-supplied weight row -- config 3 constraints
//   separable quadratic 1/2 sum a_j (x_j - b_j)^2, a, b from the counter hash  -- config 2 objective
//   mean constraint  sum x / n + offset                                         -- config 2 / 4 constraint
//
// Per-variable expressions use un-fused IEEE operations in the same order as tests/problems.py
// (numpy), so device gradients are bit-identical to the host callbacks' gradients.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/nlopt_b200_device.cuh"
#include "synth.cuh"

namespace {

double g_cb_seconds = 0.0;
struct Tick {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Tick() { g_cb_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// ---- device functors ---------------------------------------------------------------------------------
struct RosenbrockDev {
    static constexpr int halo = 1;           // reads x[jl - 1] and x[jl + 1] across shard boundaries
    __device__ double operator()(unsigned long long j, unsigned long long n, long long jl, long long,
                                 const double *x, double *grad_j) const
    {
        const double xj = x[jl];
        double term = 0.0, gsum = 0.0;
        if (j + 1 < n) {
            const double d = __dsub_rn(x[jl + 1], __dmul_rn(xj, xj)), e = __dsub_rn(1.0, xj);
            term = __dadd_rn(__dmul_rn(__dmul_rn(100.0, d), d), __dmul_rn(e, e));
            gsum = __dadd_rn(0.0, __dsub_rn(__dmul_rn(__dmul_rn(-400.0, xj), d), __dmul_rn(2.0, e)));
        }
        if (j > 0) {
            const double xm = x[jl - 1];
            gsum = __dadd_rn(gsum, __dmul_rn(200.0, __dsub_rn(xj, __dmul_rn(xm, xm))));
        }
        if (grad_j) *grad_j = gsum;
        return term;
    }
    double finish(double s) const { return s; }
};

struct LinearDev {
    const double *w;        // device, this rank's shard of the weight row
    double b;
    __device__ double operator()(unsigned long long, unsigned long long, long long jl, long long, const double *x,
                                 double *grad_j) const
    {
        const double wj = w[jl];
        if (grad_j) *grad_j = wj;
        return __dmul_rn(wj, x[jl]);
    }
    double finish(double s) const { return s - b; }
};

struct QuadraticDev {
    unsigned long long seed;
    __device__ double operator()(unsigned long long j, unsigned long long, long long jl, long long, const double *x,
                                 double *grad_j) const
    {
        const double a = __dadd_rn(1.0, nb200::u01(seed, 0, j));
        const double b = __dsub_rn(__dmul_rn(2.0, nb200::u01(seed, 1, j)), 1.0);
        const double d = __dsub_rn(x[jl], b);
        const double ad = __dmul_rn(a, d);
        if (grad_j) *grad_j = ad;
        return __dmul_rn(ad, d);
    }
    double finish(double s) const { return 0.5 * s; }
};

// synthetic SIMP compliance (BASELINE config 4): f(x) = sum_j a_j / (eps + (1 - eps) x_j^3),
// a_j = 0.5 + u01(seed, 0, j).  Same expression order as nb200p_simp_host below.
struct SimpDev {
    unsigned long long seed;
    double eps;
    __device__ double operator()(unsigned long long j, unsigned long long, long long jl, long long, const double *x,
                                 double *grad_j) const
    {
        const double a = __dadd_rn(0.5, nb200::u01(seed, 0, j));
        const double xj = x[jl], x2 = __dmul_rn(xj, xj), x3 = __dmul_rn(x2, xj);
        const double ome = __dsub_rn(1.0, eps);
        const double d = __dadd_rn(eps, __dmul_rn(ome, x3));
        if (grad_j) *grad_j = -__ddiv_rn(__dmul_rn(__dmul_rn(a, __dmul_rn(ome, 3.0)), x2), __dmul_rn(d, d));
        return __ddiv_rn(a, d);
    }
    double finish(double s) const { return s; }
};

struct MeanDev {
    double inv_n, offset;
    __device__ double operator()(unsigned long long, unsigned long long, long long jl, long long, const double *x,
                                 double *grad_j) const
    {
        if (grad_j) *grad_j = inv_n;
        return x[jl];
    }
    double finish(double s) const { return s * inv_n + offset; }
};

}  // namespace

struct nb200p_lin_data {
    const double *w;
    double b;
};
struct nb200p_quad_data {
    unsigned long long seed;
};
struct nb200p_mean_data {
    double offset;
};
struct nb200p_simp_data {
    unsigned long long seed;
    double eps;
};

struct nb200p_problem_s {
    RosenbrockDev rosen;
    QuadraticDev quad;
    std::vector<LinearDev *> lin;
    std::vector<MeanDev *> mean;
    std::vector<double *> dev_rows;
    std::vector<nb200p_lin_data *> lin_host;
    SimpDev simp;
    std::vector<void *> misc_host;          // small data records of the host callbacks (freed with the problem)
};

extern "C" {

nb200p_problem_s *nb200p_create(void) { return new nb200p_problem_s; }

void nb200p_destroy(nb200p_problem_s *p)
{
    if (!p) return;
    for (double *d : p->dev_rows) cudaFree(d);
    for (LinearDev *l : p->lin) delete l;
    for (MeanDev *m : p->mean) delete m;
    for (nb200p_lin_data *l : p->lin_host) delete l;
    for (void *q : p->misc_host) std::free(q);
    delete p;
}

double nb200p_callback_seconds(void) { return g_cb_seconds; }
void nb200p_reset_callback_seconds(void) { g_cb_seconds = 0.0; }

// ---- device registration -------------------------------------------------------------------------------
int nb200p_set_rosenbrock_device(nb200p_problem_s *p, nlopt_opt opt)
{
    return nlopt_b200::set_min_objective(opt, &p->rosen);
}

int nb200p_add_linear_device(nb200p_problem_s *p, nlopt_opt opt, const double *w_host_full, double b, double tol)
{
    const unsigned n = nlopt_get_dimension(opt);
    unsigned long long j0 = 0, cnt = n;
    nlopt_b200_shard_range(n, nlopt_b200_comm_rank(), nlopt_b200_comm_world(), &j0, &cnt);
    double *w = nullptr;
    if (cudaMalloc(&w, (cnt ? cnt : 1) * sizeof(double)) != cudaSuccess) return NLOPT_OUT_OF_MEMORY;
    cudaMemcpy(w, w_host_full + j0, cnt * sizeof(double), cudaMemcpyHostToDevice);
    p->dev_rows.push_back(w);
    LinearDev *l = new LinearDev{w, b};
    p->lin.push_back(l);
    return nlopt_b200::add_inequality_constraint(opt, l, tol);
}

int nb200p_set_quadratic_device(nb200p_problem_s *p, nlopt_opt opt, unsigned long long seed)
{
    p->quad.seed = seed;
    return nlopt_b200::set_min_objective(opt, &p->quad);
}

int nb200p_add_mean_device(nb200p_problem_s *p, nlopt_opt opt, double offset, double tol)
{
    MeanDev *m = new MeanDev{1.0 / (double) nlopt_get_dimension(opt), offset};
    p->mean.push_back(m);
    return nlopt_b200::add_inequality_constraint(opt, m, tol);
}

int nb200p_set_simp_device(nb200p_problem_s *p, nlopt_opt opt, unsigned long long seed, double eps)
{
    p->simp.seed = seed;
    p->simp.eps = eps;
    return nlopt_b200::set_min_objective(opt, &p->simp);
}

// ---- host callbacks (nlopt_func shape; work with any NLopt-ABI library) ----------------------------------
double nb200p_simp_host(unsigned n, const double *x, double *grad, void *data)
{
    Tick t;
    const nb200p_simp_data *sd = static_cast<const nb200p_simp_data *>(data);
    const double ome = 1.0 - sd->eps;
    double f = 0.0;
    for (unsigned j = 0; j < n; ++j) {
        const double a = 0.5 + nb200::u01(sd->seed, 0, j);
        const double x2 = x[j] * x[j], x3 = x2 * x[j];
        const double d = sd->eps + ome * x3;
        if (grad) grad[j] = -(((a * (ome * 3.0)) * x2) / (d * d));
        f += a / d;
    }
    return f;
}

// sharded forms (nlopt_b200_sfunc): this rank's variables only, additive value contribution
double nb200p_simp_sharded(unsigned n_local, unsigned long long j0, unsigned long long, const double *x, double *grad, void *data)
{
    Tick t;
    const nb200p_simp_data *sd = static_cast<const nb200p_simp_data *>(data);
    const double ome = 1.0 - sd->eps;
    double f = 0.0;
    for (unsigned jl = 0; jl < n_local; ++jl) {
        const double a = 0.5 + nb200::u01(sd->seed, 0, j0 + jl);
        const double x2 = x[jl] * x[jl], x3 = x2 * x[jl];
        const double d = sd->eps + ome * x3;
        if (grad) grad[jl] = -(((a * (ome * 3.0)) * x2) / (d * d));
        f += a / d;
    }
    return f;
}

double nb200p_mean_sharded(unsigned n_local, unsigned long long j0, unsigned long long n, const double *x, double *grad, void *data)
{
    Tick t;
    const double inv_n = 1.0 / (double) n;
    double s = 0.0;
    for (unsigned jl = 0; jl < n_local; ++jl) s += x[jl];
    if (grad)
        for (unsigned jl = 0; jl < n_local; ++jl) grad[jl] = inv_n;
    return s * inv_n + (j0 == 0 ? static_cast<const nb200p_mean_data *>(data)->offset : 0.0);
}

void *nb200p_make_simp_data(nb200p_problem_s *p, unsigned long long seed, double eps)
{
    nb200p_simp_data *d = static_cast<nb200p_simp_data *>(std::malloc(sizeof(nb200p_simp_data)));
    d->seed = seed;
    d->eps = eps;
    p->misc_host.push_back(d);
    return d;
}

void *nb200p_make_mean_data(nb200p_problem_s *p, double offset)
{
    nb200p_mean_data *d = static_cast<nb200p_mean_data *>(std::malloc(sizeof(nb200p_mean_data)));
    d->offset = offset;
    p->misc_host.push_back(d);
    return d;
}

void *nb200p_make_quad_data(nb200p_problem_s *p, unsigned long long seed)
{
    nb200p_quad_data *d = static_cast<nb200p_quad_data *>(std::malloc(sizeof(nb200p_quad_data)));
    d->seed = seed;
    p->misc_host.push_back(d);
    return d;
}

double nb200p_rosenbrock_host(unsigned n, const double *x, double *grad, void *)
{
    Tick t;
    double f = 0.0;
    if (grad)
        for (unsigned j = 0; j < n; ++j) grad[j] = 0.0;
    for (unsigned j = 0; j + 1 < n; ++j) {
        const double d = x[j + 1] - x[j] * x[j], e = 1.0 - x[j];
        f += 100.0 * d * d + e * e;
        if (grad) {
            grad[j] += -400.0 * x[j] * d - 2.0 * e;
            grad[j + 1] += 200.0 * d;
        }
    }
    return f;
}

double nb200p_linear_host(unsigned n, const double *x, double *grad, void *data)
{
    Tick t;
    const nb200p_lin_data *d = static_cast<const nb200p_lin_data *>(data);
    double s = 0.0;
    for (unsigned j = 0; j < n; ++j) s += d->w[j] * x[j];
    if (grad) std::memcpy(grad, d->w, (size_t) n * sizeof(double));
    return s - d->b;
}

double nb200p_quadratic_host(unsigned n, const double *x, double *grad, void *data)
{
    Tick t;
    const unsigned long long seed = static_cast<const nb200p_quad_data *>(data)->seed;
    double s = 0.0;
    for (unsigned j = 0; j < n; ++j) {
        const double a = 1.0 + nb200::u01(seed, 0, j), b = 2.0 * nb200::u01(seed, 1, j) - 1.0;
        const double d = x[j] - b, ad = a * d;
        if (grad) grad[j] = ad;
        s += ad * d;
    }
    return 0.5 * s;
}

double nb200p_mean_host(unsigned n, const double *x, double *grad, void *data)
{
    Tick t;
    const double inv_n = 1.0 / (double) n;
    double s = 0.0;
    for (unsigned j = 0; j < n; ++j) s += x[j];
    if (grad)
        for (unsigned j = 0; j < n; ++j) grad[j] = inv_n;
    return s * inv_n + static_cast<const nb200p_mean_data *>(data)->offset;
}

// data-record helpers for the host callbacks (w_host must stay alive)
void *nb200p_make_linear_data(nb200p_problem_s *p, const double *w_host, double b)
{
    nb200p_lin_data *d = new nb200p_lin_data{w_host, b};
    p->lin_host.push_back(d);
    return d;
}

}  // extern "C"
