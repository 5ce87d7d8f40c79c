// nlopt_b200_device.cuh -- supply the objective / constraints as __device__ code.
//
// NLopt's callbacks (src/api/nlopt.h:60-62) are host functions reading x and writing the gradient
// in host memory.  For the B200 path that means one D2H of x and one H2D per gradient row per
// inner iteration.  This header removes the trip: a user functor with a __device__ operator() is
// instantiated into a map + fixed-tree-reduce kernel in the USER's translation unit and registered
// through the plain C entry points nlopt_b200_set_min_objective_device /
// nlopt_b200_add_inequality_constraint_device (include/nlopt_b200.h).
//
// Functor concept (separable-sum functions  F(x) = finish( sum_j term_j )):
//
//   struct MyF {
//       // value contribution of variable j and d F / d x_j (write iff grad_j != nullptr).
//       // x points at this rank's shard; jl is the index inside it, j = j0 + jl the global index,
//       // n_local the shard length, n the global length.  A functor that declares
//       //     static constexpr int halo = 1;
//       // may also read x[jl-1] and x[jl+1] for every variable whose global neighbour exists (j > 0, j + 1 < n):
//       // with several ranks the library fills the cells x[-1] and x[n_local] from the neighbouring ranks.
//       __device__ double operator()(unsigned long long j, unsigned long long n, long long jl,
//                                    long long n_local, const double *x, double *grad_j) const;
//       // optional constant / scaling applied once to the global sum on the host
//       double finish(double sum) const { return sum; }
//   };
//
// Usage:   nlopt_b200::set_min_objective(opt, &functor);     // functor must outlive opt
//          nlopt_b200::add_inequality_constraint(opt, &cfunctor, tol);
//
// The reduction is deterministic AND independent of the number of ranks: the variables are cut into the library's
// groups and 8 virtual shards (a function of n alone, nlopt_b200_shard_geometry); one CTA reduces one group with a
// fixed thread->variable map and a fixed shuffle / shared-memory tree, a second kernel folds the group sums of each
// virtual shard in a fixed order, and the library adds the 8 shard sums of all ranks in index order -- all in un-fused
// IEEE double adds.  Nothing synchronises the host per function: the callbacks of a point are enqueued back to back
// and the library collects all values with one copy (nlopt_b200_dfunc2, include/nlopt_b200.h).
#pragma once

#include <cuda_runtime.h>

#include "nlopt_b200.h"

namespace nlopt_b200 {

namespace detail {

constexpr int kThreads = 256;
constexpr int kBlocks = 1184;            // 8 CTAs per SM on a 148-SM B200

struct Workspace {
    double *partials = nullptr;          // [kBlocks]
    unsigned *ticket = nullptr;
    double *result_host = nullptr;       // pinned
    double *result_dev = nullptr;
};

inline Workspace &workspace()
{
    static Workspace w;
    if (!w.partials) {
        cudaMalloc(&w.partials, kBlocks * sizeof(double));
        cudaMalloc(&w.ticket, sizeof(unsigned));
        cudaMemset(w.ticket, 0, sizeof(unsigned));
        cudaMalloc(&w.result_dev, sizeof(double));
        cudaHostAlloc(&w.result_host, sizeof(double), cudaHostAllocDefault);
    }
    return w;
}

__device__ __forceinline__ double block_sum(double v, double *smem)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) s = __dadd_rn(s, smem[w]);
    __syncthreads();
    return s;                            // valid in thread 0
}

template <class F>
__global__ void __launch_bounds__(kThreads) map_reduce_kernel(F f, unsigned long long j0, unsigned long long n,
                                                              long long n_local, const double *x, double *grad,
                                                              double *partials, unsigned *ticket, double *result)
{
    __shared__ double smem[kThreads / 32];
    __shared__ int last;
    // contiguous chunk per block, so the summation order is a function of n_local only
    const long long per = (n_local + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long) blockIdx.x * per;
    long long hi = lo + per;
    if (hi > n_local) hi = n_local;
    double acc = 0.0;
    for (long long jl = lo + threadIdx.x; jl < hi; jl += kThreads)
        acc = __dadd_rn(acc, f(j0 + (unsigned long long) jl, n, jl, n_local, x, grad ? grad + jl : nullptr));
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    acc = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) acc = __dadd_rn(acc, __ldcg(partials + b));
    const double total = block_sum(acc, smem);
    if (threadIdx.x == 0) {
        *result = total;
        *ticket = 0;
    }
}

template <class F>
double evaluate(const F &f, unsigned n_local, unsigned long long j0, unsigned long long n, const double *x_dev,
                double *grad_dev, cudaStream_t s)
{
    Workspace &w = workspace();
    map_reduce_kernel<F><<<kBlocks, kThreads, 0, s>>>(f, j0, n, (long long) n_local, x_dev, grad_dev, w.partials,
                                                      w.ticket, w.result_dev);
    cudaMemcpyAsync(w.result_host, w.result_dev, sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return *w.result_host;
}

// ---- asynchronous, rank-count-independent form (nlopt_b200_dfunc2) ---------------------------------------------
struct Workspace2 {
    double *partials = nullptr;
    unsigned cap = 0;
};
inline double *partials2(unsigned groups)
{
    static Workspace2 w;
    if (groups > w.cap) {
        if (w.partials) cudaFree(w.partials);
        w.cap = groups + 64;
        cudaMalloc(&w.partials, (size_t) w.cap * sizeof(double));
    }
    return w.partials;
}

// one CTA per group: thread t takes variables lo + t, lo + t + 256, ... of the group
template <class F>
__global__ void __launch_bounds__(kThreads) map_group_kernel(F f, nlopt_b200_shard sh, const double *x, double *grad, double *partials)
{
    __shared__ double smem[kThreads / 32];
    const unsigned g = sh.group0 + blockIdx.x;
    const unsigned long long c_lo = (unsigned long long) g * sh.nchunks / sh.groups_total - sh.chunk0;
    const unsigned long long c_hi = (unsigned long long) (g + 1) * sh.nchunks / sh.groups_total - sh.chunk0;
    long long lo = (long long) (c_lo * 512), hi = (long long) (c_hi * 512);
    if (hi > (long long) sh.n_local) hi = (long long) sh.n_local;
    double acc = 0.0;
    for (long long jl = lo + threadIdx.x; jl < hi; jl += kThreads)
        acc = __dadd_rn(acc, f(sh.j0 + (unsigned long long) jl, sh.n, jl, (long long) sh.n_local, x, grad ? grad + jl : nullptr));
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// one CTA per local virtual shard: its P group sums in a fixed order
__global__ void __launch_bounds__(kThreads) fold_groups_kernel(const double *partials, unsigned P, double *vsums /* at vshard0 */)
{
    __shared__ double smem[kThreads / 32];
    const double *base = partials + (size_t) blockIdx.x * P;
    double acc = 0.0;
    for (unsigned r = threadIdx.x; r < P; r += kThreads) acc = __dadd_rn(acc, base[r]);
    const double s = block_sum(acc, smem);
    if (threadIdx.x == 0) vsums[blockIdx.x] = s;
}

template <class F, class = void>
struct halo_of { static constexpr int value = 0; };
template <class F>
struct halo_of<F, decltype((void) F::halo)> { static constexpr int value = F::halo; };

template <class F>
void trampoline2(const nlopt_b200_shard *sh, const double *x_dev, double *grad_dev, double *vsums_dev, void *data, void *stream)
{
    const F *f = static_cast<const F *>(data);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (sh->groups_local == 0) return;
    double *part = partials2(sh->groups_local);
    map_group_kernel<F><<<sh->groups_local, kThreads, 0, s>>>(*f, *sh, x_dev, grad_dev, part);
    fold_groups_kernel<<<sh->local_vshards, kThreads, 0, s>>>(part, sh->groups_per_vshard, vsums_dev + sh->vshard0);
}

template <class F>
double finish2(double total, void *data)
{
    return static_cast<const F *>(data)->finish(total);
}

template <class F>
struct Bound {
    const F *f;
    unsigned long long n;
};

template <class F>
double trampoline(unsigned n_local, unsigned long long j0, const double *x_dev, double *grad_dev, void *data,
                  void *stream)
{
    const Bound<F> *b = static_cast<const Bound<F> *>(data);
    const double sum = evaluate(*b->f, n_local, j0, b->n, x_dev, grad_dev, static_cast<cudaStream_t>(stream));
    // the constant of finish() must enter the cross-rank sum exactly once: rank owning j = 0 adds it
    return j0 == 0 ? b->f->finish(sum) : b->f->finish(sum) - b->f->finish(0.0);
}

}  // namespace detail

// `f` (host object holding the functor's parameters) must stay alive and unchanged while `opt` uses it.
template <class F>
nlopt_result set_min_objective(nlopt_opt opt, const F *f)
{
    return nlopt_b200_set_min_objective_device2(opt, &detail::trampoline2<F>, &detail::finish2<F>, const_cast<F *>(f),
                                                detail::halo_of<F>::value);
}

template <class F>
nlopt_result add_inequality_constraint(nlopt_opt opt, const F *f, double tol)
{
    return nlopt_b200_add_inequality_constraint_device2(opt, &detail::trampoline2<F>, &detail::finish2<F>, const_cast<F *>(f), tol,
                                                        detail::halo_of<F>::value);
}

// the first form of the interface (one synchronous evaluation per call, nlopt_b200_dfunc), kept for callers that
// want a value right away: rank-local sums, summed over ranks by the library
template <class F>
nlopt_result set_min_objective_sync(nlopt_opt opt, const F *f)
{
    auto *b = new detail::Bound<F>{f, nlopt_get_dimension(opt)};      // lives as long as the process
    return nlopt_b200_set_min_objective_device(opt, &detail::trampoline<F>, b);
}

template <class F>
nlopt_result add_inequality_constraint_sync(nlopt_opt opt, const F *f, double tol)
{
    auto *b = new detail::Bound<F>{f, nlopt_get_dimension(opt)};
    return nlopt_b200_add_inequality_constraint_device(opt, &detail::trampoline<F>, b, tol);
}

}  // namespace nlopt_b200
