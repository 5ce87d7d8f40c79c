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
//       // n_local the shard length, n the global length.  Neighbours x[jl-1], x[jl+1] may be read
//       // when they are inside the shard (stencil functions need world size 1 or their own halo).
//       __device__ double operator()(unsigned long long j, unsigned long long n, long long jl,
//                                    long long n_local, const double *x, double *grad_j) const;
//       // optional constant / scaling applied once to the global sum on the host
//       double finish(double sum) const { return sum; }
//   };
//
// Usage:   nlopt_b200::set_min_objective(opt, &functor);     // functor must outlive opt
//          nlopt_b200::add_inequality_constraint(opt, &cfunctor, tol);
//
// The reduction is deterministic: a fixed grid, a fixed thread->variable map and a fixed
// shuffle / shared-memory / last-block tree, all in un-fused IEEE double adds.
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
    auto *b = new detail::Bound<F>{f, nlopt_get_dimension(opt)};      // lives as long as the process
    return nlopt_b200_set_min_objective_device(opt, &detail::trampoline<F>, b);
}

template <class F>
nlopt_result add_inequality_constraint(nlopt_opt opt, const F *f, double tol)
{
    auto *b = new detail::Bound<F>{f, nlopt_get_dimension(opt)};
    return nlopt_b200_add_inequality_constraint_device(opt, &detail::trampoline<F>, b, tol);
}

}  // namespace nlopt_b200
