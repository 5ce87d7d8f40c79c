// synth.cuh -- deterministic synthetic instance of one dual evaluation, generated on the device.
// Same counter-based hash and same un-fused floating-point expressions as tests/synth.py
// (kernel_instance), so a test can compare device-generated arrays with host-generated ones bit
// for bit, and the bench can fill n = 1e7..1e8 without pushing gigabytes over PCIe.
#pragma once

#include <cuda_runtime.h>

namespace nb200 {

__host__ __device__ inline unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__host__ __device__ inline double u01(unsigned long long seed, unsigned k, unsigned long long j)
{
    const unsigned long long base = (seed + k) * 0x9E3779B97F4A7C15ull;
    return (double) (mix64(base + j) >> 11) * 0x1.0p-53;
}

struct SynthArgs {
    double *x, *lb, *ub, *sigma, *g, *G;
    unsigned long long ld, n_local, j0, seed;
    int m;
};

__global__ void synth_fill_kernel(SynthArgs a)
{
    for (unsigned long long t = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; t < a.n_local;
         t += (unsigned long long) gridDim.x * blockDim.x) {
        const unsigned long long j = a.j0 + t;
        const double cls = u01(a.seed, 99, j);
        double lb = -2.0, ub = 2.0;
        double sigma = __dmul_rn(__dadd_rn(0.05, __dmul_rn(0.95, u01(a.seed, 0, j))), 2.0);
        const double x = __dadd_rn(lb, __dmul_rn(__dadd_rn(0.25, __dmul_rn(0.5, u01(a.seed, 1, j))), __dsub_rn(ub, lb)));
        if (cls < 0.001) { lb = x; ub = x; sigma = 0.0; }
        else if (cls < 0.002) { lb = -HUGE_VAL; ub = HUGE_VAL; }
        a.x[t] = x;
        a.lb[t] = lb;
        a.ub[t] = ub;
        a.sigma[t] = sigma;
        a.g[t] = __dmul_rn(__dsub_rn(__dmul_rn(2.0, u01(a.seed, 2, j)), 1.0), 10.0);
        for (int i = 0; i < a.m; ++i)
            a.G[(unsigned long long) i * a.ld + t] = __dsub_rn(__dmul_rn(2.0, u01(a.seed, 3 + i, j)), 1.0);
    }
}

}  // namespace nb200
