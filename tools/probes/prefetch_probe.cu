// prefetch_probe.cu -- does cp.async.bulk.prefetch.L2 make a later streaming read faster?  (B200)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/prefetch_probe tools/probes/prefetch_probe.cu && build/prefetch_probe
// Each of 444 CTAs owns a contiguous slice of `per` bytes of a buffer.  Phase 0: flush the L2 by streaming a 512 MB
// buffer.  Variant A: read the slice (ld.global.nc.L1::no_allocate.v2.f64, 256 threads) and time it with %globaltimer.
// Variant B: one thread issues a bulk L2 prefetch of the slice, the CTA spins `wait_ns`, then reads and times the read.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void flush(const double2 *p, size_t n, double *sink) {
    double s = 0; for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) { double2 v = p[i]; s += v.x + v.y; }
    if (s == 1.2345) *sink = s;
}
__global__ void probe(const double2 *buf, size_t per_pairs, int mode, unsigned long long wait_ns, unsigned long long *tmin, unsigned long long *tmax, double *sink) {
    const double2 *mine = buf + (size_t) blockIdx.x * per_pairs;
    if (mode == 1 && threadIdx.x == 0) {
        // several prefetches of <= 16 KB like the solver issues them
        for (size_t off = 0; off < per_pairs; off += 1024) {
            size_t np = per_pairs - off < 1024 ? per_pairs - off : 1024;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(mine + off), "r"((unsigned) (np * 16)) : "memory");
        }
    }
    if (mode == 2) {   // per-thread prefetch.global.L2 of each 128-byte line
        for (size_t i = threadIdx.x * 8; i < per_pairs; i += blockDim.x * 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(mine + i));
    }
    if (mode) { const unsigned long long t0 = gt(); while (gt() - t0 < wait_ns) __nanosleep(100); }
    __syncthreads();
    const unsigned long long t0 = gt();
    double s = 0;
    for (size_t i = threadIdx.x; i < per_pairs; i += blockDim.x) {
        double2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(mine + i));
        s += v.x + v.y;
    }
    __syncthreads();
    const unsigned long long t1 = gt();
    if (threadIdx.x == 0) { atomicMin(tmin, t0); atomicMax(tmax, t1); }
    if (s == 1.2345) *sink = s;
}
int main() {
    const int ctas = 444;
    double2 *big, *buf; double *sink; unsigned long long *t;
    const size_t bigN = (512ull << 20) / 16;
    cudaMalloc(&big, bigN * 16); cudaMemset(big, 0, bigN * 16);
    cudaMalloc(&sink, 8); cudaMalloc(&t, 16);
    for (size_t kb : {36, 72, 108, 216}) {
        const size_t per_pairs = kb * 1024 / 16;
        cudaMalloc(&buf, per_pairs * 16 * ctas); cudaMemset(buf, 0, per_pairs * 16 * ctas);
        for (int mode = 0; mode < 3; ++mode)
            for (unsigned long long w : {5000ull, 20000ull}) {
                if (mode == 0 && w != 5000ull) continue;
                double best = 1e9;
                for (int rep = 0; rep < 5; ++rep) {
                    flush<<<1184, 256>>>(big, bigN, sink);
                    unsigned long long init[2] = {~0ull, 0ull};
                    cudaMemcpy(t, init, 16, cudaMemcpyHostToDevice);
                    probe<<<ctas, 256>>>(buf, per_pairs, mode, w, t, t + 1, sink);
                    unsigned long long h[2]; cudaMemcpy(h, t, 16, cudaMemcpyDeviceToHost);
                    const double us = (h[1] - h[0]) * 1e-3; if (us < best) best = us;
                }
                printf("slice %zu KB x %d CTAs = %.1f MB  mode %d (0 none, 1 bulk prefetch, 2 per-line prefetch) wait %llu ns : read phase %.2f us  (%.0f GB/s)\n",
                       kb, ctas, kb * ctas / 1024.0, mode, w, best, kb * 1024.0 * ctas / best * 1e-3);
            }
        cudaFree(buf);
    }
    return 0;
}
