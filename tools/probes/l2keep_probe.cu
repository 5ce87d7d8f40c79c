// l2keep_probe.cu -- can a streaming kernel keep part of its operands resident in the B200's L2 across passes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l2keep_probe tools/probes/l2keep_probe.cu && build/l2keep_probe
// One "pass" = 444 CTAs x 256 threads stream KEEP MB with policy P_keep and then STREAM MB with policy P_stream
// (ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64).  Passes repeat back to back inside one kernel (grid-stride over
// the same addresses, like the generations of the persistent dual-solve kernel); the time of the last passes is
// reported as effective GB/s over KEEP + STREAM.  Variants: policies (evict_last / evict_first / none) x
// cudaLimitPersistingL2CacheSize (0 / max).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ double2 ldp(const double2 *p, unsigned long long pol, int use) {
    double2 v;
    if (use) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    else asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__global__ void passes(const double2 *keep, size_t nkeep, const double2 *strm, size_t nstrm, int mode, int npass, unsigned long long *tt, double *sink) {
    unsigned long long pk, ps;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pk));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(ps));
    double s = 0;
    const size_t stride = (size_t) gridDim.x * blockDim.x, t0i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    for (int p = 0; p < npass; ++p) {
        if (blockIdx.x == 0 && threadIdx.x == 0) tt[p] = gt();
        for (size_t i = t0i; i < nkeep; i += stride) { double2 v = ldp(keep + i, pk, mode); s += v.x + v.y; }
        for (size_t i = t0i; i < nstrm; i += stride) { double2 v = ldp(strm + i, ps, mode); s += v.x + v.y; }
        // crude grid barrier per pass through a counter would distort; passes of different CTAs simply overlap a little
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) tt[npass] = gt();
    if (s == 1.2345) *sink = s;
}
int main() {
    int dev = 0, maxp = 0, l2 = 0;
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    printf("L2 %d MB, max persisting %d MB\n", l2 >> 20, maxp >> 20);
    double2 *keep, *strm; double *sink; unsigned long long *tt;
    cudaMalloc(&keep, 128ull << 20); cudaMalloc(&strm, 512ull << 20); cudaMemset(keep, 0, 128ull << 20); cudaMemset(strm, 0, 512ull << 20);
    cudaMalloc(&sink, 8); cudaMalloc(&tt, 8 * 64);
    const int npass = 24;
    for (int lim = 0; lim < 2; ++lim) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim ? (size_t) maxp : 0);
        for (int keep_mb : {0, 20, 30, 40, 60, 90})
            for (int total_mb : {90, 180})
                for (int mode = 0; mode < 2; ++mode) {
                    if (keep_mb > total_mb) continue;
                    const size_t nk = ((size_t) keep_mb << 20) / 16, ns = ((size_t) (total_mb - keep_mb) << 20) / 16;
                    passes<<<444, 256>>>(keep, nk, strm, ns, mode, npass, tt, sink);
                    unsigned long long h[64]; cudaMemcpy(h, tt, 8 * (npass + 1), cudaMemcpyDeviceToHost);
                    const double us = (h[npass] - h[npass / 2]) * 1e-3 / (npass - npass / 2);
                    printf("persist-limit %s  working set %3d MB, keep %2d MB, %s : %.2f us per pass  (%.0f GB/s effective)\n", lim ? "max" : "0  ", total_mb, keep_mb,
                           mode ? "keep=evict_last rest=evict_first" : "no policy                       ", us, total_mb * 1048576.0 / us * 1e-3);
                }
    }
    return 0;
}
