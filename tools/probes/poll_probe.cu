// Micro-probe behind the design of the persistent solve kernel's folder CTA: what do nanosleep, strong / weak
// 16-byte loads and a cross-SM tagged-slot hand-off actually cost on this GPU?   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ ulonglong2 ld_vol(const ulonglong2 *p) { ulonglong2 v; asm volatile("ld.volatile.global.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ ulonglong2 ld_gpu(const ulonglong2 *p) { ulonglong2 v; asm volatile("ld.relaxed.gpu.global.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ ulonglong2 ld_cg(const ulonglong2 *p) { ulonglong2 v; asm volatile("ld.global.cg.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_gpu(ulonglong2 *p, unsigned long long a, unsigned long long b) { asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }

template <int MODE> __device__ __forceinline__ ulonglong2 ldm(const ulonglong2 *p) { return MODE == 0 ? ld_vol(p) : MODE == 1 ? ld_gpu(p) : ld_cg(p); }

__global__ void sleep_probe(unsigned long long *out)
{
    const unsigned ns[6] = {20, 40, 100, 250, 1000, 4000};
    for (int i = 0; i < 6; ++i) {
        unsigned long long t0 = gt();
        for (int k = 0; k < 200; ++k) __nanosleep(ns[i]);
        out[i] = (gt() - t0) / 200;
    }
    unsigned long long t0 = gt(), c0 = clock64();
    while (gt() - t0 < 100000) {}
    out[6] = clock64() - c0;          // cycles per 100 us
    unsigned long long a = gt(), b;   // globaltimer resolution
    do { b = gt(); } while (b == a);
    out[7] = b - a;
}

// every thread of the block: `reps` polls of `nload` independent slots stride `stride` (in 16-byte units)
template <int MODE>
__global__ void load_probe(const ulonglong2 *buf, int nload, int stride, int lane_stride, int reps, unsigned long long *out)
{
    const ulonglong2 *p = buf + (size_t) threadIdx.x * lane_stride;
    unsigned long long acc = 0;
    __syncthreads();
    unsigned long long t0 = gt();
    for (int r = 0; r < reps; ++r) {
        ulonglong2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < nload) v[k] = ldm<MODE>(p + (size_t) k * stride);
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < nload) acc += v[k].x + v[k].y;
        p += (acc & 1);   // dependency between rounds
    }
    unsigned long long t1 = gt();
    __syncthreads();
    if (threadIdx.x == 0) { out[0] = (t1 - t0) / reps; out[1] = acc; }
}

// ping-pong between CTA 0 and CTA 1 (different SMs) through tagged slots; half round trip = hand-off latency
template <int MODE>
__global__ void pingpong(ulonglong2 *slots, int rounds, int sleep_ns, unsigned long long *out)
{
    if (threadIdx.x != 0) return;
    ulonglong2 *mine = slots + 64 * blockIdx.x, *other = slots + 64 * (1 - blockIdx.x);
    unsigned long long t0 = gt();
    for (unsigned long long r = 1; r <= (unsigned long long) rounds; ++r) {
        if (blockIdx.x == 0) st_gpu(other, r, r);
        for (;;) { ulonglong2 v = ldm<MODE>(mine); if (v.y == r) break; if (sleep_ns) __nanosleep(sleep_ns); }
        if (blockIdx.x == 1) st_gpu(other, r, r);
    }
    if (blockIdx.x == 0) out[0] = (gt() - t0) / rounds / 2;
}

int main()
{
    unsigned long long *out; cudaMallocManaged(&out, 64 * 8);
    ulonglong2 *buf; cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 0, 64 << 20);
    sleep_probe<<<1, 1>>>(out); cudaDeviceSynchronize();
    printf("nanosleep(20,40,100,250,1000,4000) actual ns: %llu %llu %llu %llu %llu %llu | SM clock %.0f MHz | globaltimer step %llu ns\n",
           out[0], out[1], out[2], out[3], out[4], out[5], out[6] / 100.0, out[7]);
    const char *names[3] = {"volatile(sys)", "relaxed.gpu", "weak .cg"};
    for (int threads : {1, 32, 256}) for (int nload : {1, 7}) for (int ls : {1, 8}) {
        unsigned long long r[3];
        load_probe<0><<<1, threads>>>(buf, nload, 1224, ls, 200, out); cudaDeviceSynchronize(); r[0] = out[0];
        load_probe<1><<<1, threads>>>(buf, nload, 1224, ls, 200, out); cudaDeviceSynchronize(); r[1] = out[0];
        load_probe<2><<<1, threads>>>(buf, nload, 1224, ls, 200, out); cudaDeviceSynchronize(); r[2] = out[0];
        printf("poll round: %3d threads x %d slots, lane stride %3d B : %s %llu ns | %s %llu ns | %s %llu ns\n", threads, nload, ls * 16,
               names[0], r[0], names[1], r[1], names[2], r[2]);
    }
    for (int sl : {0, 20, 100}) {
        unsigned long long r[3];
        cudaMemset(buf, 0, 4096); pingpong<0><<<2, 32>>>(buf, 2000, sl, out); cudaDeviceSynchronize(); r[0] = out[0];
        cudaMemset(buf, 0, 4096); pingpong<1><<<2, 32>>>(buf, 2000, sl, out); cudaDeviceSynchronize(); r[1] = out[0];
        cudaMemset(buf, 0, 4096); pingpong<2><<<2, 32>>>(buf, 2000, sl, out); cudaDeviceSynchronize(); r[2] = out[0];
        printf("slot hand-off SM->SM, nanosleep(%d) in the poll loop: %s %llu ns | %s %llu ns | %s %llu ns\n", sl, names[0], r[0], names[1], r[1], names[2], r[2]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
