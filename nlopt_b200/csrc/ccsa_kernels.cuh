// ccsa_kernels.cuh -- the CUDA kernels of the MMA/CCSAQ hot path (sm_100a).
//
//   dual_eval_kernel     : one dual evaluation  y -> x*(y), val, g0, w, g_1..g_m over this rank's shard
//                          (reference: static dual_func, src/algs/mma/mma.c:59-137 and
//                           src/algs/mma/ccsa_quadratic.c:79-148); dual_eval_tma_kernel: same, operands staged by TMA
//   dual_solve_kernel    : a whole dual solve (mma.c:275-288) in one persistent cooperative launch -- the default path
//   sigma_init_kernel    : mma.c:202-210
//   end_outer_kernel     : nlopt_stop_x norms (src/util/stop.c:98-108) + sigma update (mma.c:431-442,
//                          ccsa_quadratic.c:577-590) + xprev/xprevprev rotation (mma.c:264-265), one pass
//   penalty_axpy_kernel  : gradient of the augmented-Lagrangian objective (src/algs/auglag/auglag.c:47-48, :59-60)
//
// Arithmetic contract: every per-variable expression is evaluated with the reference's operation
// order using __dmul_rn/__dadd_rn/__dsub_rn/__ddiv_rn/__dsqrt_rn, which nvcc never contracts into
// FMAs -- the reference is built with -ffp-contract=off (CMakeLists.txt:281-284).  x*(y) is
// therefore bit-identical to the reference; only the ORDER of the n-term sums differs (fixed
// tree, see below), which is the documented parity tolerance.
//
// Reduction contract (deterministic and independent of the grid and of the number of GPUs): the global
// index space is cut into S = 8*P groups whose boundaries depend on n only (geometry.hpp).  A group is
// reduced with a fixed lane->element map and a fixed shuffle / shared-memory tree into one record of m+3
// sums.  The P records of a "virtual shard" (8 shards; each rank owns 8/world) are folded in the canonical
// shard order (fold_shard_records), the rank's shard sums in index order, and the 8 shard sums of all ranks
// in index order.  Who does the folding differs by kernel -- the warp that completes a shard (atomic
// ticket) in the one-evaluation kernels, a dedicated folder CTA polling tagged records in the solve
// kernel -- the operations and their order do not, so every path gives the same bits.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "dual_mma.hpp"

namespace nb200 {

constexpr int kBlock = 256;              // threads per CTA of every kernel here
constexpr int kWarps = kBlock / 32;
constexpr int kVirtualShards = 8;        // V: fixed, so 1/2/4/8 ranks give bit-identical sums
constexpr int kMaxParamM = 32;           // multipliers that travel as kernel parameters
constexpr int kMaxNV = 3 + 16;           // accumulators one CTA carries: val, g0, w, <=16 g_i

// ---- exact-rounding arithmetic (never fused) ---------------------------------------------
__device__ __forceinline__ double mulx(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double addx(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double subx(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double divx(double a, double b) { return __ddiv_rn(a, b); }

// streaming loads: read-once data must not displace anything in L1
__device__ __forceinline__ double2 ld_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(double2 *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- tagged 128-bit mailbox slots (cross-rank exchange) ---------------------------------------------------
__device__ __forceinline__ void box_put(double *slot, double value, unsigned long long tag)
{
    asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(slot), "l"(__double_as_longlong(value)), "l"(tag) : "memory");
}
__device__ __forceinline__ bool box_get(const double *slot, unsigned long long tag, double *value)
{
    long long v;
    unsigned long long t;
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(slot) : "memory");
    *value = __longlong_as_double(v);
    return t == tag;
}
constexpr int kBoxStride = 24;           // slots per virtual shard record (comm.hpp)

// The same slots inside one GPU (group records and published multipliers of the persistent solve kernel):
// gpu-scope relaxed accesses are served by the L2.  (The .volatile = system-scope forms above, needed across
// NVLink, measured ~9 ns per lane request when 32 lanes hit 32 different lines.)
__device__ __forceinline__ void slot_put(double *slot, double value, unsigned long long tag)
{
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(slot), "l"(__double_as_longlong(value)), "l"(tag) : "memory");
}
// Polling many slots per thread: strong (volatile / relaxed) loads of one thread complete one after the other
// (measured: the folder's tail grew with the number of strong loads per thread, ~0.3 us each on an idle memory
// system), weak loads pipeline.  ld.global.cg always reads the L2 -- where the producers' stores land -- and the
// tag travels in the same 16 bytes as the value, so a slot whose tag matches is complete.
__device__ __forceinline__ bool slot_peek(const double *slot, unsigned long long tag, double *value)
{
    long long v;
    unsigned long long t;
    asm volatile("ld.global.cg.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(slot) : "memory");
    *value = __longlong_as_double(v);
    return t == tag;
}
__device__ __forceinline__ int ld_gpu_s32(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool slot_get(const double *slot, unsigned long long tag, double *value)
{
    long long v;
    unsigned long long t;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(slot) : "memory");
    *value = __longlong_as_double(v);
    return t == tag;
}

// Lane k < nv: write this rank's shard records into every peer's mailbox, then gather all 8 records of
// sum k from the own mailbox and fold them in index order.  Returns the total in lanes < nv; *timed_out is
// warp-uniform.
// VS_LOCAL: the shard sums were written by this CTA (shared memory or own global writes ordered by a barrier): plain loads.
template <bool VS_LOCAL = false>
__device__ __forceinline__ double box_exchange(double *const *box, int rank, int world, unsigned long long seq,
                                               const double *vsums, int nvp, unsigned local_vshards, unsigned v0,
                                               int nv, int lane, int *timed_out)
{
    const int buf = (int) (seq & 1ull);
    double total = 0.0;
    int to = 0;
    if (lane < nv) {
        for (unsigned v = 0; v < local_vshards; ++v) {
            const double val = VS_LOCAL ? vsums[(unsigned long long) v * nvp + lane] : __ldcg(vsums + (unsigned long long) v * nvp + lane);
            for (int r = 0; r < world; ++r)
                box_put(box[r] + 2ull * (((unsigned long long) buf * 8 + v0 + v) * kBoxStride + lane), val, seq);
        }
        const double *mine = box[rank] + 2ull * ((unsigned long long) buf * 8 * kBoxStride + lane);
        const long long t0 = clock64();
        double x[kVirtualShards];
        for (;;) {                         // all 8 slots are fetched together: one L2 round trip per poll
            bool ok[kVirtualShards];
#pragma unroll
            for (int v = 0; v < kVirtualShards; ++v) ok[v] = box_get(mine + 2ull * v * kBoxStride, seq, &x[v]);
            bool all = true;
#pragma unroll
            for (int v = 0; v < kVirtualShards; ++v) all = all && ok[v];
            if (all) break;
            if (clock64() - t0 > 20000000000ll) { to = 1; break; }          // ~10 s: a peer died
        }
        total = x[0];
#pragma unroll
        for (int v = 1; v < kVirtualShards; ++v) total = addx(total, x[v]);
    }
    *timed_out = __any_sync(0xffffffffu, to);
    return total;
}

// ---- arguments of one dual evaluation -------------------------------------------------------
constexpr int kGroupWarps = 8;           // warps that sweep one group together (fixed: part of the reduction order)
constexpr int kChunkPairs = 32 * kGroupWarps;   // 256 double2 pairs = 512 variables = 4 KB per array per sweep step

struct DualArgs {
    // shard-local arrays (16-byte aligned, padded with sigma = 0 lanes)
    const double *x, *lb, *ub, *sigma, *g;
    const double *G;              // m rows of ld doubles
    double *xcur;                 // written iff STORE
    unsigned long long ld;        // row stride of G in doubles
    // group geometry (global, depends on n only): group s covers chunks [s*nchunks/S, (s+1)*nchunks/S)
    unsigned long long nchunks;   // ceil(ceil(n/2) / 256) over ALL ranks
    unsigned long long chunk0;    // first global chunk of this rank
    unsigned nseg_total;          // S = 8 * P
    unsigned seg0;                // first global group of this rank
    unsigned segs_per_vshard;     // P
    unsigned local_vshards;       // 8 / world
    // reduction workspace
    double *grouprecs;            // [local groups][nvp]       group records
    double *vsums;                // [local_vshards][nvp]
    unsigned *tickets;            // [local_vshards + 1], zero between launches
    double *out_dev;              // [8][nvp] all-rank exchange buffer (this rank's slots filled)
    volatile double *out_host;    // mapped pinned [nvp]; written when publish_host
    volatile unsigned long long *flag_host;
    unsigned long long seq;
    int publish_host;             // 1: single rank, results + flag go straight to the host
    int nvp;                      // stride of one record (>= 3 + chunk size)
    // fused cross-rank exchange over NVLink peer memory (null box[0]: NCCL path instead)
    double *box[8];               // box[r]: rank r's mailbox as mapped into this process (CUDA IPC)
    int rank, world;
    // the multipliers and penalties
    int m;                        // total number of constraints (rows of G)
    int cons0, cons_n;            // this launch accumulates g_i for i in [cons0, cons0 + cons_n)
    unsigned active;              // bit i clear: constraint i switched off (MMA, NaN value)
    double rho, half_rho, u_ccsaq;    // u_ccsaq = rho + sum_i rhoc_i y_i (ccsa_quadratic.c:116-120)
    double y[kMaxParamM], rhoc[kMaxParamM], half_rhoc[kMaxParamM];
};

// the per-evaluation scalars as the point functions see them: for the one-evaluation kernel they alias
// the __grid_constant__ parameter block (constant-bank operands), for the persistent solve kernel y lives
// in shared memory and changes every generation
struct Multipliers {
    const double *y, *rhoc, *half_rhoc;
    double rho, half_rho, u_ccsaq;
    unsigned active;
    int m, cons0, cons_n;
};

// pair range [p_lo, p_hi) of global group `seg`, relative to the start of this rank's shard
__device__ __forceinline__ void group_pairs(unsigned long long nchunks, unsigned nseg_total, unsigned long long chunk0,
                                            unsigned seg, unsigned long long *p_lo, unsigned long long *p_hi)
{
    *p_lo = ((unsigned long long) seg * nchunks / nseg_total - chunk0) * kChunkPairs;
    *p_hi = ((unsigned long long) (seg + 1) * nchunks / nseg_total - chunk0) * kChunkPairs;
}

// ---- block-level reduction with a fixed tree (end_outer_kernel) ----------------------------------
template <int NV, int BLOCK = kBlock>
__device__ __forceinline__ void block_reduce_to(double (&acc)[NV], double *smem /* [(BLOCK/32)*NV] */, double *out)
{
    constexpr int kWarpsB = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = addx(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) smem[warp * NV + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = smem[threadIdx.x];
#pragma unroll
        for (int w = 1; w < kWarpsB; ++w) s = addx(s, smem[w * NV + threadIdx.x]);
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// true in exactly one CTA: the one whose ticket completes `total`
__device__ __forceinline__ bool is_last_arrival(unsigned *ticket, unsigned total, int *s_flag)
{
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *s_flag = (atomicAdd(ticket, 1u) == total - 1u);
    __syncthreads();
    return *s_flag != 0;
}

// ---- per-variable closed forms ------------------------------------------------------------------
// MAXM rows of grad_c are kept in registers; FULL means m == MAXM with every constraint active, which
// strips the per-row predicates from the unrolled loops (the common case m in {1,2,4,8,16}).
// MMA: mma.c:96-129.
template <int MAXM, bool FULL, class MU>
__device__ __forceinline__ double mma_point(const MU &a, double x, double lb, double ub, double s, double g,
                                            const double (&Gr)[MAXM > 0 ? MAXM : 1], const double *Gcol,
                                            unsigned long long ld, double (&acc)[3 + (MAXM > 0 ? MAXM : 1)])
{
    if (s == 0) return x;                                    // fixed variable, mma.c:96-99
    const double ag_s = mulx(fabs(g), s);
    double u = g;
    double v = addx(ag_s, a.half_rho);
    if (FULL || a.m <= MAXM) {
#pragma unroll
        for (int i = 0; i < MAXM; ++i)
            if (FULL || (i < a.m && ((a.active >> i) & 1u))) {
                u = addx(u, mulx(Gr[i], a.y[i]));
                v = addx(v, mulx(addx(mulx(fabs(Gr[i]), s), a.half_rhoc[i]), a.y[i]));
            }
    } else {
        for (int i = 0; i < a.m; ++i)
            if ((a.active >> i) & 1u) {
                const double gi = Gcol[(unsigned long long) i * ld];
                u = addx(u, mulx(gi, a.y[i]));
                v = addx(v, mulx(addx(mulx(fabs(gi), s), a.half_rhoc[i]), a.y[i]));
            }
    }
    const double s2 = mulx(s, s);
    u = mulx(u, s2);
    const double r = divx(u, mulx(v, s));
    double dx = divx(divx(u, v), subx(-1.0, __dsqrt_rn(fabs(subx(1.0, mulx(r, r))))));   // mma.c:108
    double xc = addx(x, dx);
    if (xc > ub) xc = ub; else if (xc < lb) xc = lb;        // mma.c:110-111
    const double lim = mulx(0.9, s), hi = addx(x, lim), lo = subx(x, lim);
    if (xc > hi) xc = hi; else if (xc < lo) xc = lo;        // mma.c:112-113
    dx = subx(xc, x);
    const double dx2 = mulx(dx, dx);
    const double dinv = divx(1.0, subx(s2, dx2));
    acc[0] = addx(acc[0], mulx(addx(mulx(u, dx), mulx(v, dx2)), dinv));                  // mma.c:119
    const double c = mulx(s2, dx);
    acc[1] = addx(acc[1], mulx(addx(mulx(g, c), mulx(addx(ag_s, a.half_rho), dx2)), dinv));   // mma.c:123
    acc[2] = addx(acc[2], mulx(mulx(0.5, dx2), dinv));                                  // mma.c:125
#pragma unroll
    for (int k = 0; k < MAXM; ++k) {
        const int i = FULL ? k : a.cons0 + k;
        if (FULL || (k < a.cons_n && ((a.active >> i) & 1u))) {
            const double gi = (FULL || a.m <= MAXM) ? Gr[k] : Gcol[(unsigned long long) i * ld];
            acc[3 + k] = addx(acc[3 + k],
                              mulx(addx(mulx(gi, c), mulx(addx(mulx(fabs(gi), s), a.half_rhoc[i]), dx2)), dinv));   // mma.c:127
        }
    }
    return xc;
}

// CCSAQ: ccsa_quadratic.c:111-140
template <int MAXM, bool FULL, class MU>
__device__ __forceinline__ double ccsaq_point(const MU &a, double x, double lb, double ub, double s, double g,
                                              const double (&Gr)[MAXM > 0 ? MAXM : 1], const double *Gcol,
                                              unsigned long long ld, double (&acc)[3 + (MAXM > 0 ? MAXM : 1)])
{
    if (s == 0) return x;                                    // ccsa_quadratic.c:111-114
    double v = g;
    if (FULL || a.m <= MAXM) {
#pragma unroll
        for (int i = 0; i < MAXM; ++i)
            if (FULL || i < a.m) v = addx(v, mulx(Gr[i], a.y[i]));
    } else {
        for (int i = 0; i < a.m; ++i) v = addx(v, mulx(Gcol[(unsigned long long) i * ld], a.y[i]));
    }
    const double u = a.u_ccsaq;
    const double s2 = mulx(s, s);
    double dx = divx(mulx(-s2, v), u);                       // ccsa_quadratic.c:122
    if (fabs(dx) > s) dx = copysign(s, dx);                  // ccsa_quadratic.c:126
    double xc = addx(x, dx);
    if (xc > ub) xc = ub; else if (xc < lb) xc = lb;
    dx = subx(xc, x);
    const double dx2 = mulx(dx, dx);
    acc[0] = addx(acc[0], addx(mulx(v, dx), divx(mulx(mulx(0.5, u), dx2), s2)));         // ccsa_quadratic.c:134
    const double q = divx(mulx(0.5, dx2), s2);
    acc[1] = addx(acc[1], addx(mulx(g, dx), mulx(a.rho, q)));                            // :137
    acc[2] = addx(acc[2], q);                                                            // :138
#pragma unroll
    for (int k = 0; k < MAXM; ++k) {
        const int i = FULL ? k : a.cons0 + k;
        if (FULL || k < a.cons_n) {
            const double gi = (FULL || a.m <= MAXM) ? Gr[k] : Gcol[(unsigned long long) i * ld];
            acc[3 + k] = addx(acc[3 + k], addx(mulx(gi, dx), mulx(a.rhoc[i], q)));       // :139-140
        }
    }
    return xc;
}

// ---- the dual evaluation kernel ---------------------------------------------------------------------
// Persistent CTAs (grid sized to the machine).  A *group* is a contiguous run of 512-variable chunks;
// the 8 warps of a group slot sweep it together -- sweep step t reads one 4 KB-contiguous chunk per array,
// warp w taking lanes [32w, 32w+32) of it -- but every warp keeps its OWN m+3 accumulators over the group
// and folds them with a fixed xor-butterfly into a warp record.  The streaming loop has no barrier; one
// slot barrier per group hands the 8 warp records to warp 0 through shared memory.
// Fold tree (all in fixed order, all un-fused adds):
//   warp record -> group record (8 warp records in warp order, by warp 0 of the slot)
//               -> virtual-shard sum (P group records, by the warp that completes the shard)
//               -> rank sum / exchange buffer (8/world shard sums, by the warp that completes the rank).
// A record depends only on n (the cuts) -- never on the grid size, on which CTA swept the group or on the
// number of ranks -- so the m+3 sums are bit-identical for every launch geometry and every world size.
template <int NV>
__device__ __forceinline__ void warp_fold(double (&acc)[NV])
{
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[k] = addx(acc[k], __shfl_xor_sync(0xffffffffu, acc[k], off));
    }
}

__device__ __forceinline__ bool warp_is_last(unsigned *ticket, unsigned total, int lane)
{
    unsigned t = 0;
    __threadfence();
    if (lane == 0) t = atomicAdd(ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    __threadfence();
    return t == total - 1u;
}

// Virtual-shard sum of P group records, canonical order (the order the solve kernel's folder CTA produces with
// its 256 threads): chain t in [0, 256) adds records t, t+256, ... in index order; the 32 chains of "fold warp"
// w = t / 32 meet in an xor butterfly; the 8 fold-warp results are added in warp order.  Here one warp plays
// the 8 fold warps in turn.  Every lane returns all NV sums.
template <int NV>
__device__ __forceinline__ void fold_shard_records(const double *base, unsigned P, int nvp, int lane, double (&tot)[NV])
{
#pragma unroll 1
    for (int w = 0; w < kGroupWarps; ++w) {
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        for (unsigned r = 32 * w + lane; r < P; r += 32 * kGroupWarps)
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] = addx(acc[k], __ldcg(base + (unsigned long long) r * nvp + k));
        warp_fold<NV>(acc);
#pragma unroll
        for (int k = 0; k < NV; ++k) tot[k] = w == 0 ? acc[k] : addx(tot[k], acc[k]);
    }
}

// Sweep one group: this warp's lanes of every chunk of group `gl`, m+3 lane accumulators.
template <int VARIANT, int MAXM, bool FULL, int UNROLL, class MU>
__device__ __forceinline__ void sweep_group(const DualArgs &a, const MU &mu, bool store, unsigned gl, int sub,
                                            int lane, double (&acc)[3 + (MAXM > 0 ? MAXM : 1)])
{
    constexpr int MR = MAXM > 0 ? MAXM : 1;
    const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
    const double2 *lb2 = reinterpret_cast<const double2 *>(a.lb);
    const double2 *ub2 = reinterpret_cast<const double2 *>(a.ub);
    const double2 *s2v = reinterpret_cast<const double2 *>(a.sigma);
    const double2 *g2 = reinterpret_cast<const double2 *>(a.g);
    const bool in_regs = FULL || mu.m <= MAXM;
    unsigned long long p_lo, p_hi;
    group_pairs(a.nchunks, a.nseg_total, a.chunk0, a.seg0 + gl, &p_lo, &p_hi);

    for (unsigned long long p0 = p_lo + sub * 32 + lane; p0 < p_hi; p0 += (unsigned long long) kChunkPairs * UNROLL) {
        double2 vx[UNROLL], vlb[UNROLL], vub[UNROLL], vs[UNROLL], vg[UNROLL];
        double Ga[UNROLL][MR], Gb[UNROLL][MR];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned long long p = p0 + (unsigned long long) kChunkPairs * u;
            const bool live = u == 0 || p < p_hi;
            vs[u] = make_double2(0.0, 0.0);      // sigma = 0 lanes are skipped by both formulas
            vx[u] = vlb[u] = vub[u] = vg[u] = make_double2(0.0, 0.0);
            if (live) {
                vx[u] = ld_stream(x2 + p); vlb[u] = ld_stream(lb2 + p); vub[u] = ld_stream(ub2 + p);
                vs[u] = ld_stream(s2v + p); vg[u] = ld_stream(g2 + p);
            }
#pragma unroll
            for (int i = 0; i < MR; ++i) {
                Ga[u][i] = 0.0;
                Gb[u][i] = 0.0;
                if (MAXM > 0 && in_regs && (FULL || i < mu.m) && live) {
                    const double2 t = ld_stream(reinterpret_cast<const double2 *>(a.G + (unsigned long long) i * a.ld) + p);
                    Ga[u][i] = t.x;
                    Gb[u][i] = t.y;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const unsigned long long p = p0 + (unsigned long long) kChunkPairs * u;
            const bool live = u == 0 || p < p_hi;
            const double *col = a.G + 2 * p;
            double2 xc;
            if (VARIANT == 0) {
                xc.x = mma_point<MAXM, FULL>(mu, vx[u].x, vlb[u].x, vub[u].x, vs[u].x, vg[u].x, Ga[u], col, a.ld, acc);
                xc.y = mma_point<MAXM, FULL>(mu, vx[u].y, vlb[u].y, vub[u].y, vs[u].y, vg[u].y, Gb[u], col + 1, a.ld, acc);
            } else {
                xc.x = ccsaq_point<MAXM, FULL>(mu, vx[u].x, vlb[u].x, vub[u].x, vs[u].x, vg[u].x, Ga[u], col, a.ld, acc);
                xc.y = ccsaq_point<MAXM, FULL>(mu, vx[u].y, vlb[u].y, vub[u].y, vs[u].y, vg[u].y, Gb[u], col + 1, a.ld, acc);
            }
            if (store && live) st_stream(reinterpret_cast<double2 *>(a.xcur) + p, xc);
        }
    }
}

template <int VARIANT, int MAXM, bool FULL, bool STORE, int BLOCK, int UNROLL, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) dual_eval_kernel(const __grid_constant__ DualArgs a)
{
    constexpr int MR = MAXM > 0 ? MAXM : 1;
    constexpr int NV = 3 + MR;
    constexpr int SLOTS = BLOCK / (32 * kGroupWarps);        // groups a CTA sweeps at a time
    static_assert(BLOCK % (32 * kGroupWarps) == 0, "a CTA holds whole group slots");
    const int lane = threadIdx.x & 31;
    const int sub = (threadIdx.x >> 5) % kGroupWarps;        // which eighth of every chunk this warp owns
    const unsigned slot = blockIdx.x * SLOTS + (threadIdx.x >> 5) / kGroupWarps;
    const unsigned nslots = gridDim.x * SLOTS;
    const unsigned ngroups = a.segs_per_vshard * a.local_vshards;
    __shared__ double s_rec[SLOTS][2][kGroupWarps * NV];
    int parity = 0;
    for (unsigned gl = slot; gl < ngroups; gl += nslots) {
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        sweep_group<VARIANT, MAXM, FULL, UNROLL>(a, a, STORE, gl, sub, lane, acc);   // multipliers = the parameter block itself

        // warp record -> shared memory; group record = the 8 warp records added in warp order by warp 0 of
        // the slot.  One slot barrier per group; the record buffer is double-buffered across iterations so
        // the next group's writers can never overtake this group's reader.
        warp_fold<NV>(acc);
        double *srec = s_rec[(threadIdx.x >> 5) / kGroupWarps][parity];
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) srec[sub * NV + k] = acc[k];
        }
        if (SLOTS == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"((int) ((threadIdx.x >> 5) / kGroupWarps) + 1), "r"(32 * kGroupWarps) : "memory");
        parity ^= 1;
        if (sub != 0) continue;
        if (lane < NV) {
            double s = srec[lane];
#pragma unroll
            for (int w = 1; w < kGroupWarps; ++w) s = addx(s, srec[w * NV + lane]);
            a.grouprecs[(unsigned long long) gl * a.nvp + lane] = s;
        }
        __syncwarp();

        // virtual-shard sum, by the warp that completes the shard
        const unsigned vs_local = gl / a.segs_per_vshard;
        if (!warp_is_last(a.tickets + vs_local, a.segs_per_vshard, lane)) continue;
        fold_shard_records<NV>(a.grouprecs + (unsigned long long) vs_local * a.segs_per_vshard * a.nvp, a.segs_per_vshard,
                               a.nvp, lane, acc);
        if (lane == 0) {
            double *rec = a.vsums + (unsigned long long) vs_local * a.nvp;
#pragma unroll
            for (int k = 0; k < NV; ++k) rec[k] = acc[k];
        }

        // rank sum, by the warp that completes the last virtual shard
        if (!warp_is_last(a.tickets + a.local_vshards, a.local_vshards, lane)) continue;
        if (lane < NV) {
            if (a.publish_host) {
                double s = __ldcg(a.vsums + lane);
                for (unsigned v = 1; v < a.local_vshards; ++v) s = addx(s, __ldcg(a.vsums + (unsigned long long) v * a.nvp + lane));
                a.out_host[lane] = s;
                __threadfence_system();
            } else if (a.box[0] == nullptr) {
                const unsigned v0 = a.seg0 / a.segs_per_vshard;
                for (unsigned v = 0; v < a.local_vshards; ++v)
                    a.out_dev[(unsigned long long) (v0 + v) * a.nvp + lane] = __ldcg(a.vsums + (unsigned long long) v * a.nvp + lane);
            }
        }
        if (!a.publish_host && a.box[0] != nullptr) {
            // ---- all-gather fused into the kernel: tagged NVLink peer stores (comm.hpp layout) ----
            int timed_out = 0;
            const double total = box_exchange(a.box, a.rank, a.world, a.seq, a.vsums, a.nvp, a.local_vshards,
                                              a.seg0 / a.segs_per_vshard, NV, lane, &timed_out);
            if (lane < NV) {
                a.out_host[lane] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : total;
                __threadfence_system();
            }
        }
        __syncwarp();
        if (lane == 0) {
            for (unsigned v = 0; v <= a.local_vshards; ++v) a.tickets[v] = 0;    // ready for the next launch
            if (a.publish_host || a.box[0] != nullptr) {
                *a.flag_host = a.seq;
                __threadfence_system();
            }
        }
    }
}

// ---- the dual evaluation kernel, TMA-staged form -------------------------------------------------------------
// Same groups, same per-warp accumulators, same fold tree (=> same bits) as dual_eval_kernel; what changes is
// how the operands arrive.  A producer warp issues 1-D TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx): one 4 KB chunk of each of the 5+m arrays per stage, STAGES stages deep, running ahead
// across group boundaries so the pipeline never drains.  The 8 consumer warps wait on the stage's "full"
// barrier, read their double2 lanes from shared memory, and release the stage on its "empty" barrier.
// No load ever occupies a consumer register before it is needed, so many more bytes are in flight per SM
// than the register form can hold at the same occupancy.
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

constexpr int kTmaBlock = 32 * (kGroupWarps + 1);        // 8 consumer warps + 1 producer warp
constexpr unsigned kChunkBytes = kChunkPairs * 16;       // 4 KB per array per stage

template <int VARIANT, int MAXM, bool STORE, int STAGES, int MINB>
__global__ void __launch_bounds__(kTmaBlock, MINB) dual_eval_tma_kernel(const __grid_constant__ DualArgs a)
{
    constexpr int MR = MAXM > 0 ? MAXM : 1;
    constexpr int NV = 3 + MR;
    constexpr int NARR = 5 + MAXM;                            // x lb ub sigma g + MAXM gradient rows
    extern __shared__ __align__(128) unsigned char s_raw[];
    double2 *s_tile = reinterpret_cast<double2 *>(s_raw);     // [STAGES][NARR][kChunkPairs]
    __shared__ unsigned long long s_full[STAGES], s_empty[STAGES];
    __shared__ double s_rec[2][kGroupWarps * NV];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned ngroups = a.segs_per_vshard * a.local_vshards;

    if (threadIdx.x == 0) {
        for (int st = 0; st < STAGES; ++st) {
            mbar_init(&s_full[st], 1);                        // one arrive.expect_tx by the producer
            mbar_init(&s_empty[st], kGroupWarps);             // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kGroupWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const double *src[NARR];
            src[0] = a.x; src[1] = a.lb; src[2] = a.ub; src[3] = a.sigma; src[4] = a.g;
#pragma unroll
            for (int i = 0; i < MAXM; ++i) src[5 + i] = a.G + (unsigned long long) i * a.ld;
            int st = 0;
            unsigned phase = 0;
            for (unsigned gl = blockIdx.x; gl < ngroups; gl += gridDim.x) {
                unsigned long long p_lo, p_hi;
                group_pairs(a.nchunks, a.nseg_total, a.chunk0, a.seg0 + gl, &p_lo, &p_hi);
                for (unsigned long long p = p_lo; p < p_hi; p += kChunkPairs) {
                    mbar_wait(&s_empty[st], phase ^ 1u);      // passes at once on a fresh barrier
                    mbar_expect_tx(&s_full[st], NARR * kChunkBytes);
#pragma unroll
                    for (int k = 0; k < NARR; ++k)
                        tma_bulk_load(s_tile + ((size_t) st * NARR + k) * kChunkPairs, src[k] + 2 * p, kChunkBytes, &s_full[st]);
                    if (++st == STAGES) { st = 0; phase ^= 1u; }
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int sub = warp;
    int st = 0;
    unsigned phase = 0;
    int parity = 0;
    for (unsigned gl = blockIdx.x; gl < ngroups; gl += gridDim.x) {
        unsigned long long p_lo, p_hi;
        group_pairs(a.nchunks, a.nseg_total, a.chunk0, a.seg0 + gl, &p_lo, &p_hi);
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        for (unsigned long long p = p_lo; p < p_hi; p += kChunkPairs) {
            mbar_wait(&s_full[st], phase);
            const double2 *t = s_tile + (size_t) st * NARR * kChunkPairs + sub * 32 + lane;
            const double2 vx = t[0], vlb = t[kChunkPairs], vub = t[2 * kChunkPairs], vs = t[3 * kChunkPairs],
                          vg = t[4 * kChunkPairs];
            double Ga[MR], Gb[MR];
#pragma unroll
            for (int i = 0; i < MR; ++i) {
                Ga[i] = 0.0; Gb[i] = 0.0;
                if (MAXM > 0) { const double2 g2 = t[(5 + i) * kChunkPairs]; Ga[i] = g2.x; Gb[i] = g2.y; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);         // operands are in registers: release the stage
            if (++st == STAGES) { st = 0; phase ^= 1u; }
            double2 xc;
            if (VARIANT == 0) {
                xc.x = mma_point<MAXM, true>(a, vx.x, vlb.x, vub.x, vs.x, vg.x, Ga, nullptr, a.ld, acc);
                xc.y = mma_point<MAXM, true>(a, vx.y, vlb.y, vub.y, vs.y, vg.y, Gb, nullptr, a.ld, acc);
            } else {
                xc.x = ccsaq_point<MAXM, true>(a, vx.x, vlb.x, vub.x, vs.x, vg.x, Ga, nullptr, a.ld, acc);
                xc.y = ccsaq_point<MAXM, true>(a, vx.y, vlb.y, vub.y, vs.y, vg.y, Gb, nullptr, a.ld, acc);
            }
            if (STORE) st_stream(reinterpret_cast<double2 *>(a.xcur) + p + sub * 32 + lane, xc);
        }

        warp_fold<NV>(acc);
        double *srec = s_rec[parity];
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) srec[sub * NV + k] = acc[k];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * kGroupWarps) : "memory");      // the 8 consumer warps only
        parity ^= 1;
        if (sub != 0) continue;
        if (lane < NV) {
            double s = srec[lane];
#pragma unroll
            for (int w = 1; w < kGroupWarps; ++w) s = addx(s, srec[w * NV + lane]);
            a.grouprecs[(unsigned long long) gl * a.nvp + lane] = s;
        }
        __syncwarp();
        const unsigned vs_local = gl / a.segs_per_vshard;
        if (!warp_is_last(a.tickets + vs_local, a.segs_per_vshard, lane)) continue;
        fold_shard_records<NV>(a.grouprecs + (unsigned long long) vs_local * a.segs_per_vshard * a.nvp, a.segs_per_vshard,
                               a.nvp, lane, acc);
        if (lane == 0) {
            double *rec = a.vsums + (unsigned long long) vs_local * a.nvp;
#pragma unroll
            for (int k = 0; k < NV; ++k) rec[k] = acc[k];
        }
        if (!warp_is_last(a.tickets + a.local_vshards, a.local_vshards, lane)) continue;
        if (lane < NV) {
            if (a.publish_host) {
                double s = __ldcg(a.vsums + lane);
                for (unsigned v = 1; v < a.local_vshards; ++v) s = addx(s, __ldcg(a.vsums + (unsigned long long) v * a.nvp + lane));
                a.out_host[lane] = s;
                __threadfence_system();
            } else if (a.box[0] == nullptr) {
                const unsigned v0 = a.seg0 / a.segs_per_vshard;
                for (unsigned v = 0; v < a.local_vshards; ++v)
                    a.out_dev[(unsigned long long) (v0 + v) * a.nvp + lane] = __ldcg(a.vsums + (unsigned long long) v * a.nvp + lane);
            }
        }
        if (!a.publish_host && a.box[0] != nullptr) {
            int timed_out = 0;
            const double total = box_exchange(a.box, a.rank, a.world, a.seq, a.vsums, a.nvp, a.local_vshards,
                                              a.seg0 / a.segs_per_vshard, NV, lane, &timed_out);
            if (lane < NV) {
                a.out_host[lane] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : total;
                __threadfence_system();
            }
        }
        __syncwarp();
        if (lane == 0) {
            for (unsigned v = 0; v <= a.local_vshards; ++v) a.tickets[v] = 0;
            if (a.publish_host || a.box[0] != nullptr) {
                *a.flag_host = a.seq;
                __threadfence_system();
            }
        }
    }
}

// ---- the persistent dual-SOLVE kernel: one launch per dual solve ------------------------------------------
//   The m-dimensional dual optimiser (DualMachine, dual_mma.hpp -- the same code the
// host runs) moves into the kernel: all CTAs stay resident (cooperative launch) and walk *generations*.
// Generation g = one dual evaluation at the trial multipliers y_g.
//
//   sweeper CTAs (all but the last): claim groups from a monotonic counter (claim c -> generation
//     c / ngroups + 1, group c % ngroups), sweep them exactly like dual_eval_kernel (same warp records,
//     same group records => same bits) and drop each group record into a *tagged* 16-byte slot
//     {value, tag(launch, generation)} with one 128-bit store.  No fence, no ticket, no atomic on the
//     record path: the timeline of the previous (ticket) design showed warp 0 of every CTA spending
//     ~4.6 us per group in fence + atomic + fence under full memory load while its seven sibling warps
//     waited at the next slot barrier (profiles/r01_trace_summary_ticket_design.txt).
//   the folder CTA (the last one): warp v polls the P slots of local virtual shard v in index order
//     (lane l takes records l, l+32, ...; then the xor butterfly -- the fold tree of dual_eval_kernel),
//     one CTA barrier hands the <= 8 shard sums to warp 0, which exchanges them over the NVLink mailbox
//     when there are several ranks, feeds F and grad F to the DualMachine held in ITS shared memory,
//     and publishes y_{g+1} as tagged slots the sweepers poll (again no fence: a slot is valid iff its
//     tag is the awaited generation).
// Versus one launch per evaluation this removes launch latency, the PCIe result hop and the host turn-
// around from every evaluation; the host sees one launch and one result per dual solve.
//
// Timeline instrumentation (tools/trace_solve.py builds a separate library with -DNB200_TRACE; the product build
// contains none of it).  Per generation g, 16 counters at trace[16 g]: 0 published | 1 ~min / 2 max "CTA saw it" |
// 3 ~min / 4 max "group record stored" | 5 all shard sums in | 6 totals ready | 7 optimiser done | 8 sum / 9 count
// of per-group sweep times.  Row 0 holds the CTA start times.  All in %globaltimer nanoseconds.
#ifdef NB200_TRACE
constexpr int kTraceGens = 512;
__device__ __forceinline__ unsigned long long nb_gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define NB_TR(...) __VA_ARGS__
#else
#define NB_TR(...)
#endif

constexpr int kPubSlots = kMaxParamM + 2;     // y_i | u_ccsaq | flags
struct SolveState {                       // device global; the head is zeroed by the host before every launch
    unsigned long long claim;             // monotonic group-claim counter
    int done;                             // 1: leave
    int pad;
    double pub[2 * kPubSlots];            // tagged slots {value, tag}: trial multipliers of the generation in flight,
                                          // u = rho + sum rhoc_i y_i, and (as an integer) bit 0 = also store x*(y)
};

struct SolveArgs {
    DualArgs d;                           // arrays, geometry, workspace, exchange boxes (d.y: the warm start)
    SolveState *st;
    double *grouptags;                    // [nvp][local groups] tagged slots {value, tag}
    unsigned long long tag0;              // launch id << 40; generation g carries tag0 | g
    double fval;                          // objective value at x
    double cval[kMaxParamM];              // constraint values with switched-off ones zeroed (mma.c:78)
    double lo[kMaxParamM], hi[kMaxParamM];    // box of the multipliers
    DualStop stop;
    volatile double *res_host;            // mapped pinned: raw sums [24] | y [32] | nevals | ret | generations
    NB_TR(unsigned long long *trace;)
};

struct SharedMultipliers {                // what the point functions read in the solve kernel
    const double *y, *rhoc, *half_rhoc;   // y in shared memory; penalties from the parameter block
    double rho, half_rho, u_ccsaq;
    unsigned active;
    int m, cons0, cons_n;
};

// The folder CTA's loop (kept out of line so that its registers do not weigh on the sweep loop).
template <int NV>
__device__ __noinline__ void solve_folder(const SolveArgs &sa)
{
    const DualArgs &a = sa.d;
    SolveState *st = sa.st;
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x >> 5;
    const unsigned ngroups = a.segs_per_vshard * a.local_vshards;
    __shared__ int s_exit;
    __shared__ DualMachine s_mach;
    __shared__ double s_grad[kMaxParamM];
    __shared__ double s_vs[kVirtualShards * NV];      // the shard sums of the generation in flight
    __shared__ double s_w[kVirtualShards * kGroupWarps * NV];     // per shard: the 8 fold-warp results
    __shared__ double s_gt[kMaxParamM], s_wt[kMaxParamM];
    __shared__ int s_has[kMaxParamM];
    // ================================ the folder CTA ================================
    int final_pass = 0;
    if (threadIdx.x == 0) {
        s_exit = 0;
        const int rc = s_mach.start(a.m, a.y, sa.lo, sa.hi, sa.stop);      // d.y carries the warm start
        if (rc != kRetSuccess) {          // start point outside the box: report, publish nothing
            sa.res_host[24 + kMaxParamM + 1] = (double) rc;
            __threadfence_system();
            *a.flag_host = a.seq;
            __threadfence_system();
            *reinterpret_cast<volatile int *>(&st->done) = 1;
            s_exit = 1;
        } else {
            double u = a.rho;
            for (int i = 0; i < a.m; ++i) u = addx(u, mulx(a.rhoc[i], s_mach.y[i]));
            NB_TR(sa.trace[16] = nb_gtime();)
            for (int i = 0; i < a.m; ++i) slot_put(st->pub + 2 * i, s_mach.y[i], sa.tag0 | 1ull);
            slot_put(st->pub + 2 * kMaxParamM, u, sa.tag0 | 1ull);
            slot_put(st->pub + 2 * (kMaxParamM + 1), __longlong_as_double(0ll), sa.tag0 | 1ull);
        }
    }
    for (int i = threadIdx.x; i < kVirtualShards * kGroupWarps * NV; i += 32 * kGroupWarps) s_w[i] = 0.0;
    const unsigned fw_all = (a.segs_per_vshard + 31u) / 32u;
    const unsigned fw_per = fw_all < (unsigned) kGroupWarps ? fw_all : (unsigned) kGroupWarps;      // non-empty fold warps per shard
    const unsigned nitems = a.local_vshards * fw_per;
    __syncthreads();
    if (s_exit) return;
    const long long t_start = clock64();
    for (unsigned long long gen = 1;; ++gen) {
        const unsigned long long tag = sa.tag0 | gen;
        // ---- shard sums, canonical order (see fold_shard_records).  Work item (v, w) = fold warp w of local
        // shard v: chains t = 32 w + lane over records t, t + 256, ...  Items are dealt round-robin to the 8
        // physical warps in (v, w) order -- shards complete roughly in index order, and when P <= 32 (small n, or
        // one shard per rank with 8 GPUs) all shards are polled side by side.  Empty fold warps contribute the
        // +0.0 parked in s_w at start-up.
        for (unsigned item = sub; item < nitems; item += kGroupWarps) {
            const unsigned v = item / fw_per, w = item % fw_per;
            double acc[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] = 0.0;
            // slot (group gl, sum k) lives at [k][gl]: the 32 lanes of a poll read 512 contiguous bytes
            const double *base = sa.grouptags + 2ull * (unsigned long long) v * a.segs_per_vshard;
            // warp-uniform control flow: a warp polls until all of its lanes have their record (measured: a warp
            // whose lanes left the poll loop at different times took ~9 us per shard instead of < 1 us)
            for (unsigned r0 = 32u * w; r0 < a.segs_per_vshard; r0 += 32 * kGroupWarps) {
                const unsigned r = r0 + lane;
                const bool has = r < a.segs_per_vshard;
                const double *rec = base + 2ull * (has ? r : 0u);
                double val[NV];
                for (;;) {                // the NV slots of a record are fetched together: one round trip per poll
                    bool all = true;
#pragma unroll
                    for (int k = 0; k < NV; ++k) all = slot_peek(rec + 2ull * k * ngroups, tag, &val[k]) && all;
                    if (__all_sync(0xffffffffu, all || !has)) break;
                    __nanosleep(20);
                }
                if (has) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) acc[k] = addx(acc[k], val[k]);
                }
            }
            warp_fold<NV>(acc);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < NV; ++k) s_w[(v * kGroupWarps + w) * NV + k] = acc[k];
            }
        }
        NB_TR(if (threadIdx.x == 0 && gen < kTraceGens) sa.trace[16 * gen + 14] = nb_gtime();)
        __syncthreads();
        if (sub == 0 && lane < NV) {
            for (unsigned v = 0; v < a.local_vshards; ++v) {
                double t = s_w[(v * kGroupWarps) * NV + lane];
#pragma unroll
                for (int w = 1; w < kGroupWarps; ++w) t = addx(t, s_w[(v * kGroupWarps + w) * NV + lane]);
                s_vs[v * NV + lane] = t;
            }
        }
        __syncwarp();
        // ---- warp 0: totals (exchange if sharded), the dual optimiser's turn, publication ----
        if (sub == 0) {
            NB_TR(if (lane == 0 && gen < kTraceGens) sa.trace[16 * gen + 5] = nb_gtime();)
            double total = 0.0;               // lane k < NV holds sum k
            int timed_out = 0;
            if (a.box[0] == nullptr) {
                if (lane < NV) {
                    total = s_vs[lane];
                    for (unsigned v = 1; v < a.local_vshards; ++v) total = addx(total, s_vs[v * NV + lane]);
                }
            } else {
                total = box_exchange<true>(a.box, a.rank, a.world, a.seq + gen, s_vs, NV, a.local_vshards,
                                           a.seg0 / a.segs_per_vshard, NV, lane, &timed_out);     // one tag per generation
            }
            NB_TR(if (lane == 0 && gen < kTraceGens) sa.trace[16 * gen + 6] = nb_gtime();)
            // F and grad F from the sums, constants added in the reference's order (mma.c:75-78, :135)
            int finished = 0, next_final = 0;
            if (!final_pass) {
                if (lane >= 3 && lane < 3 + a.m) s_grad[lane - 3] = -addx(sa.cval[lane - 3], total);   // -g_i(y)
                const double sum0 = __shfl_sync(0xffffffffu, total, 0);
                __syncwarp();
                if (lane == 0) {
                    const double *yt = s_mach.trial();
                    double val = sa.fval;
                    for (int i = 0; i < a.m; ++i) val = addx(val, mulx(yt[i], sa.cval[i]));
                    val = addx(val, sum0);
                    const double elapsed = (double) (clock64() - t_start) * 5e-10;     // ~2 GHz; only feeds maxtime
                    finished = timed_out ? 1 : (s_mach.feed_pre(-val, s_grad, elapsed) ? 1 : 0);
                    if (timed_out) s_mach.ret = kRetFailure;
                }
                finished = __shfl_sync(0xffffffffu, finished, 0);
                if (!finished) {          // the m terms of the next trial point side by side (divisions, square root)
                    __syncwarp();
                    if (lane < a.m) s_has[lane] = s_mach.step_term(lane, &s_gt[lane], &s_wt[lane]) ? 1 : 0;
                    __syncwarp();
                    if (lane == 0) s_mach.step_sum(s_gt, s_wt, s_has);
                    __syncwarp();
                }
                if (finished && !timed_out) next_final = 1;          // one more pass at the solution, storing x*(y)
            }
            __syncwarp();
            NB_TR(if (lane == 0 && gen < kTraceGens) sa.trace[16 * gen + 7] = nb_gtime();)
            if (final_pass || timed_out) {
                // publish the result of the solve: raw sums of the final pass, multipliers, counts
                if (lane < NV) sa.res_host[lane] = total;
                if (lane < a.m) sa.res_host[24 + lane] = s_mach.y[lane];
                if (lane == 0) {
                    sa.res_host[24 + kMaxParamM] = (double) s_mach.nevals;
                    sa.res_host[24 + kMaxParamM + 1] = (double) s_mach.ret;
                    sa.res_host[24 + kMaxParamM + 2] = (double) gen;
                }
                __threadfence_system();
                __syncwarp();
                if (lane == 0) {
                    *a.flag_host = a.seq;
                    __threadfence_system();
                    *reinterpret_cast<volatile int *>(&st->done) = 1;
                    s_exit = 1;
                }
            } else {
                // publish generation gen + 1
                const double *trial = next_final ? s_mach.y : s_mach.ycur;
                const unsigned long long ntag = sa.tag0 | (gen + 1);
                NB_TR(if (lane == 0 && gen + 1 < kTraceGens) sa.trace[16 * (gen + 1)] = nb_gtime();)
                if (lane < a.m) slot_put(st->pub + 2 * lane, trial[lane], ntag);
                if (lane == 0) {
                    double u = a.rho;
                    for (int i = 0; i < a.m; ++i) u = addx(u, mulx(a.rhoc[i], trial[i]));
                    slot_put(st->pub + 2 * kMaxParamM, u, ntag);
                    slot_put(st->pub + 2 * (kMaxParamM + 1), __longlong_as_double((long long) next_final), ntag);
                }
                final_pass = next_final;
            }
            final_pass = __shfl_sync(0xffffffffu, final_pass, 0);
        }
        __syncthreads();
        if (s_exit) return;
    }
}

template <int VARIANT, int MAXM, bool FULL, int BLOCK, int UNROLL, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) dual_solve_kernel(const __grid_constant__ SolveArgs sa)
{
    constexpr int MR = MAXM > 0 ? MAXM : 1;
    constexpr int NV = 3 + MR;
    static_assert(BLOCK == 32 * kGroupWarps, "one group slot per CTA");
    static_assert(kGroupWarps == kVirtualShards, "the folder CTA gives one warp to each virtual shard");
    if (blockIdx.x == gridDim.x - 1) {
        solve_folder<NV>(sa);
        return;
    }
    const DualArgs &a = sa.d;
    SolveState *st = sa.st;
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x >> 5;
    const unsigned ngroups = a.segs_per_vshard * a.local_vshards;
    __shared__ double s_rec[2][kGroupWarps * NV];
    __shared__ double s_y[kMaxParamM];
    __shared__ double s_u;
    __shared__ int s_store, s_exit;
    __shared__ unsigned long long s_claim[2];

    // ================================ sweeper CTAs ================================
    unsigned long long next_c = 0;
    if (threadIdx.x == 0) { s_exit = 0; s_claim[0] = atomicAdd(&st->claim, 1ull); }
    NB_TR(if (threadIdx.x == 0) { const unsigned long long t = nb_gtime(); atomicMax(&sa.trace[1], ~t); atomicMax(&sa.trace[2], t); })
    __syncthreads();

    int parity = 0;
    unsigned long long my_gen = 0;        // generation whose multipliers are in s_y
    for (int it = 0;; ++it) {
        const unsigned long long c = s_claim[it & 1];
        const unsigned long long want = c / ngroups + 1;
        const unsigned gl = (unsigned) (c % ngroups);
        // wait until generation `want` is published (or the solve has finished); refresh the multipliers
        if (want != my_gen) {
            if (sub == 0) {                   // warp 0 polls, warp-uniformly: lane i < m: y_i, lane m: u, the others: flags
                const int slot = lane < a.m ? lane : (lane == a.m ? kMaxParamM : kMaxParamM + 1);
                const unsigned long long tag = sa.tag0 | want;
                double v;
                int ex = 0;
                unsigned spins = 0;
                for (;;) {
                    if (__all_sync(0xffffffffu, slot_get(st->pub + 2 * slot, tag, &v))) break;
                    if ((++spins & 7u) == 0u && __any_sync(0xffffffffu, ld_gpu_s32(&st->done))) {
                        ex = !__all_sync(0xffffffffu, slot_get(st->pub + 2 * slot, tag, &v));     // published before done was raised?
                        break;
                    }
                    __nanosleep(20);
                }
                if (ex) s_exit = 1;
                else if (lane < a.m) s_y[lane] = v;
                else if (lane == a.m) s_u = v;
                else if (lane == a.m + 1) s_store = (int) (__double_as_longlong(v) & 1ll);
            }
            __syncthreads();
            if (s_exit) return;
            NB_TR(if (threadIdx.x == 0 && want < kTraceGens) { const unsigned long long t = nb_gtime();
                      atomicMax(&sa.trace[16 * want + 1], ~t); atomicMax(&sa.trace[16 * want + 2], t); })
            my_gen = want;
        }
        // claim the next group now; the result is parked in a register until the sweep is over
        if (threadIdx.x == 0) next_c = atomicAdd(&st->claim, 1ull);

        SharedMultipliers mu;
        mu.y = s_y; mu.rhoc = a.rhoc; mu.half_rhoc = a.half_rhoc;
        mu.rho = a.rho; mu.half_rho = a.half_rho; mu.u_ccsaq = s_u;
        mu.active = a.active; mu.m = a.m; mu.cons0 = 0; mu.cons_n = a.m;
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        NB_TR(const unsigned long long tr_s0 = nb_gtime();)
        sweep_group<VARIANT, MAXM, FULL, UNROLL>(a, mu, s_store != 0, gl, sub, lane, acc);

        warp_fold<NV>(acc);
        double *srec = s_rec[parity];
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) srec[sub * NV + k] = acc[k];
        }
        if (threadIdx.x == 0) s_claim[(it + 1) & 1] = next_c;
        __syncthreads();
        parity ^= 1;
        if (sub == 0 && lane < NV) {
            double s = srec[lane];
#pragma unroll
            for (int w = 1; w < kGroupWarps; ++w) s = addx(s, srec[w * NV + lane]);
            slot_put(sa.grouptags + 2ull * ((unsigned long long) lane * ngroups + gl), s, sa.tag0 | my_gen);
        }
        NB_TR(if (sub == 0 && lane == 0 && my_gen < kTraceGens) { const unsigned long long t = nb_gtime(); unsigned long long *r = sa.trace + 16 * my_gen;
                  atomicMax(r + 3, ~t); atomicMax(r + 4, t); atomicAdd(r + 8, t - tr_s0); atomicAdd(r + 9, 1ull); })
    }
}

// Gradient of the augmented-Lagrangian objective (auglag.c:47-48, :59-60): g_j += coef_k * row_k[j] for the
// K penalty rows in index order, separate multiply and add like the reference's loop (=> bit-identical gradient).
// Rows with coef == 0 flagged by `skip` are left out entirely (an inactive inequality adds nothing, auglag.c:57).
constexpr int kPenaltyRowsPerLaunch = 16;
struct PenaltyCoefs {
    double c[kPenaltyRowsPerLaunch];
    int row[kPenaltyRowsPerLaunch];       // index of the row in the scratch block
    int count;
};
__global__ void __launch_bounds__(kBlock) penalty_axpy_kernel(double *__restrict__ g, const double *__restrict__ rows,
                                                               unsigned long long ld, unsigned long long n_local,
                                                               const __grid_constant__ PenaltyCoefs pc)
{
    const unsigned long long stride = (unsigned long long) gridDim.x * blockDim.x;
    for (unsigned long long j = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x; j < n_local; j += stride) {
        double v = g[j];
        for (int k = 0; k < pc.count; ++k) v = addx(v, mulx(pc.c[k], rows[(unsigned long long) pc.row[k] * ld + j]));
        g[j] = v;
    }
}

// After the all-gather (several ranks): fold the 8 shard sums in index order and publish.
__global__ void publish_kernel(const double *all_vsums /* [8][nvp] */, int nv, int nvp, volatile double *out_host,
                               volatile unsigned long long *flag_host, unsigned long long seq)
{
    if (threadIdx.x < nv) {
        double s = all_vsums[threadIdx.x];
        for (int v = 1; v < kVirtualShards; ++v) s = addx(s, all_vsums[v * nvp + threadIdx.x]);
        out_host[threadIdx.x] = s;
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *flag_host = seq;
        __threadfence_system();
    }
}

__global__ void fill_kernel(double *dst, double value, unsigned long long n_local)
{
    for (unsigned long long j = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; j < n_local;
         j += (unsigned long long) gridDim.x * blockDim.x)
        dst[j] = value;
}

// ---- sigma initialisation, mma.c:202-210 ---------------------------------------------------------
__device__ __forceinline__ bool dev_isinf(double v) { return fabs(v) >= HUGE_VAL * 0.99 || isinf(v); }

__global__ void sigma_init_kernel(double *sigma, const double *lb, const double *ub, const double *sigma_init,
                                  double sigma_min, unsigned long long n_local)
{
    for (unsigned long long j = blockIdx.x * (unsigned long long) blockDim.x + threadIdx.x; j < n_local;
         j += (unsigned long long) gridDim.x * blockDim.x) {
        double s;
        if (sigma_init && sigma_init[j] > 0) s = sigma_init[j];
        else if (dev_isinf(ub[j]) || dev_isinf(lb[j])) s = 1.0;
        else s = mulx(0.5, subx(ub[j], lb[j]));
        sigma[j] = s > sigma_min ? s : sigma_min;
    }
}

// ---- fused end-of-outer-iteration pass -------------------------------------------------------------
struct EndOuterArgs {
    const double *xcur;
    double *xprev, *xprevprev, *sigma;
    const double *lb, *ub;
    const double *w;          // x weights or null (stop.c:37-79)
    const double *xtol_abs;   // or null
    unsigned long long n_local, nchunks, chunk0;
    unsigned nseg_total, seg0, segs_per_vshard, local_vshards;
    double *partials, *vsums;
    unsigned *tickets;
    double *out_dev;
    volatile double *out_host;
    volatile unsigned long long *flag_host;
    unsigned long long seq;
    int publish_host, nvp;
    int update_sigma;         // k > 1
    double kappa;             // 0.01 (mma.c:439) or 1e-8 (ccsa_quadratic.c:587)
    double sigma_min;
};

__global__ void __launch_bounds__(kBlock) end_outer_kernel(const __grid_constant__ EndOuterArgs a)
{
    constexpr int NV = 3;     // sum w|dx|, sum w|x|, count of |dx| >= xtol_abs
    __shared__ double s_red[kWarps * NV];
    __shared__ int s_flag;
    const unsigned seg = a.seg0 + blockIdx.x;
    unsigned long long p_lo, p_hi;
    group_pairs(a.nchunks, a.nseg_total, a.chunk0, seg, &p_lo, &p_hi);
    double acc[NV] = {0.0, 0.0, 0.0};
    for (unsigned long long p = p_lo + threadIdx.x; p < p_hi; p += kBlock) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const unsigned long long j = 2 * p + h;
            if (j >= a.n_local) break;
            const double xc = a.xcur[j], xp = a.xprev[j];
            const double d = fabs(subx(xc, xp));
            if (a.w) {
                acc[0] = addx(acc[0], mulx(a.w[j], d));
                acc[1] = addx(acc[1], mulx(a.w[j], fabs(xc)));
            } else {
                acc[0] = addx(acc[0], d);
                acc[1] = addx(acc[1], fabs(xc));
            }
            if (a.xtol_abs && d >= a.xtol_abs[j]) acc[2] = addx(acc[2], 1.0);
            if (a.update_sigma) {
                const double xpp = a.xprevprev[j];
                const double osc = mulx(subx(xc, xp), subx(xp, xpp));
                double s = mulx(a.sigma[j], osc < 0 ? 0.7 : (osc > 0 ? 1.2 : 1.0));
                const double lo = a.lb[j], hi = a.ub[j];
                if (!dev_isinf(hi) && !dev_isinf(lo)) {
                    const double range = subx(hi, lo);
                    const double top = mulx(10.0, range), bot = mulx(a.kappa, range);
                    s = s < top ? s : top;
                    s = s > bot ? s : bot;
                }
                a.sigma[j] = s > a.sigma_min ? s : a.sigma_min;
            }
            a.xprevprev[j] = xp;
            a.xprev[j] = xc;
        }
    }
    block_reduce_to<NV>(acc, s_red, a.partials + (unsigned long long) blockIdx.x * a.nvp);
    const unsigned vs_local = blockIdx.x / a.segs_per_vshard;
    if (!is_last_arrival(a.tickets + vs_local, a.segs_per_vshard, &s_flag)) return;
    acc[0] = acc[1] = acc[2] = 0.0;
    {
        const double *base = a.partials + (unsigned long long) vs_local * a.segs_per_vshard * a.nvp;
        for (unsigned sgi = threadIdx.x; sgi < a.segs_per_vshard; sgi += kBlock)
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] = addx(acc[k], __ldcg(base + (unsigned long long) sgi * a.nvp + k));
    }
    block_reduce_to<NV>(acc, s_red, a.vsums + (unsigned long long) vs_local * a.nvp);
    if (!is_last_arrival(a.tickets + a.local_vshards, a.local_vshards, &s_flag)) return;
    if (threadIdx.x < NV) {
        if (a.publish_host) {
            double s = __ldcg(a.vsums + threadIdx.x);
            for (unsigned v = 1; v < a.local_vshards; ++v) s = addx(s, __ldcg(a.vsums + (unsigned long long) v * a.nvp + threadIdx.x));
            a.out_host[threadIdx.x] = s;
            __threadfence_system();
        } else {
            const unsigned v0 = a.seg0 / a.segs_per_vshard;
            for (unsigned v = 0; v < a.local_vshards; ++v)
                a.out_dev[(unsigned long long) (v0 + v) * a.nvp + threadIdx.x] = __ldcg(a.vsums + (unsigned long long) v * a.nvp + threadIdx.x);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (unsigned v = 0; v <= a.local_vshards; ++v) a.tickets[v] = 0;
        if (a.publish_host) {
            *a.flag_host = a.seq;
            __threadfence_system();
        }
    }
}

}  // namespace nb200
