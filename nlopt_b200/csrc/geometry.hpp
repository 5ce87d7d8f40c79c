// geometry.hpp -- how the n variables are cut into chunks, groups, virtual shards and rank shards.
//
// Everything here is a function of n alone (plus two tuning constants), never of the number of
// GPUs or of the launch geometry:
//   chunk  = 256 double2 pairs = 512 variables (one 4 KB sweep step of an 8-warp group slot);
//   group s of S = 8*P covers chunks [s*NC/S, (s+1)*NC/S), NC = ceil(ceil(n/2)/256);
//   virtual shard v = groups [v*P, (v+1)*P);
//   rank r of `world` (1, 2, 4 or 8) owns virtual shards [r*8/world, (r+1)*8/world): one
//   contiguous block of variables starting at a multiple of 512.
// Fixed cuts + fixed fold order = sums that are bit-identical for every world size and every grid
// (SURVEY.md 8(e)).
#pragma once

#include <cstdint>

#ifndef NB200_MIN_GROUP_CHUNKS
#define NB200_MIN_GROUP_CHUNKS 4
#endif

namespace nb200 {

constexpr unsigned kV = 8;              // virtual shards (== kVirtualShards in ccsa_kernels.cuh)
constexpr unsigned kChunkVars = 512;    // variables per chunk (== 2 * kChunkPairs)

struct Geometry {
    unsigned long long n = 0, nchunks = 0;
    unsigned P = 1, S = kV;
    int world = 1, rank = 0;
    unsigned seg0 = 0, nseg_local = kV, local_vshards = kV;
    unsigned long long chunk0 = 0, chunk1 = 0;   // this rank's chunk range
    unsigned long long j0 = 0, n_local = 0;      // this rank's variable range
    unsigned long long ld = 0;                   // padded local length: whole chunks (multiple of 512 doubles)

    // tuning constants of choose_P (process-wide)
    static unsigned &fill_div() { static unsigned v = 888; return v; }      // rule 0: groups wanted before groups start to grow
    static unsigned &group_base() { static unsigned v = 440; return v; }    // rule 1: see below
    static int &rule() { static int v = 1; return v; }
    static unsigned &min_group_chunks() { static unsigned v = NB200_MIN_GROUP_CHUNKS; return v; }   // rule 1: smallest group, in chunks

    static unsigned long long cut(unsigned s, unsigned long long nchunks, unsigned S)
    {
        return (unsigned long long) s * nchunks / S;
    }

    // Number of groups S = 8 P as a function of n alone.
    // Rule 1 (default).  The persistent kernels run 3 CTAs on each of the 148 SMs: 443 sweepers + the folder.  A rank
    // owns S / world groups per generation, so S is chosen from base * {1, 2, 4, 8} with base = 440: 1, 2, 4 or 8 ranks
    // then see a whole number of sweeper "waves" (3520 groups: 7.95 / 3.97 / 1.99 / 0.99 waves) -- the largest such S
    // that keeps a group at min_group_chunks = 4 chunks or more (measured, profiles/r02b_call3_geometry_async.txt: with
    // 2 the mid-size problems ran two half-length groups per CTA and paid the group boundary twice -- n = 1e6, m = 4:
    // 16.8 -> 14.2 us per evaluation, n = 1.25e6: 18.6 -> 16.5, n = 5e6: 61.7 -> 59.2).  Large n (groups would exceed `target_chunks` chunks): S = nchunks /
    // target_chunks, small groups even out the tail of a generation.  Small n: one chunk per group.
    // Rule 0 (round 1): about 1000 groups for mid-size n, `target_chunks` chunks per group for large n.
    static unsigned choose_P(unsigned long long nchunks, unsigned target_chunks, unsigned pmax, unsigned group_base_arg = 0)
    {
        const unsigned gbase = group_base_arg ? group_base_arg : group_base();
        unsigned long long want;
        if (rule() == 0) {
            unsigned long long fill = nchunks / fill_div();
            if (fill < 1) fill = 1;
            if (fill < target_chunks) target_chunks = (unsigned) fill;
            want = (nchunks + (unsigned long long) kV * target_chunks - 1) / ((unsigned long long) kV * target_chunks);
        } else {
            const unsigned long long base = gbase / kV ? gbase / kV : 1;                    // P of the smallest candidate
            const unsigned long long big = (nchunks + (unsigned long long) kV * target_chunks - 1) / ((unsigned long long) kV * target_chunks);
            if (big >= 8 * base) want = big;
            else {
                want = 0;
                for (unsigned long long f = 8; f >= 1; f >>= 1)
                    if (kV * base * f * min_group_chunks() <= nchunks) { want = base * f; break; }
                if (!want) {                                                              // at most one chunk per group
                    const unsigned long long one = (nchunks + kV - 1) / kV;
                    want = one < base ? one : base;
                }
            }
        }
        if (want < 1) want = 1;
        if (want > pmax) want = pmax;
        return (unsigned) want;
    }

    // group_base = 0: the process-wide default (rule 1: number of sweeper CTAs of the kernel that will run, rounded down
    // to a multiple of 8; the backends pass the value that matches their solve kernel -- it must be the same on all ranks)
    static Geometry make(unsigned long long n, int world, int rank, unsigned target_chunks, unsigned pmax, unsigned group_base_arg = 0)
    {
        Geometry g;
        g.n = n;
        const unsigned long long npairs = (n + 1) / 2;
        g.nchunks = (npairs + kChunkVars / 2 - 1) / (kChunkVars / 2);
        if (g.nchunks < 1) g.nchunks = 1;
        g.P = choose_P(g.nchunks, target_chunks, pmax, group_base_arg);
        g.S = kV * g.P;
        g.world = world;
        g.rank = rank;
        g.local_vshards = kV / (unsigned) world;
        g.nseg_local = g.local_vshards * g.P;
        g.seg0 = (unsigned) rank * g.nseg_local;
        g.chunk0 = cut(g.seg0, g.nchunks, g.S);
        g.chunk1 = cut(g.seg0 + g.nseg_local, g.nchunks, g.S);
        g.j0 = g.chunk0 * kChunkVars;
        if (g.j0 > n) g.j0 = n;
        unsigned long long j1 = g.chunk1 * kChunkVars;
        if (j1 > n) j1 = n;
        g.n_local = j1 > g.j0 ? j1 - g.j0 : 0;
        g.ld = (g.chunk1 - g.chunk0) * kChunkVars;
        if (g.ld == 0) g.ld = kChunkVars;
        return g;
    }
};

constexpr unsigned kDefaultTargetChunks = 8;     // ~4096 variables per group
constexpr unsigned kDefaultPmax = 1u << 16;

}  // namespace nb200
