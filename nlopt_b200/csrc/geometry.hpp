// geometry.hpp -- how the n variables are cut into segments, virtual shards and rank shards.
//
// Everything here is a function of n alone (plus the two tuning constants), never of the number
// of GPUs: S = 8 * P segments, segment s covers the double2 "pairs" [s*NP/S, (s+1)*NP/S) with
// NP = ceil(n/2); virtual shard v = segments [v*P, (v+1)*P); rank r of `world` (1, 2, 4 or 8)
// owns virtual shards [r*8/world, (r+1)*8/world), i.e. one contiguous block of variables
// starting at an even index.  Fixed cuts + fixed fold order = sums that are bit-identical for
// every world size.
#pragma once

#include <cstdint>

namespace nb200 {

constexpr unsigned kV = 8;   // virtual shards (== kVirtualShards in ccsa_kernels.cuh)

struct Geometry {
    unsigned long long n = 0, npairs = 0;
    unsigned P = 1, S = kV;
    int world = 1, rank = 0;
    unsigned seg0 = 0, nseg_local = kV, local_vshards = kV;
    unsigned long long pair0 = 0, pair1 = 0;   // this rank's pair range
    unsigned long long j0 = 0, n_local = 0;    // this rank's variable range
    unsigned long long ld = 0;                 // padded local length (multiple of 32 doubles)

    static unsigned long long cut(unsigned s, unsigned long long npairs, unsigned S)
    {
        return (unsigned long long) s * npairs / S;
    }

    // P grows with n until a segment holds about `target_pairs` pairs, capped at pmax
    static unsigned choose_P(unsigned long long npairs, unsigned target_pairs, unsigned pmax)
    {
        unsigned long long want = (npairs + (unsigned long long) kV * target_pairs - 1) / ((unsigned long long) kV * target_pairs);
        if (want < 1) want = 1;
        if (want > pmax) want = pmax;
        return (unsigned) want;
    }

    static Geometry make(unsigned long long n, int world, int rank, unsigned target_pairs, unsigned pmax)
    {
        Geometry g;
        g.n = n;
        g.npairs = (n + 1) / 2;
        g.P = choose_P(g.npairs, target_pairs, pmax);
        g.S = kV * g.P;
        g.world = world;
        g.rank = rank;
        g.local_vshards = kV / (unsigned) world;
        g.nseg_local = g.local_vshards * g.P;
        g.seg0 = (unsigned) rank * g.nseg_local;
        g.pair0 = cut(g.seg0, g.npairs, g.S);
        g.pair1 = cut(g.seg0 + g.nseg_local, g.npairs, g.S);
        g.j0 = 2 * g.pair0;
        unsigned long long j1 = 2 * g.pair1;
        if (j1 > n) j1 = n;
        g.n_local = j1 > g.j0 ? j1 - g.j0 : 0;
        unsigned long long padded = 2 * (g.pair1 - g.pair0);
        g.ld = (padded + 31) / 32 * 32;
        if (g.ld == 0) g.ld = 32;
        return g;
    }
};

constexpr unsigned kDefaultTargetPairs = 1024;   // ~2048 variables per CTA before P saturates
constexpr unsigned kDefaultPmax = 296;           // 8 * 296 = 2368 = 16 * 148 CTAs on one GPU

}  // namespace nb200
