"""N > 1 on CPU (gloo, world_size 2): the sharding protocol of SURVEY.md 8(e).

Every rank owns 8/world of the 8 fixed virtual shards (contiguous blocks of variables whose cuts
depend on n only), produces one record of m+3 sums per virtual shard, the records are all-gathered
and every rank folds the 8 records in index order.  The test checks, with the oracle standing in
for the per-shard kernel, that (i) the cuts reported by the library tile [0, n) identically for
world = 1, 2, 4, 8, (ii) the gathered fold is bit-identical to the single-process fold, and
(iii) it agrees with one un-sharded oracle evaluation to rounding."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_bindings as ob
import synth
from nlopt_b200 import _capi

N, M = 100003, 3


def shard(n, rank, world):
    L = _capi.default_library()
    j0, cnt = C.c_ulonglong(), C.c_ulonglong()
    L.nlopt_b200_shard_range(n, rank, world, C.byref(j0), C.byref(cnt))
    return j0.value, cnt.value


def vshard_record(variant, inst, v):
    """m+3 raw sums of virtual shard v (constants zeroed, as the device kernel produces them)."""
    j0, cnt = shard(inst["n"], v, 8)
    sl = slice(j0, j0 + cnt)
    sub = dict(inst, n=cnt, x=inst["x"][sl], lb=inst["lb"][sl], ub=inst["ub"][sl], sigma=inst["sigma"][sl],
               grad_f=inst["grad_f"][sl], grad_c=np.ascontiguousarray(inst["grad_c"][:, sl]), f0=0.0,
               c0=np.zeros(inst["m"]))
    r = ob.port_dual(variant, sub)
    return np.array([-r["ret"], r["g0"], r["w"], *r["gc"]])


def fold(records):
    acc = records[0].copy()
    for r in records[1:]:
        acc = acc + r
    return acc


def test_virtual_shards_tile_every_world_size(built):
    for n in (N, 10**7, 17):
        cuts8 = [shard(n, v, 8) for v in range(8)]
        for world in (1, 2, 4, 8):
            per = 8 // world
            for r in range(world):
                j0, cnt = shard(n, r, world)
                mine = cuts8[r * per:(r + 1) * per]
                assert j0 == mine[0][0] and cnt == sum(c for _, c in mine)


def _worker(rank, world, port, variant, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inst = synth.kernel_instance(N, M)
    per = 8 // world
    mine = np.stack([vshard_record(variant, inst, v) for v in range(rank * per, (rank + 1) * per)])
    gathered = [torch.zeros(per, 3 + M, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine))
    allrec = torch.cat(gathered).numpy()
    out[rank] = fold([allrec[v] for v in range(8)])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("variant", [ob.MMA, ob.CCSAQ])
def test_two_rank_fold_is_bit_identical_to_one_rank(built, variant):
    inst = synth.kernel_instance(N, M)
    single = fold([vshard_record(variant, inst, v) for v in range(8)])
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + variant
    mp.spawn(_worker, args=(2, port, variant, out), nprocs=2, join=True)
    assert np.array_equal(out[0], single) and np.array_equal(out[1], single)
    whole = ob.port_dual(variant, dict(inst, f0=0.0, c0=np.zeros(M)))
    ref = np.array([-whole["ret"], whole["g0"], whole["w"], *whole["gc"]])
    assert np.allclose(single, ref, rtol=1e-12, atol=1e-9)


# ---- the rank-count-independent protocol of the device callbacks (include/nlopt_b200.h: nlopt_b200_dfunc2) ----------
class ShardGeo(C.Structure):
    _fields_ = [("n", C.c_ulonglong), ("n_local", C.c_ulonglong), ("j0", C.c_ulonglong), ("nchunks", C.c_ulonglong),
                ("chunk0", C.c_ulonglong), ("groups_total", C.c_uint), ("group0", C.c_uint), ("groups_local", C.c_uint),
                ("groups_per_vshard", C.c_uint), ("vshard0", C.c_uint), ("local_vshards", C.c_uint), ("rank", C.c_int),
                ("world", C.c_int)]


def geometry(n, rank, world):
    g = ShardGeo()
    _capi.default_library().nlopt_b200_shard_geometry(n, rank, world, C.byref(g))
    return g


def group_range(g, k):
    """variables [lo, hi) of global group k, relative to the start of the rank that owns it"""
    lo = (k * g.nchunks // g.groups_total - g.chunk0) * 512
    hi = ((k + 1) * g.nchunks // g.groups_total - g.chunk0) * 512
    return lo, min(hi, g.n_local)


def test_shard_geometry_groups_tile_and_do_not_depend_on_the_world_size(built):
    for n in (1, 511, 513, 100003, 1250000, 10**7):
        g1 = geometry(n, 0, 1)
        assert g1.groups_total == 8 * g1.groups_per_vshard and g1.groups_local == g1.groups_total
        cover = 0
        for k in range(g1.groups_total):
            lo, hi = group_range(g1, k)
            assert lo == min(cover, max(lo, 0)) or lo >= cover
            cover = max(cover, hi)
        assert cover == n
        for world in (2, 4, 8):
            tot = 0
            for r in range(world):
                g = geometry(n, r, world)
                assert (g.groups_total, g.groups_per_vshard, g.nchunks) == (g1.groups_total, g1.groups_per_vshard, g1.nchunks)
                assert g.vshard0 == r * (8 // world) and g.local_vshards == 8 // world and g.group0 == g.vshard0 * g.groups_per_vshard
                for k in range(g.group0, g.group0 + g.groups_local):       # the same global variables as on one rank
                    lo, hi = group_range(g, k)
                    lo1, hi1 = group_range(g1, k)
                    assert (g.j0 + lo, g.j0 + max(hi, lo)) == (lo1, max(hi1, lo1))
                tot += g.n_local
            assert tot == n


def test_group_rule_wave_aligned_counts_and_smallest_group(built):
    """geometry.hpp rule 1: between one chunk per group (small n) and 8 chunks per group (large n) the number of groups is
    440 x {1, 2, 4, 8} -- whole sweeper waves of the 3-CTAs/SM persistent kernels on 1, 2, 4 or 8 ranks -- and the largest
    of these that keeps every group at 4 chunks or more."""
    for n, want in ((300_000, 440), (600_000, 440), (1_000_000, 440), (1_250_000, 440), (2_500_000, 880),
                    (5_000_000, 1760), (10_000_000, 3520)):
        g = geometry(n, 0, 1)
        assert g.groups_total == want, (n, g.groups_total)
        if n >= 1_000_000:
            sizes = [(k + 1) * g.nchunks // g.groups_total - k * g.nchunks // g.groups_total for k in range(g.groups_total)]
            assert min(sizes) >= 4
    for n in (20_000_000, 100_000_000):                           # large n: 8 chunks per group
        g = geometry(n, 0, 1)
        assert g.groups_total == (g.nchunks + 63) // 64 * 8
    for n in (1, 1000, 100_000):                                  # small n: one chunk per group (at least the 8 virtual shards)
        g = geometry(n, 0, 1)
        assert g.groups_total == 8 * max(1, (g.nchunks + 7) // 8)


def _value_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 100003
    x = synth.u01(7, n)
    g = geometry(n, rank, world)
    block = np.zeros((3, 8))                                     # [1 + m][8] virtual-shard sums, this rank's slots only
    for f in range(3):
        for v in range(g.vshard0, g.vshard0 + g.local_vshards):
            s = 0.0
            for k in range(v * g.groups_per_vshard, (v + 1) * g.groups_per_vshard):
                lo, hi = group_range(g, k)
                s += float(np.sum(x[g.j0 + lo:g.j0 + hi] ** (f + 1))) if hi > lo else 0.0
            block[f, v] = s
    t = torch.from_numpy(block)
    dist.all_reduce(t)                                           # every slot is non-zero on exactly one rank: exact
    out[rank] = [float(np.add.reduce(t.numpy()[f])) if False else float(sum_in_order(t.numpy()[f])) for f in range(3)]
    dist.barrier()
    dist.destroy_process_group()


def sum_in_order(a):
    s = a[0]
    for v in a[1:]:
        s = s + v
    return s


def test_device_callback_values_are_bit_identical_for_two_ranks(built):
    mgr = mp.Manager()
    one, two = mgr.dict(), mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_value_worker, args=(1, port, one), nprocs=1, join=True)
    mp.spawn(_value_worker, args=(2, port + 1, two), nprocs=2, join=True)
    assert list(two[0]) == list(one[0]) and list(two[1]) == list(one[0])
