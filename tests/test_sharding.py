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
