"""Multi-GPU parity of the sharded path (run on the GPU box).

  python tools/multigpu_check.py single             -> gpurun_out/mg_single.json  (1 GPU: sums + optimisation results)
  torchrun --nproc-per-node N tools/multigpu_check.py sharded   -> compares with mg_single.json

Checks: (i) the m+3 sums of a dual evaluation are BIT-IDENTICAL for world = 1 and world = N (fixed
cuts, fixed fold order); (ii) x*(y) gathered from the shards equals the single-GPU x*(y) bit for bit;
(iii) short CCSAQ / MMA runs with device callbacks on the separable quadratic problem stay on the same
trajectory on every world size (replicated host logic fed by identical dual sums; the user objective's own
reduction differs in rounding across world sizes and the optimiser amplifies that, so f is compared to 1e-7
relative; the mailbox and the NCCL exchange give bit-identical f at the same world size)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "mg_single.json")
CASES = [(0, 1_000_003, 4), (1, 1_000_003, 4), (1, 3_000_000, 1), (0, 500_000, 16)]


def dual_case(variant, n, m):
    from gpu_dual import DualHandle
    import synth
    h = DualHandle(variant, n=n, m=m, synthetic_seed=synth.SEED0)
    i = np.arange(m, dtype=float)
    h.set_scalars(1.0, 1.0, -0.1 * (i + 1), 1.0 + 0.1 * i)
    r = h.eval(0.5 * (i + 1), want_xcur=True)
    j0, cnt = h.query("j0"), h.query("n_local")
    xc = r["xcur"][j0:j0 + cnt]
    return dict(ret=r["ret"].hex(), g0=r["g0"].hex(), w=r["w"].hex(), gc=[v.hex() for v in r["gc"]],
                xsum=float(np.sum(xc)), xhash=int(np.bitwise_xor.reduce(xc.view(np.uint64))), j0=j0, cnt=cnt)


def opt_case(alg_name, n):
    import nlopt_b200 as nl
    from nlopt_b200.problems import Problem
    import torch
    alg = getattr(nl, alg_name)
    o = nl.opt(alg, n)
    o.set_lower_bounds(-1.0); o.set_upper_bounds(1.0)
    p = Problem()
    p.quadratic_device(o)
    o.set_maxeval(12)
    L = o._lib
    j0, cnt = C.c_ulonglong(), C.c_ulonglong()
    L.nlopt_b200_shard_range(n, L.nlopt_b200_comm_rank(), L.nlopt_b200_comm_world(), C.byref(j0), C.byref(cnt))
    x = torch.full((cnt.value,), -0.5, dtype=torch.float64, device="cuda")
    o.optimize_device(x.data_ptr())
    st = o.get_stats()
    return dict(f=o.last_optimum_value().hex(), ret=o.last_optimize_result(), evals=o.get_numevals(),
                dual_evals=st["dual_evals"], xsum=float(x.sum().item()))


def main():
    mode = sys.argv[1]
    if mode == "single":
        res = {"dual": [dual_case(*c) for c in CASES], "opt": [opt_case("LD_CCSAQ", 2_000_000), opt_case("LD_MMA", 2_000_000)]}
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        json.dump(res, open(OUT, "w"), indent=1)
        print("single-GPU results written:", json.dumps(res["opt"]))
        return
    import torch
    import torch.distributed as dist
    from nlopt_b200._capi import default_library
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = default_library()
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        assert L.nlopt_b200_comm_unique_id(raw) == 0
        idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    assert L.nlopt_b200_comm_init(bytes(idbuf.cpu().numpy().tobytes()), rank, world, local) == 0
    want = json.load(open(OUT))
    ok = True
    for c, w in zip(CASES, want["dual"]):
        g = dual_case(*c)
        same = g["ret"] == w["ret"] and g["g0"] == w["g0"] and g["w"] == w["w"] and g["gc"] == w["gc"]
        # x*(y): xor of the shard's bit patterns, combined over ranks, equals the single-GPU xor
        t = torch.tensor([g["xhash"] & 0x7FFFFFFFFFFFFFFF, g["xhash"] >> 63], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        hx = 0
        for tt in gathered:
            lo, hi = int(tt[0].item()), int(tt[1].item())
            hx ^= lo | (hi << 63)
        same = same and hx == w["xhash"]
        ok = ok and same
        if rank == 0:
            print("dual", c, "bit-identical to 1 GPU:", same, flush=True)
    for name, w in zip(("LD_CCSAQ", "LD_MMA"), want["opt"]):
        g = opt_case(name, 2_000_000)
        # the user's objective reduction (map_reduce_kernel + all-reduce) is not world-size independent,
        # so f differs in the last bits; the solver path fed by it must stay on the same trajectory
        f1, fN = float.fromhex(w["f"]), float.fromhex(g["f"])
        same = abs(fN - f1) <= 1e-7 * abs(f1) and g["ret"] == w["ret"] and g["evals"] == w["evals"] \
            and abs(g["dual_evals"] - w["dual_evals"]) <= 0.1 * w["dual_evals"] + 2
        ok = ok and same
        if rank == 0:
            print("opt", name, "identical to 1 GPU:", same, g, w, flush=True)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL", "world", world, flush=True)
    L.nlopt_b200_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
