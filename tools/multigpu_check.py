"""Multi-GPU parity of the sharded path (run on the GPU box).

  python tools/multigpu_check.py single             -> gpurun_out/mg_single.json  (1 GPU: sums + optimisation results)
  torchrun --nproc-per-node N tools/multigpu_check.py sharded   -> compares with mg_single.json

Checks: (i) the m+3 sums of a dual evaluation are BIT-IDENTICAL for world = 1 and world = N (fixed
cuts, fixed fold order); (ii) x*(y) gathered from the shards equals the single-GPU x*(y) bit for bit;
(iii) short CCSAQ / MMA runs with device callbacks -- the separable quadratic problem and the chained Rosenbrock
problem of BASELINE config 3, whose stencil needs the one-element halo exchange -- are BIT-IDENTICAL on every world
size: f*, the evaluation counts, the dual-evaluation count and the xor-hash of x* (the device callbacks reduce over the
same n-only groups and virtual shards as the dual kernels, include/nlopt_b200_device.cuh)."""
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


def opt_case(alg_name, n, problem="quadratic"):
    import nlopt_b200 as nl
    from nlopt_b200.problems import Problem, rosen_x0
    import torch
    alg = getattr(nl, alg_name)
    o = nl.opt(alg, n)
    p = Problem()
    L = o._lib
    j0, cnt = C.c_ulonglong(), C.c_ulonglong()
    L.nlopt_b200_shard_range(n, L.nlopt_b200_comm_rank(), L.nlopt_b200_comm_world(), C.byref(j0), C.byref(cnt))
    if problem == "quadratic":
        o.set_lower_bounds(-1.0); o.set_upper_bounds(1.0)
        p.quadratic_device(o)
        x = torch.full((cnt.value,), -0.5, dtype=torch.float64, device="cuda")
    else:
        o.set_lower_bounds(-2.0); o.set_upper_bounds(2.0)
        p.rosenbrock_device(o, 4)
        x = torch.from_numpy(rosen_x0(n)[j0.value:j0.value + cnt.value].copy()).cuda()
    o.set_maxeval(12)
    o.optimize_device(x.data_ptr())
    st = o.get_stats()
    xh = int(np.bitwise_xor.reduce(x.cpu().numpy().view(np.uint64))) if cnt.value else 0
    return dict(f=o.last_optimum_value().hex(), ret=o.last_optimize_result(), evals=o.get_numevals(),
                dual_evals=st["dual_evals"], xhash=xh)


OPT_CASES = [("LD_CCSAQ", 2_000_000, "quadratic"), ("LD_MMA", 2_000_000, "quadratic"), ("LD_CCSAQ", 3_000_001, "rosenbrock"),
             ("LD_MMA", 1_500_000, "rosenbrock")]


def main():
    mode = sys.argv[1]
    if mode == "single":
        res = {"dual": [dual_case(*c) for c in CASES], "opt": [opt_case(*c) for c in OPT_CASES]}
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
    for c, w in zip(OPT_CASES, want["opt"]):
        g = opt_case(*c)
        t = torch.tensor([g["xhash"] & 0x7FFFFFFFFFFFFFFF, g["xhash"] >> 63], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        hx = 0
        for tt in gathered:
            hx ^= int(tt[0].item()) | (int(tt[1].item()) << 63)
        same = g["f"] == w["f"] and g["ret"] == w["ret"] and g["evals"] == w["evals"] and g["dual_evals"] == w["dual_evals"] \
            and hx == w["xhash"]
        ok = ok and same
        if rank == 0:
            print("opt", c, "bit-identical to 1 GPU:", same, g["f"], w["f"], g["dual_evals"], w["dual_evals"], flush=True)
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL", "world", world, flush=True)
    L.nlopt_b200_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
