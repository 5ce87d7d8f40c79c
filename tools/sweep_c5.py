"""BASELINE config 5: sweep n = 1e3..1e8, m in {1, 4, 16}: dual-evaluation throughput and achieved HBM GB/s of the hot
path at 1/2/4/8 GPUs next to the reference's CPU dual function on the host cores.

  python tools/sweep_c5.py                       # 1 GPU  (adds the CPU reference per point, n <= 1e7)
  torchrun --nproc-per-node N tools/sweep_c5.py  # N GPUs (one process group, all points inside)

One point = one whole dual solve (mma.c:275-288) on the deterministic synthetic instance of SURVEY.md 8(d), generated on
the device (bit-identical to tests/synth.py), through the kernel-level C ABI nlopt_b200_dual_solve: warm start
y_i = 0.5 (i + 1), ftol_rel = 0, maxeval = 60 -> 60 evaluations + the final one that stores x*(y).  m <= 16 with the
fused persistent kernel where the library uses it (everything except MMA with m > 8, which runs one TMA-staged launch per
evaluation).  Reported per point: microseconds per dual evaluation (CUDA events around the kernels, max over ranks), whole-job
dual-evals/s, per-GPU GB/s = 8 n_local (5 + m) / time and its fraction of MEASURED_PEAKS.json hbm_gbs, and -- 1 GPU only -- the
reference's dual_func on one host core for the same instance (median of 3, n <= 1e7).  Output: one JSON per world size,
gpurun_out/sweep_c5_n{N}.json (copy to profiles/)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    import torch
    from nlopt_b200._capi import default_library, c_double_p
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    L = default_library()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            assert L.nlopt_b200_comm_unique_id(raw) == 0
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        assert L.nlopt_b200_comm_init(bytes(idbuf.cpu().numpy().tobytes()), rank, world, local) == 0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        peak = 6650.0
    ns = [int(float(v)) for v in os.environ.get("SWEEP_N", "1e3,1e4,1e5,1e6,1e7,1e8").split(",")]
    ms_ = [int(v) for v in os.environ.get("SWEEP_M", "1,4,16").split(",")]
    P = lambda v: v.ctypes.data_as(c_double_p)      # noqa: E731
    rows = []
    cpu_timer = None
    for variant, vname in ((0, "LD_MMA"), (1, "LD_CCSAQ")):
        for n in ns:
            for m in ms_:
                h = L.nlopt_b200_dual_create(variant, n, m)
                if not h:
                    continue
                for kv in filter(None, os.environ.get("SWEEP_CFG", "").split(",")):      # e.g. SWEEP_CFG=solve_async=3,group_base=288
                    k, v = kv.split("=")
                    if L.nlopt_b200_dual_configure(h, k.encode(), int(v)) != 0:
                        print("configure failed", kv, L.nlopt_b200_dual_errmsg(h).decode(), flush=True)
                L.nlopt_b200_dual_fill_synthetic(h, 0x5EED0000)
                i = np.arange(m, dtype=np.float64)
                c0, rhoc = -0.1 * (i + 1.0), 1.0 + 0.1 * i
                L.nlopt_b200_dual_set_scalars(h, 1.0, 1.0, P(c0), P(rhoc))
                lo, hi = np.zeros(m), np.full(m, 1e40)
                out = np.zeros(3 + m)
                best = None
                for rep in range(3):
                    y = 0.5 * (i + 1.0)
                    res, nev, kms = C.c_int(0), C.c_long(0), C.c_double(0.0)
                    if L.nlopt_b200_dual_solve(h, P(y), P(lo), P(hi), 0.0, 60, P(out), C.byref(res), C.byref(nev), C.byref(kms)) != 0:
                        print("solve failed", L.nlopt_b200_dual_errmsg(h).decode(), flush=True)
                        break
                    us = 1e3 * kms.value / max(1, nev.value)
                    if rep and (best is None or us < best[0]):
                        best = (us, nev.value, res.value, float(out[0]))
                nl = L.nlopt_b200_dual_query(h, b"n_local")
                groups = L.nlopt_b200_dual_query(h, b"segments")
                L.nlopt_b200_dual_destroy(h)
                if best is None:
                    continue
                us = best[0]
                if dist is not None:
                    t = torch.tensor([us], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    us = float(t.item())
                gbs = 8.0 * nl * (5 + m) / us * 1e-3
                row = dict(alg=vname, n=n, m=m, n_gpus=world, us_per_eval=us, evals_per_s=1e6 / us, gbs_per_gpu=gbs, frac_of_peak=gbs / peak,
                           evals=best[1], result=best[2], value=best[3], groups=groups)
                if world == 1 and n <= 10_000_000 and os.environ.get("SWEEP_CPU", "1") == "1":
                    import bench
                    import synth
                    inst = synth.kernel_instance(n, m)
                    tm = bench.RefDualTimer(inst, variant)
                    ts = tm.run(3, warm=1)
                    row["cpu_ms_per_eval"] = 1e3 * float(np.median(ts))
                    row["cpu_kind"] = tm.kind
                    row["speedup_vs_cpu"] = row["cpu_ms_per_eval"] * 1e3 / us
                    del inst, tm
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
    if rank == 0:
        summary = {}
        for vname in ("LD_MMA", "LD_CCSAQ"):
            for m in ms_:
                pts = [r for r in rows if r["alg"] == vname and r["m"] == m and "cpu_ms_per_eval" in r]
                cross = [r["n"] for r in pts if r["speedup_vs_cpu"] > 1.0]
                if pts:
                    summary[f"{vname}_m{m}_crossover_n"] = min(cross) if cross else None
        outp = os.path.join(ROOT, "gpurun_out", f"sweep_c5_n{world}{os.environ.get('SWEEP_TAG', '')}.json")
        os.makedirs(os.path.dirname(outp), exist_ok=True)
        json.dump({"n_gpus": world, "peak_gbs": peak, "rows": rows, "crossover": summary,
                   "what": "one whole dual solve of 60 + 1 evaluations per point (nlopt_b200_dual_solve), best of 2 after a warm-up"}, open(outp, "w"), indent=1)
        print("wrote", outp, json.dumps(summary), flush=True)
    if world > 1:
        L.nlopt_b200_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
