#!/usr/bin/env python
"""bench.py -- CCSA dual-evaluations per second (BASELINE.json metric) on 1..8 B200.

Workload (config.workload): BASELINE configs[2] -- NLOPT_LD_CCSAQ on the n = 1e7 chained-Rosenbrock
problem with m = 4 dense linear inequality constraints, fp64, box [-2,2]^n.

  step      = one inner CCSA iteration: one dual solve (K_i dual evaluations -- by default inside ONE
              persistent dual_solve_kernel launch that also runs the m-dimensional dual optimiser; with
              b200_fused_solve=0 K_i launches of dual_eval_kernel driven from the host), the final
              evaluation that materialises x*(y), one objective + constraint evaluation, acceptance.
              `--steps K` runs exactly K of them (maxeval = K + 1), `--warmup W` a separate
              W-iteration run first.
  value     = dual evaluations / second with everything resident in HBM (__device__ objective and
              constraints, x on the device): dual evaluations performed in the K steps /
              device-timed duration of the nlopt_b200_optimize_device call (setup, objective and
              constraint evaluations, acceptance and stopping tests included).
  e2e       = the same through plain nlopt_optimize(): HOST x, HOST callbacks (C functions of
              libnlopt_b200_problems.so); every step pulls x*(y) to pinned host memory and pushes
              (1+m) gradient rows back.  Rate = dual evaluations / (wall - time inside the user's
              callbacks), the definition BASELINE.md uses for the reference; the rate including
              callback time is reported next to it.
  roofline  = dominant kernel (dual_solve_kernel / dual_eval_kernel): algorithmic bytes 8 n (5+m) per
              dual evaluation (+8 n on the evaluations that store x*), divided by the CUDA-event
              duration of those kernels measured inside the timed region (for the persistent kernel
              this includes its in-kernel optimiser steps and generation hand-offs), against
              MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference = the reference's own dual_func (oracle/_ref, include-trick on
              the unmodified src/algs/mma/ccsa_quadratic.c) on the same arrays, one thread.

Launch: python bench.py [--gpus N --steps K --warmup W] ; for N > 1 under torchrun (one rank per GPU).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "ccsa_dual_evals_per_sec"
UNIT = "dual-evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--m", type=int, default=4)
    ap.add_argument("--alg", default="ccsaq", choices=["ccsaq", "mma"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--param", action="append", default=[], help="extra nlopt_set_param name=value (tuning experiments)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed region runs

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        return len(self.rows)

    def stop(self, i0=0, i1=None):
        """summary of the samples taken between two mark()s (the timed region); if the region was too short to
        catch one, the nearest samples around it are used and `window` says so"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        i1 = len(self.rows) if i1 is None else i1
        rows, window = self.rows[i0:i1], "timed region"
        if not rows:
            rows, window = self.rows[max(0, i0 - 3):i1 + 3], "nearest samples around the (short) timed region"
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own dual function on the same state, host cores

def c3_state_at_x0(n, m):
    """The arrays one dual evaluation of the first outer iteration reads for the config-3 instance:
    x0, box, sigma_0 = (ub-lb)/2, grad f(x0), the m weight rows; f0, c0 at x0; rho = rhoc = 1."""
    from nlopt_b200.problems import linear_weights, rosen_x0
    x = rosen_x0(n)
    lb, ub = np.full(n, -2.0), np.full(n, 2.0)
    d = x[1:] - x[:-1] ** 2
    e = 1.0 - x[:-1]
    g = np.zeros(n)
    g[:-1] += -400.0 * x[:-1] * d - 2.0 * e
    g[1:] += 200.0 * d
    f0 = float(np.sum(100.0 * d * d + e * e))
    G = np.empty((m, n))
    c0 = np.empty(m)
    for k in range(m):
        G[k] = linear_weights(k, n)
        c0[k] = float(np.dot(G[k], x)) - (0.5 + 0.1 * k)
    return dict(n=n, m=m, x=x, lb=lb, ub=ub, sigma=0.5 * (ub - lb), grad_f=g, grad_c=G, f0=f0, rho=1.0,
                c0=c0, rhoc=np.ones(m), y=np.zeros(m))


def cpu_dual_rate(n, m, variant, evals, warm=1):
    """dual-evals/s of the reference's static dual_func (oracle/_ref) -- or of the oracle port when the
    reference could not be built -- on one host core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_bindings as ob
    inst = c3_state_at_x0(n, m)
    kind = "reference" if ob.ref_dual_available() else "port"
    fn = ob.ref_dual if kind == "reference" else ob.port_dual
    rng = np.random.default_rng(0)
    ys = [np.abs(rng.standard_normal(m)) * 10.0 for _ in range(evals + warm)]
    for y in ys[:warm]:
        fn(variant, inst, y)
    times = []
    for y in ys[warm:]:
        t0 = time.perf_counter()
        fn(variant, inst, y)
        times.append(time.perf_counter() - t0)
    return kind, times


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    variant = 1 if a.alg == "ccsaq" else 0
    kind, times = cpu_dual_rate(a.n, a.m, variant, a.steps, a.warmup)
    total = float(np.sum(times))
    val = len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, step="one evaluation of the reference's dual_func (ccsa_quadratic.c:79-148 / "
                                  "mma.c:59-137) on the n-variable state at x0 -- the bounded sample of the workload"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"{len(times)} dual evaluations at n={a.n}, m={a.m}, single thread (the reference is single-threaded)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


def workload_config(a, step):
    return {"workload": f"NLOPT_LD_{a.alg.upper()} n={a.n} chained-Rosenbrock + {a.m} dense linear inequality "
                        f"constraints, box [-2,2]^n, fp64 (BASELINE configs[2])",
            "n": a.n, "m": a.m, "algorithm": "LD_" + a.alg.upper(), "step": step,
            "l2": f"each dual evaluation streams {8 * a.n * (5 + a.m) / 1e6:.0f} MB (> 126 MB L2); no flush needed",
            "sharding": "contiguous blocks of variables, one rank per GPU"}


# ------------------------------------------------------------------------------------------------------

def main():
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    import torch.distributed as dist
    import nlopt_b200 as nl
    from nlopt_b200._capi import default_library
    from nlopt_b200.problems import Problem, rosen_x0

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    L = default_library()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            assert L.nlopt_b200_comm_unique_id(raw) == 0
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        assert L.nlopt_b200_comm_init(bytes(idbuf.cpu().numpy().tobytes()), rank, world, local) == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, m = a.n, a.m
    alg = nl.LD_CCSAQ if a.alg == "ccsaq" else nl.LD_MMA
    j0, cnt = C.c_ulonglong(0), C.c_ulonglong(0)
    L.nlopt_b200_shard_range(n, rank, world, C.byref(j0), C.byref(cnt))
    x0 = rosen_x0(n)

    def make_opt(device_callbacks):
        o = nl.opt(alg, n)
        o.set_lower_bounds(-2.0)
        o.set_upper_bounds(2.0)
        p = Problem()
        if device_callbacks:
            p.rosenbrock_device(o, m)
        else:
            p.rosenbrock_host(o, m)
        o.set_param("b200_time_kernels", 1)
        for kv in a.param:
            k, v = kv.split("=")
            o.set_param(k, float(v))
        return o, p

    def timed(run, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        run(steps)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (t1 - t0) * 1e3)   # the call is host-synchronous; both agree
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    # ---- value: device-resident ------------------------------------------------------------------------
    od, pd = make_opt(True)
    x0dev = torch.from_numpy(x0[j0.value:j0.value + cnt.value].copy()).cuda()     # inputs are resident in HBM before the timed region
    xdev = x0dev.clone()

    def run_dev(steps):
        xdev.copy_(x0dev)
        od.set_maxeval(steps + 1)
        od.optimize_device(xdev.data_ptr())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()            # nvidia-smi needs ~0.1 s to deliver its first sample: start before the warm-up
    run_dev(a.warmup)
    i0 = sampler.mark()
    ms_dev = timed(run_dev, a.steps)
    i1 = sampler.mark()
    clocks = sampler.stop(i0, i1) if rank == 0 else None
    sd = od.get_stats()
    f_dev = od.last_optimum_value()
    value = sd["dual_evals"] / (ms_dev * 1e-3)
    n_local = cnt.value
    kern_bytes = 8.0 * n_local * ((5 + m) * sd["dual_evals"] + sd["dual_solves"])
    kern_s = sd["seconds_dual_kernel"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = kern_bytes / kern_s / 1e9 if kern_s > 0 else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    launches_dev = sd["kernel_launches"] + od.get_numevals() * (1 + m)

    # ---- e2e: host buffers + host callbacks through nlopt_optimize ---------------------------------------
    e2e = None
    if not a.no_e2e:
        oh, ph = make_opt(False)

        def run_host(steps):
            oh.set_maxeval(steps + 1)
            run_host.x = oh.optimize(x0)

        run_host(max(1, min(a.warmup, 2)))
        ph.reset_callback_seconds()
        ms_host = timed(run_host, a.steps)
        sh = oh.get_stats()
        cb_s = sh["seconds_callbacks"]
        if world > 1:      # ranks wait for the slowest rank's callbacks inside the next dual solve
            t = torch.tensor([cb_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cb_s = float(t.item())
        solver_s = ms_host * 1e-3 - cb_s
        e2e = {"value": sh["dual_evals"] / solver_s, "unit": UNIT,
               "h2d_bytes_per_step": sh["h2d_bytes"] / a.steps, "d2h_bytes_per_step": sh["d2h_bytes"] / a.steps,
               "value_incl_user_callbacks": sh["dual_evals"] / (ms_host * 1e-3),
               "seconds_in_user_callbacks": cb_s, "dual_evals": sh["dual_evals"],
               "note": "rate over wall time minus time inside the user's host callbacks (BASELINE.md definition); "
                       "includes all H2D/D2H copies, launches and the host-side dual optimiser",
               "f_after_steps": oh.last_optimum_value(),
               "wall_breakdown_s": {k: sh[k] for k in ("seconds_total", "seconds_setup", "seconds_dual_wall", "seconds_eval_wall",
                                                        "seconds_glue_wall", "seconds_callbacks")}}

    # ---- cpu baseline: reference dual_func on host cores (rank 0, N = 1 only) ----------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        kind, times = cpu_dual_rate(n, m, 1 if a.alg == "ccsaq" else 0, evals=24, warm=1)
        cpu = {"value": len(times) / float(np.sum(times)), "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"{len(times)} evaluations of the reference's dual_func at n={n}, m={m} (~{np.sum(times):.1f} s), "
                         f"single thread -- the reference has no threading; host has {os.cpu_count()} cores"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, step="one inner CCSA iteration (dual solve + x*(y) + candidate evaluation)"),
            "dual_evals": sd["dual_evals"], "dual_solves": sd["dual_solves"], "f_after_steps": f_dev,
            "wall_breakdown_s": {k: sd[k] for k in ("seconds_total", "seconds_setup", "seconds_dual_wall", "seconds_eval_wall",
                                                     "seconds_glue_wall", "seconds_dual_kernel")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": "dual_solve_kernel (persistent; sweeps = dual_eval_kernel body)", "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                         "avg_launch_us": 1e6 * kern_s / max(1, sd["dual_evals"]),
                         "kernel_share_of_step": kern_s / (ms_dev * 1e-3)},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches_dev),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        L.nlopt_b200_comm_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
