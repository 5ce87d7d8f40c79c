"""Full-size goldens from the UNMODIFIED reference (oracle/_ref), generated in the build container and committed as
tests/golden/reference_golden_big.json (the GPU box has no /root/reference):

    python tests/golden/make_golden_big.py            # ~10 minutes of single-threaded reference time

  dual_big[] : the reference's static dual_func at the BASELINE size n = 1e7, m = 4 (both variants) on the synthetic
               instance of tests/synth.py: sums as C99 hex floats; x*(y) as xor-hash, plain sum and 8 samples.
  c3[]       : BASELINE config 3 (NLOPT_LD_CCSAQ, n = 1e7 chained Rosenbrock + 4 dense linear constraints) through the
               reference's nlopt_optimize with the plain-C host callbacks of libnlopt_b200_problems.so (the same
               callbacks bench.py's e2e arm registers): f after K inner iterations (maxeval = K + 1) for the K's
               bench.py is run with, evaluation and dual-evaluation counts (verbosity = 1 output, mma.c:288-291).
  c4[]       : BASELINE config 4's problem (NLOPT_LD_MMA, synthetic SIMP compliance + volume constraint, host
               callbacks) at n = 1e6: converged run (xtol_rel = 1e-6) and a fixed 12-iteration run.
"""
import json
import os
import re
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, ROOT)
sys.path.insert(0, TESTS)
import numpy as np  # noqa: E402

import nlopt_b200 as nl  # noqa: E402
import oracle_bindings as ob  # noqa: E402
import synth  # noqa: E402
from nlopt_b200 import Library  # noqa: E402
from nlopt_b200.problems import Problem, rosen_x0  # noqa: E402


def xhash(x):
    return int(np.bitwise_xor.reduce(np.ascontiguousarray(x).view(np.uint64))) if x.size else 0


class CaptureStdout:
    """collect what C code prints to fd 1 (the reference's verbosity output)"""

    def __enter__(self):
        sys.stdout.flush()
        self.tmp = tempfile.TemporaryFile(mode="w+b")
        self.saved = os.dup(1)
        os.dup2(self.tmp.fileno(), 1)
        return self

    def __exit__(self, *a):
        import ctypes
        ctypes.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.tmp.seek(0)
        self.text = self.tmp.read().decode(errors="replace")
        self.tmp.close()


def dual_counts(text):
    return [int(v) for v in re.findall(r"dual converged in (\d+) iter", text)]


def main():
    assert ob.ref_dual_available(), "build oracle/_ref first (python __graft_entry__.py)"
    ref = Library(ob.REF_SO, extensions=False)
    out = {"generator": "tests/golden/make_golden_big.py", "reference": "NLopt 2.11.0 (oracle/_ref, gcc -O3 -ffp-contract=off)",
           "dual_big": [], "c3": [], "c4": []}
    n, m = 10_000_000, 4
    inst = synth.kernel_instance(n, m)
    for variant in (ob.MMA, ob.CCSAQ):
        t0 = time.time()
        r = ob.ref_dual(variant, inst)
        idx = np.linspace(0, n - 1, 8).astype(int)
        out["dual_big"].append(dict(variant=variant, n=n, m=m, seed=synth.SEED0, ret=float(r["ret"]).hex(), g0=float(r["g0"]).hex(),
                                    w=float(r["w"]).hex(), gc=[float(v).hex() for v in r["gc"]], x_xor=xhash(r["xcur"]),
                                    x_sum=float(np.sum(r["xcur"])).hex(), x_idx=idx.tolist(),
                                    x_samples=[float(r["xcur"][i]).hex() for i in idx], seconds=time.time() - t0))
        print("dual_big", variant, time.time() - t0, flush=True)
    del inst

    x0 = rosen_x0(n)
    for alg_name, alg, ks in (("LD_CCSAQ", nl.LD_CCSAQ, (8, 20)), ("LD_MMA", nl.LD_MMA, (8,))):
        for K in ks:
            o = nl.opt(alg, n, library=ref)
            o.set_lower_bounds(-2.0); o.set_upper_bounds(2.0)
            p = Problem()
            p.rosenbrock_host(o, m)
            o.set_maxeval(K + 1)
            o.set_param("verbosity", 1)
            p.reset_callback_seconds()
            t0 = time.time()
            with CaptureStdout() as cap:
                x = o.optimize(x0)
            wall = time.time() - t0
            counts = dual_counts(cap.text)
            out["c3"].append(dict(alg=alg_name, n=n, m=m, steps=K, maxeval=K + 1, ret=o.last_optimize_result(), numevals=o.get_numevals(),
                                  minf=float(o.last_optimum_value()).hex(), minf_dec=repr(float(o.last_optimum_value())),
                                  x_xor=xhash(x), x_sum=float(np.sum(x)).hex(), dual_counts=counts, dual_evals=int(sum(counts)),
                                  wall_s=wall, callback_s=p.callback_seconds()))
            print("c3", alg_name, K, o.last_optimum_value(), sum(counts), wall, flush=True)

    n4 = 1_000_000
    for kw in (dict(xtol_rel=1e-6, maxeval=400), dict(maxeval=13)):
        o = nl.opt(nl.LD_MMA, n4, library=ref)
        o.set_lower_bounds(0.0); o.set_upper_bounds(1.0)
        p = Problem()
        p.simp_host(o)
        for k, v in kw.items():
            getattr(o, "set_" + k)(v)
        o.set_param("verbosity", 1)
        t0 = time.time()
        with CaptureStdout() as cap:
            x = o.optimize(np.full(n4, 0.4))
        counts = dual_counts(cap.text)
        out["c4"].append(dict(alg="LD_MMA", n=n4, m=1, **kw, ret=o.last_optimize_result(), numevals=o.get_numevals(),
                              minf=float(o.last_optimum_value()).hex(), minf_dec=repr(float(o.last_optimum_value())), x_xor=xhash(x),
                              x_sum=float(np.sum(x)).hex(), x_first=[float(v).hex() for v in x[:4]], dual_evals=int(sum(counts)),
                              wall_s=time.time() - t0))
        print("c4", kw, o.last_optimum_value(), o.get_numevals(), sum(counts), flush=True)
    with open(os.path.join(HERE, "reference_golden_big.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote reference_golden_big.json")


if __name__ == "__main__":
    main()
