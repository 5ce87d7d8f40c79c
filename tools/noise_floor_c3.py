"""How far do rounding-level perturbations move BASELINE config 3 (NLOPT_LD_CCSAQ, n = 1e7 chained Rosenbrock + 4 dense
linear constraints) after K inner iterations?  The end-to-end parity bar for this config is judged against this floor
(SURVEY.md 8(c) measured it at n = 1e3 only).  Runs, all on the CPU with the oracle port (bit-identical to the
reference, tests/test_oracle_port.py):
  ref_c      : the reference itself with the sequential-sum C callbacks        (tests/golden/reference_golden_big.json)
  port_np    : the port, numpy callbacks (pairwise sums inside f and the constraints: last-bit changes of f, c)
  port_wide  : the same, dual sums accumulated in 80-bit long double (-DPORT_WIDE_SUMS): a nearly exact summation
Output: profiles/r02_c3_noise_floor.json.     python tools/noise_floor_c3.py [K] [n]"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_bindings as ob  # noqa: E402
import problems as P  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
    m = 4
    wide = "/tmp/liboracle_port_wide.so"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DPORT_WIDE_SUMS", "-o", wide,
                           os.path.join(ROOT, "oracle", "ccsa_port.c"), "-lm"])
    cons = [P.lin_constraint(k, n) for k in range(m)]
    lb, ub = np.full(n, -2.0), np.full(n, 2.0)
    out = {"n": n, "m": m, "K": K, "algorithm": "LD_CCSAQ", "runs": {}}
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_golden_big.json")))
        for g in gold["c3"]:
            if g["alg"] == "LD_CCSAQ" and g["steps"] == K and g["n"] == n:
                out["runs"]["ref_c"] = dict(f=float.fromhex(g["minf"]), dual_evals=g["dual_evals"])
    except Exception:
        pass
    for name, lib in (("port_np", ob.PORT_SO), ("port_wide", wide)):
        ob._port = None
        ob.PORT_SO = lib
        t0 = time.time()
        r = ob.port_minimize(ob.CCSAQ, P.rosen_f, cons, [1e-8] * m, lb, ub, P.rosen_x0(n), maxeval=K + 1)
        out["runs"][name] = dict(f=float(r["minf"]), dual_evals=int(r["dual_evals"]),
                                 seconds=time.time() - t0)
        print(name, out["runs"][name], flush=True)
    fs = {k: v["f"] for k, v in out["runs"].items()}
    base = fs.get("port_wide")
    out["relative_to_port_wide"] = {k: (v - base) / abs(base) for k, v in fs.items()}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"r02_c3_noise_floor_K{K}_n{n}.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
