"""Generate tests/golden/reference_golden.json by running the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/Makefile) in this container.  Commit the JSON; the GPU box has no
/root/reference and compares against these vectors.

    python tests/golden/make_golden.py

Contents:
  dual[]   : the reference's static dual_func (mma.c:59-137 / ccsa_quadratic.c:79-148) on the synthetic
             instance of tests/synth.py: return value, gval, wval, gcval[] as C99 hex floats, and for x*(y)
             the xor of the 64-bit patterns, the plain sum and 8 sampled entries.
  solve[]  : whole nlopt_optimize runs of the reference library: result code, evaluations, f* (hex), x* hash.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, ROOT)
sys.path.insert(0, TESTS)
import numpy as np  # noqa: E402

import nlopt_b200 as nl  # noqa: E402
import oracle_bindings as ob  # noqa: E402
import problems as P  # noqa: E402
import synth  # noqa: E402
from nlopt_b200 import Library  # noqa: E402

DUAL_CASES = [(v, n, m) for v in (ob.MMA, ob.CCSAQ) for (n, m) in
              ((1, 0), (5, 2), (1000, 1), (4097, 4), (100001, 1), (100000, 4), (60000, 16), (9999, 20), (250000, 8))]


def xhash(x):
    return int(np.bitwise_xor.reduce(np.ascontiguousarray(x).view(np.uint64))) if x.size else 0


def main():
    assert ob.ref_dual_available(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build_oracle()')"
    ref = Library(ob.REF_SO, extensions=False)
    out = {"generator": "tests/golden/make_golden.py", "reference": "NLopt 2.11.0 (oracle/_ref, gcc -O3 -ffp-contract=off)",
           "dual": [], "solve": []}
    for variant, n, m in DUAL_CASES:
        inst = synth.kernel_instance(n, m)
        r = ob.ref_dual(variant, inst)
        idx = np.linspace(0, n - 1, 8).astype(int)
        out["dual"].append(dict(variant=variant, n=n, m=m, seed=synth.SEED0, ret=float(r["ret"]).hex(), g0=float(r["g0"]).hex(),
                                w=float(r["w"]).hex(), gc=[float(v).hex() for v in r["gc"]], x_xor=xhash(r["xcur"]),
                                x_sum=float(np.sum(r["xcur"])).hex(), x_idx=idx.tolist(),
                                x_samples=[float(r["xcur"][i]).hex() for i in idx]))

    def run(alg, n, f, cons, tols, lb, ub, x0, **kw):
        o = nl.opt(alg, n, library=ref)
        o.set_lower_bounds(lb); o.set_upper_bounds(ub); o.set_min_objective(f)
        for c, t in zip(cons, tols):
            o.add_inequality_constraint(c, t)
        for k, v in kw.items():
            if k in ("xtol_rel", "maxeval", "stopval"):
                getattr(o, "set_" + k)(v)
            else:
                o.set_param(k, v)
        x = o.optimize(x0)
        return dict(ret=o.last_optimize_result(), numevals=o.get_numevals(), minf=float(o.last_optimum_value()).hex(),
                    x_xor=xhash(x), x_sum=float(np.sum(x)).hex(), x_first=[float(v).hex() for v in x[:4]])

    for name, alg in (("LD_MMA", nl.LD_MMA), ("LD_CCSAQ", nl.LD_CCSAQ)):
        n, m = 20000, 4
        cons = [P.lin_constraint(k, n) for k in range(m)]
        g = run(alg, n, P.rosen_f, cons, [1e-8] * m, np.full(n, -2.0), np.full(n, 2.0), P.rosen_x0(n), maxeval=30)
        out["solve"].append(dict(problem="rosenbrock+4lin", alg=name, n=n, m=m, maxeval=30, **g))
        n = 100000
        f, c = P.quad_problem(n)
        g = run(alg, n, f, [c], [0.0], np.full(n, -1.0), np.full(n, 1.0), np.full(n, -0.5), xtol_rel=1e-6, maxeval=300)
        out["solve"].append(dict(problem="quadratic+mean", alg=name, n=n, m=1, xtol_rel=1e-6, **g))
    n = 100000
    f, c = P.simp_problem(n)
    g = run(nl.LD_MMA, n, f, [c], [0.0], np.zeros(n), np.ones(n), np.full(n, 0.4), xtol_rel=1e-6, maxeval=300)
    out["solve"].append(dict(problem="simp+volume", alg="LD_MMA", n=n, m=1, xtol_rel=1e-6, **g))
    with open(os.path.join(HERE, "reference_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["dual"]), "dual and", len(out["solve"]), "solve goldens")


if __name__ == "__main__":
    main()
