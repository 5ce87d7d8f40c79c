"""NLOPT_AUGLAG / AUGLAG_EQ / LD_AUGLAG / LD_AUGLAG_EQ (the callers that default to MMA).

Host logic (run_auglag in nlopt_api.cpp: rho / lambda / mu updates, stopping, nlopt_optimize_limited, the
sub-optimiser set-up) is checked here without a GPU by linking it with the CPU test backend and comparing with
the UNMODIFIED reference library on the same callbacks.  The device side -- the augmented-Lagrangian objective with
its penalty_axpy_kernel -- is checked in the gpu-marked tests against the same reference runs."""
import numpy as np
import pytest

import nlopt_b200 as nl
import problems as P


def circle_eq(x, grad):            # h(x) = x0^2 + x1^2 + ... - 1 = 0
    if grad.size:
        grad[:] = 2 * x
    return float(np.dot(x, x) - 1.0)


def plane_eq(x, grad):             # h(x) = sum x - 0.3 = 0
    if grad.size:
        grad[:] = 1.0
    return float(np.sum(x) - 0.3)


def lin_obj(x, grad):
    n = x.size
    w = 1.0 + 0.5 * np.sin(0.37 * np.arange(n))
    if grad.size:
        grad[:] = w
    return float(np.dot(w, x))


def halfspace(x, grad):            # c(x) = 0.2 - x0 <= 0
    if grad.size:
        grad[:] = 0.0
        grad[0] = -1.0
    return float(0.2 - x[0])


def vec_ineq(result, x, grad):     # c_k(x) = x_k - 0.5 <= 0, k = 0, 1
    for k in range(2):
        result[k] = x[k] - 0.5
    if grad.size:
        grad[:] = 0.0
        grad[0, 0] = 1.0
        grad[1, 1] = 1.0


def _run(lib, alg, n, f, ineq, eq, lb, ub, x0, local=None, mineq=None, **kw):
    o = nl.opt(alg, n, library=lib)
    o.set_lower_bounds(lb); o.set_upper_bounds(ub)
    o.set_min_objective(f)
    for c, t in ineq:
        o.add_inequality_constraint(c, t)
    if mineq is not None:
        o.add_inequality_mconstraint(mineq[0], mineq[1])
    for h, t in eq:
        o.add_equality_constraint(h, t)
    if local is not None:
        lo = nl.opt(local[0], n, library=lib)
        for k, v in local[1].items():
            getattr(lo, "set_" + k)(v)
        o.set_local_optimizer(lo)
    for k, v in kw.items():
        getattr(o, "set_" + k)(v)
    x = o.optimize(np.array(x0, dtype=float))
    return dict(ret=o.last_optimize_result(), x=x, minf=o.last_optimum_value(), numevals=o.get_numevals(), opt=o)


def _same(a, b, ftol=1e-9, xtol=1e-7, evals_slack=0):
    assert a["ret"] == b["ret"], (a["ret"], b["ret"], a["opt"].get_errmsg())
    assert abs(a["numevals"] - b["numevals"]) <= evals_slack, (a["numevals"], b["numevals"])
    assert abs(a["minf"] - b["minf"]) <= ftol * max(1.0, abs(b["minf"]))
    assert np.max(np.abs(a["x"] - b["x"])) <= xtol


def _cases():
    n = 40
    out = []
    out.append(("tutorial-LD_AUGLAG", dict(alg=nl.LD_AUGLAG, n=2, f=P.tut_f, ineq=[(P.tut_c(2, 0), 1e-8), (P.tut_c(-1, 1), 1e-8)],
                                            eq=[], lb=[-np.inf, 0.05], ub=[np.inf, np.inf], x0=P.TUT_X0, kw=dict(xtol_rel=1e-4, maxeval=5000))))
    out.append(("tutorial-LD_AUGLAG_EQ", dict(alg=nl.LD_AUGLAG_EQ, n=2, f=P.tut_f, ineq=[(P.tut_c(2, 0), 1e-8), (P.tut_c(-1, 1), 1e-8)],
                                               eq=[], lb=[-np.inf, 0.05], ub=[np.inf, np.inf], x0=P.TUT_X0, kw=dict(xtol_rel=1e-4, maxeval=5000))))
    out.append(("sphere-eq-LD_AUGLAG", dict(alg=nl.LD_AUGLAG, n=n, f=lin_obj, ineq=[(halfspace, 1e-8)], eq=[(circle_eq, 1e-8)],
                                             lb=np.full(n, -2.0), ub=np.full(n, 2.0), x0=np.full(n, 0.3), kw=dict(xtol_rel=1e-6, maxeval=3000))))
    out.append(("sphere-eq-LD_AUGLAG_EQ", dict(alg=nl.LD_AUGLAG_EQ, n=n, f=lin_obj, ineq=[(halfspace, 1e-8)], eq=[(circle_eq, 1e-8)],
                                                lb=np.full(n, -2.0), ub=np.full(n, 2.0), x0=np.full(n, 0.3), kw=dict(xtol_rel=1e-6, maxeval=3000))))
    out.append(("two-eq-AUGLAG-over-CCSAQ", dict(alg=nl.AUGLAG, n=n, f=lin_obj, ineq=[], eq=[(circle_eq, 1e-8), (plane_eq, 1e-8)],
                                                  lb=np.full(n, -2.0), ub=np.full(n, 2.0), x0=np.full(n, 0.3),
                                                  local=(nl.LD_CCSAQ, dict(xtol_rel=1e-7, maxeval=400)), kw=dict(xtol_rel=1e-6, maxeval=4000))))
    out.append(("vector-ineq-AUGLAG_EQ-over-MMA", dict(alg=nl.AUGLAG_EQ, n=n, f=lin_obj, ineq=[], mineq=(vec_ineq, [1e-8, 1e-8]), eq=[(circle_eq, 1e-8)],
                                                        lb=np.full(n, -2.0), ub=np.full(n, 2.0), x0=np.full(n, 0.3),
                                                        local=(nl.LD_MMA, dict(xtol_rel=1e-7, maxeval=400)), kw=dict(xtol_rel=1e-6, maxeval=4000))))
    out.append(("maxeval-stop", dict(alg=nl.LD_AUGLAG, n=n, f=lin_obj, ineq=[(halfspace, 1e-8)], eq=[(circle_eq, 1e-8)],
                                     lb=np.full(n, -2.0), ub=np.full(n, 2.0), x0=np.full(n, 0.3), kw=dict(xtol_rel=1e-12, maxeval=57))))
    return out


def _sub_has_constraints(c):
    return c["alg"] in (nl.LD_AUGLAG_EQ, nl.AUGLAG_EQ) and (len(c["ineq"]) > 0 or c.get("mineq") is not None)


def _call(lib, c):
    return _run(lib, c["alg"], c["n"], c["f"], c["ineq"], c["eq"], c["lb"], c["ub"], c["x0"], local=c.get("local"),
                mineq=c.get("mineq"), **c["kw"])


@pytest.mark.parametrize("name,case", _cases(), ids=[n for n, _ in _cases()])
def test_auglag_host_logic_matches_reference(hosttest_lib, reflib, name, case):
    a, b = _call(hosttest_lib, case), _call(reflib, case)
    assert b["ret"] > 0
    if _sub_has_constraints(case):
        # the sub-problems are constrained MMA runs: rounding-level differences of the dual sums are amplified by
        # the flat dual optimum, so late iterates -- and which stopping test fires -- may differ
        assert a["ret"] > 0
        assert abs(a["minf"] - b["minf"]) <= 1e-5 * max(1.0, abs(b["minf"])) and np.max(np.abs(a["x"] - b["x"])) <= 2e-3
    elif case.get("local", (None,))[0] == nl.LD_CCSAQ:
        _same(a, b, ftol=1e-7, xtol=1e-6, evals_slack=b["numevals"] // 8)     # CCSAQ sub-problems: rounding-level drift
    else:
        _same(a, b)                  # penalty-only MMA sub-problems: same trajectory, same counts


def test_auglag_argument_checks(hosttest_lib):
    o = nl.opt(nl.AUGLAG, 2, library=hosttest_lib)
    o.set_min_objective(P.tut_f)
    o.set_lower_bounds([-1.0, 0.0]); o.set_upper_bounds([3.0, 10.0])
    with pytest.raises(Exception):
        o.optimize(np.array(P.TUT_X0))                       # no local optimiser: INVALID_ARGS like the reference
    assert "local optimizer" in o.get_errmsg()
    o2 = nl.opt(nl.LN_AUGLAG, 2, library=hosttest_lib)
    o2.set_min_objective(P.tut_f)
    o2.set_lower_bounds([-1.0, 0.0]); o2.set_upper_bounds([3.0, 10.0])
    with pytest.raises(Exception):
        o2.optimize(np.array(P.TUT_X0))                      # derivative-free default is not part of this library
    assert "LD_MMA" in o2.get_errmsg()


@pytest.mark.gpu
@pytest.mark.parametrize("name,case", _cases(), ids=[n for n, _ in _cases()])
def test_auglag_on_gpu_matches_reference(built, reflib, name, case):
    """Same runs through the product library: the augmented-Lagrangian objective and its gradient are assembled on
    the device (penalty rows uploaded, penalty_axpy_kernel); MMA/CCSAQ run on the GPU."""
    a, b = _call(None, case), _call(reflib, case)
    assert a["ret"] > 0 and b["ret"] > 0
    penalty_only_mma = not _sub_has_constraints(case) and case.get("local", (None,))[0] != nl.LD_CCSAQ
    if case["kw"].get("maxeval") == 57:          # short fixed-length run: same count, same point to rounding
        assert a["ret"] == b["ret"] and a["numevals"] == b["numevals"]
        assert abs(a["minf"] - b["minf"]) <= 1e-7 * max(1.0, abs(b["minf"])) and np.max(np.abs(a["x"] - b["x"])) <= 1e-5
    elif penalty_only_mma:
        assert abs(a["minf"] - b["minf"]) <= 1e-5 * max(1.0, abs(b["minf"])) and np.max(np.abs(a["x"] - b["x"])) <= 2e-3
    else:
        # constrained or CCSAQ sub-problems: the reference itself may stop on MAXEVAL near the optimum and the late
        # iterates are rounding-sensitive; require a feasible point with the same objective to 1e-3
        none = np.empty(0)
        for h, _ in case["eq"]:
            assert abs(h(a["x"], none)) <= 1e-4
        for c, _ in case["ineq"]:
            assert c(a["x"], none) <= 1e-5
        assert abs(a["minf"] - b["minf"]) <= 1e-3 * max(1.0, abs(b["minf"]))


@pytest.mark.gpu
def test_auglag_large_n_gradient_assembly(built, reflib):
    """n large enough to exercise the sharded upload + axpy path; short fixed run against the reference."""
    n = 20011
    c = dict(alg=nl.LD_AUGLAG, n=n, f=lin_obj, ineq=[(halfspace, 1e-8)], eq=[(circle_eq, 1e-8), (plane_eq, 1e-8)],
             lb=np.full(n, -2.0), ub=np.full(n, 2.0), x0=np.full(n, 0.01), kw=dict(xtol_rel=1e-10, maxeval=40))
    a, b = _call(None, c), _call(reflib, c)
    assert a["ret"] == b["ret"] and a["numevals"] == b["numevals"]
    assert abs(a["minf"] - b["minf"]) <= 1e-7 * max(1.0, abs(b["minf"]))
    assert np.max(np.abs(a["x"] - b["x"])) <= 1e-6


def test_auglag_maximize_stopval_and_forced_stop(hosttest_lib, reflib):
    """Option surface around the outer loop: maximisation (sign flip, optimize.c:1014-1024), stopval on a feasible
    point (auglag.c:272), nlopt_force_stop from inside a callback (auglag.c:39, :218) -- same outcome as the reference."""
    n = 12
    lb, ub, x0 = np.full(n, -2.0), np.full(n, 2.0), np.full(n, 0.3)

    def neg_obj(x, grad):
        v = lin_obj(x, grad)
        if grad.size:
            grad[:] = -grad
        return -v

    res = {}
    for name, lib in (("ours", hosttest_lib), ("ref", reflib)):
        o = nl.opt(nl.LD_AUGLAG, n, library=lib)
        o.set_lower_bounds(lb); o.set_upper_bounds(ub)
        o.set_max_objective(neg_obj)
        o.add_equality_constraint(circle_eq, 1e-8)
        o.set_xtol_rel(1e-6); o.set_maxeval(2000)
        x = o.optimize(x0.copy())
        res[name] = (o.last_optimize_result(), o.get_numevals(), o.last_optimum_value(), x)
    assert res["ours"][0] == res["ref"][0] and res["ours"][1] == res["ref"][1]
    assert abs(res["ours"][2] - res["ref"][2]) <= 1e-9 and np.max(np.abs(res["ours"][3] - res["ref"][3])) <= 1e-7

    for name, lib in (("ours", hosttest_lib), ("ref", reflib)):
        o = nl.opt(nl.LD_AUGLAG, n, library=lib)
        o.set_lower_bounds(lb); o.set_upper_bounds(ub)
        o.set_min_objective(lin_obj)
        o.add_inequality_constraint(halfspace, 1e-8)
        o.set_stopval(1.0); o.set_xtol_rel(1e-9); o.set_maxeval(2000)      # any feasible point with f < 1 ends the run
        x = o.optimize(x0.copy())
        res[name] = (o.last_optimize_result(), o.get_numevals(), o.last_optimum_value(), x)
    assert res["ours"][0] == res["ref"][0] == nl.STOPVAL_REACHED and res["ours"][1] == res["ref"][1]
    assert abs(res["ours"][2] - res["ref"][2]) <= 1e-9

    for name, lib in (("ours", hosttest_lib), ("ref", reflib)):
        o = nl.opt(nl.LD_AUGLAG, n, library=lib)
        calls = [0]

        def stopping_obj(x, grad, o=o, calls=calls):
            calls[0] += 1
            if calls[0] == 9:
                o.force_stop()
            return lin_obj(x, grad)

        o.set_lower_bounds(lb); o.set_upper_bounds(ub)
        o.set_min_objective(stopping_obj)
        o.add_equality_constraint(circle_eq, 1e-8)
        o.set_xtol_rel(1e-9); o.set_maxeval(500)
        try:
            o.optimize(x0.copy())
        except Exception:
            pass
        res[name] = (o.last_optimize_result(), calls[0])
    assert res["ours"] == res["ref"] and res["ours"][0] == nl.FORCED_STOP
