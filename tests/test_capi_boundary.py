"""The drop-in boundary without a GPU: libnlopt_b200.so loads, exports every symbol that
include/nlopt_b200.h declares, the object API behaves like the reference's (options.c), and
nlopt_optimize refuses to run on a machine without CUDA instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import nlopt_b200 as nl
from nlopt_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nlopt_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(nlopt_[a-z0-9_]+)\s*\(", src))
    typedefs = set(re.findall(r"\(\*\s*(nlopt_[a-z0-9_]+)\s*\)", src))
    return sorted(names - typedefs)


def test_every_declared_symbol_is_exported(built):
    exported = subprocess.run(["nm", "-D", "--defined-only", _capi.DEFAULT_LIB], capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in exported.splitlines() if l.strip())
    decl = declared_functions()
    assert len(decl) > 90
    missing = [f for f in decl if f not in exported]
    assert not missing, missing
    # and the python binder knows all of them
    known = set(_capi.STD_SYMBOLS) | set(_capi.EXT_SYMBOLS)
    assert not [f for f in decl if f not in known]


def test_reference_link_sets_are_covered(built):
    """`nm -u` of the reference's own test binaries for this path."""
    need = """nlopt_add_inequality_constraint nlopt_algorithm_name nlopt_create nlopt_destroy nlopt_get_algorithm
    nlopt_get_dimension nlopt_get_errmsg nlopt_get_param nlopt_nth_param nlopt_num_params nlopt_optimize
    nlopt_set_force_stop nlopt_set_initial_step1 nlopt_set_lower_bounds nlopt_set_lower_bounds1 nlopt_set_min_objective
    nlopt_set_munge nlopt_set_param nlopt_set_stopval nlopt_set_upper_bounds1 nlopt_set_xtol_rel
    nlopt_algorithm_from_string nlopt_set_maxeval nlopt_get_numevals nlopt_set_ftol_abs nlopt_set_ftol_rel
    nlopt_set_maxtime nlopt_set_upper_bounds nlopt_set_xtol_abs nlopt_srand nlopt_srand_time nlopt_version""".split()
    lib = C.CDLL(_capi.DEFAULT_LIB)
    for s in need:
        assert hasattr(lib, s), s


def test_names_and_enums_match_reference_abi(built):
    L = _capi.default_library()
    assert nl.LD_MMA == 24 and nl.LD_CCSAQ == 41 and nl.NUM_ALGORITHMS == 44
    assert L.nlopt_algorithm_name(24) == b"Method of Moving Asymptotes (MMA) (local, derivative)"
    assert L.nlopt_algorithm_name(41).startswith(b"CCSA (Conservative Convex Separable Approximations)")
    assert L.nlopt_algorithm_name(99) == b"UNKNOWN"
    assert L.nlopt_algorithm_from_string(b"LD_CCSAQ") == 41 and L.nlopt_algorithm_from_string(b"nope") == -1
    assert L.nlopt_algorithm_to_string(24) == b"LD_MMA"
    assert L.nlopt_result_to_string(4) == b"XTOL_REACHED" and L.nlopt_result_from_string(b"FORCED_STOP") == -5
    v = [C.c_int() for _ in range(3)]
    L.nlopt_version(*[C.byref(i) for i in v])
    assert (v[0].value, v[1].value) == (2, 11)


def test_create_rejects_bad_algorithm(built):
    L = _capi.default_library()
    assert not L.nlopt_create(-1, 3) and not L.nlopt_create(44, 3)
    h = L.nlopt_create(24, 0)
    assert h
    L.nlopt_destroy(h)


def test_object_api_semantics(built):
    o = nl.opt(nl.LD_MMA, 3)
    assert o.get_dimension() == 3 and o.get_algorithm() == nl.LD_MMA
    assert np.all(np.isneginf(o.get_lower_bounds())) and np.all(np.isposinf(o.get_upper_bounds()))
    assert o.get_stopval() == -np.inf and o.get_xtol_rel() == 0 and o.get_maxeval() == 0
    o.set_lower_bounds([0, 1, 2]); o.set_upper_bounds(5.0); o.set_upper_bound(1, 7.0)
    assert list(o.get_lower_bounds()) == [0, 1, 2] and list(o.get_upper_bounds()) == [5, 7, 5]
    with pytest.raises(ValueError):
        o.set_lower_bound(3, 0.0)
    assert "invalid bound index" in o.get_errmsg()
    # options.c:375-377: a subnormally thin interval is snapped shut
    o.set_upper_bound(0, 5e-324)
    o.set_lower_bound(0, 0.0)
    assert o.get_lower_bounds()[0] == o.get_upper_bounds()[0]
    # named parameters
    o.set_param("inner_maxeval", 123)
    assert o.get_param("inner_maxeval", 1234) == 123 and o.get_param("not a param", 1234) == 1234
    assert o.num_params() == 1 and o.nth_param(0) == "inner_maxeval" and o.has_param("inner_maxeval")
    o.set_param("inner_maxeval", 5)
    assert o.num_params() == 1 and o.get_param("inner_maxeval", 0) == 5
    # tolerances / weights
    assert list(o.get_xtol_abs()) == [0, 0, 0] and list(o.get_x_weights()) == [1, 1, 1]
    o.set_xtol_abs(1e-3); o.set_x_weights([1, 2, 3])
    assert list(o.get_xtol_abs()) == [1e-3] * 3 and list(o.get_x_weights()) == [1, 2, 3]
    with pytest.raises(ValueError):
        o.set_x_weights([1, -2, 3])
    with pytest.raises(ValueError):
        o.set_initial_step(0.0)
    o.set_initial_step([0.1, 0.2, 0.3])
    assert list(o.get_initial_step([1, 1, 1])) == [0.1, 0.2, 0.3]
    # constraints: negative tolerance, equality constraints are not an MMA/CCSAQ feature (options.c:617-622)
    f = lambda x, g: 0.0
    with pytest.raises(ValueError):
        o.add_inequality_constraint(f, -1.0)
    with pytest.raises(ValueError):
        o.add_equality_constraint(f, 0.0)
    assert "invalid algorithm for constraints" in o.get_errmsg()
    o.add_inequality_constraint(f, 1e-8)
    o.add_inequality_mconstraint(lambda r, x, g: None, [1e-8, 1e-8])
    o.remove_inequality_constraints()
    o.force_stop(); assert o.get_force_stop() == 1
    o.set_force_stop(0); assert o.get_force_stop() == 0
    # inequality constraints are refused for algorithms that cannot take them
    b = nl.opt(nl.LN_BOBYQA, 2)
    with pytest.raises(ValueError):
        b.add_inequality_constraint(f, 0.0)


def test_copy_and_munge(built):
    L = _capi.default_library()
    h = L.nlopt_create(41, 2)
    L.nlopt_set_param(h, b"rho_init", 0.5)
    L.nlopt_set_xtol_rel(h, 1e-4)
    c = L.nlopt_copy(h)
    assert L.nlopt_get_param(c, b"rho_init", 1.0) == 0.5 and L.nlopt_get_xtol_rel(c) == 1e-4
    L.nlopt_set_param(c, b"rho_init", 2.0)
    assert L.nlopt_get_param(h, b"rho_init", 1.0) == 0.5
    # munge_on_destroy is called for f_data on destroy and on objective replacement (options.c:36-47, :326)
    calls = []
    MUNGE = C.CFUNCTYPE(C.c_void_p, C.c_void_p)
    cb = MUNGE(lambda p: calls.append(p) or None)
    L.nlopt_set_munge(c, C.cast(cb, C.c_void_p), None)
    f = _capi.NLOPT_FUNC(lambda n, x, g, d: 0.0)
    L.nlopt_set_min_objective(c, f, 1234)
    L.nlopt_set_min_objective(c, f, 5678)
    assert calls[-1] == 1234
    L.nlopt_destroy(c)
    assert calls[-1] == 5678
    L.nlopt_destroy(h)


def test_only_mma_and_ccsaq_are_executable(built):
    o = nl.opt(nl.LD_SLSQP, 2)
    o.set_min_objective(lambda x, g: 0.0)
    with pytest.raises(ValueError):
        o.optimize([0.0, 0.0])
    assert "not part of this library" in o.get_errmsg()


def test_parameter_validation_matches_reference(built):
    for k, v, msg in (("rho_init", -1.0, "rho_init"), ("inner_gradients", 2, "inner_gradients"),
                      ("always_improve", 3, "always_improve"), ("sigma_min", -1.0, "sigma_min"),
                      ("dual_algorithm", float(nl.LD_SLSQP), "dual_algorithm")):
        o = nl.opt(nl.LD_MMA, 2)
        o.set_min_objective(lambda x, g: 0.0)
        o.set_param(k, v)
        with pytest.raises(ValueError):
            o.optimize([0.0, 0.0])
        assert msg in o.get_errmsg()
    o = nl.opt(nl.LD_MMA, 2)
    o.set_min_objective(lambda x, g: 0.0)
    o.set_lower_bounds(1.0)
    with pytest.raises(ValueError):
        o.optimize([0.0, 0.0])             # x0 outside the box, optimize.c:547-551
    assert "bounds 0 fail" in o.get_errmsg()


def test_n_equal_zero_shortcut(built):
    o = nl.opt(nl.LD_MMA, 0)               # optimize.c:536-539
    o.set_min_objective(lambda x, g: 42.0)
    o.optimize([])
    assert o.last_optimum_value() == 42.0 and o.last_optimize_result() == nl.SUCCESS


@pytest.mark.skipif(nl.device_count() > 0, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    o = nl.opt(nl.LD_CCSAQ, 2)
    o.set_min_objective(lambda x, g: float(x[0] ** 2))
    with pytest.raises(RuntimeError) as e:
        o.optimize([1.0, 1.0])
    assert "CUDA" in str(e.value)
    assert o.last_optimize_result() == nl.FAILURE
    L = _capi.default_library()
    assert not L.nlopt_b200_dual_create(0, 10, 1)


def test_product_library_does_not_link_the_oracle(built):
    out = subprocess.run(["ldd", _capi.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "nlopt_ref" not in out
    syms = subprocess.run(["nm", "-D", _capi.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "port_dual" not in syms and "ref_mma" not in syms


def test_shard_geometry(built):
    L = _capi.default_library()
    for n in (1, 2, 7, 1000, 12345, 10**7, 5 * 10**7 + 1):
        for world in (1, 2, 4, 8):
            end = 0
            for r in range(world):
                j0, cnt = C.c_ulonglong(), C.c_ulonglong()
                L.nlopt_b200_shard_range(n, r, world, C.byref(j0), C.byref(cnt))
                assert j0.value == end or cnt.value == 0
                assert j0.value % 2 == 0
                end = max(end, j0.value + cnt.value)
            assert end == n
