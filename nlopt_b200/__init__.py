"""nlopt_b200 -- Python face of the B200-native MMA/CCSAQ solver.

Mirrors the reference's Python module (SWIG over nlopt.hpp: src/swig/nlopt-python.i,
src/api/nlopt-in.hpp:254-609): ``opt(algorithm, n)``, ``set_min_objective(f)`` with
``f(x, grad)`` writing ``grad`` in place when ``grad.size > 0``, ``add_inequality_constraint``,
``optimize(x)`` returning the optimum as an array, ``last_optimum_value()``,
``last_optimize_result()``, the ``LD_MMA`` / ``LD_CCSAQ`` / result-code constants and the
exception mapping of nlopt.hpp ``mythrow`` (:87-103).

All arithmetic happens in ``libnlopt_b200.so`` (CUDA, sm_100a) behind the NLopt C ABI;
this module only marshals numpy arrays.  ``opt(..., library=Library(path))`` points the same
class at another library exporting that ABI (tests use it to run the unmodified reference).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import (NLOPT_B200_DFUNC, NLOPT_FUNC, NLOPT_PRECOND, NLOPT_MFUNC, Library, Stats, c_double_p,
                    default_library)

# --- algorithm ids (reference nlopt.h:72-154) ---------------------------------
_ALG_NAMES = [
    "GN_DIRECT", "GN_DIRECT_L", "GN_DIRECT_L_RAND", "GN_DIRECT_NOSCAL", "GN_DIRECT_L_NOSCAL",
    "GN_DIRECT_L_RAND_NOSCAL", "GN_ORIG_DIRECT", "GN_ORIG_DIRECT_L", "GD_STOGO", "GD_STOGO_RAND",
    "LD_LBFGS_NOCEDAL", "LD_LBFGS", "LN_PRAXIS", "LD_VAR1", "LD_VAR2", "LD_TNEWTON",
    "LD_TNEWTON_RESTART", "LD_TNEWTON_PRECOND", "LD_TNEWTON_PRECOND_RESTART", "GN_CRS2_LM",
    "GN_MLSL", "GD_MLSL", "GN_MLSL_LDS", "GD_MLSL_LDS", "LD_MMA", "LN_COBYLA", "LN_NEWUOA",
    "LN_NEWUOA_BOUND", "LN_NELDERMEAD", "LN_SBPLX", "LN_AUGLAG", "LD_AUGLAG", "LN_AUGLAG_EQ",
    "LD_AUGLAG_EQ", "LN_BOBYQA", "GN_ISRES", "AUGLAG", "AUGLAG_EQ", "G_MLSL", "G_MLSL_LDS",
    "LD_SLSQP", "LD_CCSAQ", "GN_ESCH", "GN_AGS",
]
for _i, _n in enumerate(_ALG_NAMES):
    globals()[_n] = _i
NUM_ALGORITHMS = len(_ALG_NAMES)

# --- result codes (reference nlopt.h:162-176) ---------------------------------
FAILURE, INVALID_ARGS, OUT_OF_MEMORY, ROUNDOFF_LIMITED, FORCED_STOP = -1, -2, -3, -4, -5
SUCCESS, STOPVAL_REACHED, FTOL_REACHED, XTOL_REACHED, MAXEVAL_REACHED, MAXTIME_REACHED = 1, 2, 3, 4, 5, 6


class RoundoffLimited(RuntimeError):
    """nlopt.hpp roundoff_limited (:71-74)"""


class ForcedStop(RuntimeError):
    """nlopt.hpp forced_stop (:76-79)"""


def _as_f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"dimension mismatch: expected {n}, got {a.size}")
    return a


def _ptr(a):
    return a.ctypes.data_as(c_double_p)


class opt:
    """One optimisation problem; thin owner of an ``nlopt_opt`` handle."""

    def __init__(self, algorithm, n, library: Library | None = None):
        self._lib = library or default_library()
        self._h = self._lib.nlopt_create(int(algorithm), int(n))
        if not self._h:
            raise RuntimeError("nlopt failure")       # nlopt.hpp:263
        self._n = int(n)
        self._keep = []            # ctypes thunks must outlive the handle
        self._exc = None           # exception raised inside a callback
        self._last_result = FAILURE
        self._last_optf = float("inf")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.nlopt_destroy(h)

    # ---- error mapping, nlopt.hpp:87-103 -------------------------------------
    def _check(self, ret):
        if ret >= 0:
            return ret
        msg = self._lib.nlopt_get_errmsg(self._h)
        msg = msg.decode() if msg else None
        if ret == FAILURE:
            raise RuntimeError(msg or "nlopt failure")
        if ret == OUT_OF_MEMORY:
            raise MemoryError(msg or "out of memory")
        if ret == INVALID_ARGS:
            raise ValueError(msg or "nlopt invalid argument")
        if ret == ROUNDOFF_LIMITED:
            raise RoundoffLimited(msg or "nlopt roundoff-limited")
        if ret == FORCED_STOP:
            raise ForcedStop(msg or "nlopt forced stop")
        raise RuntimeError(msg or "nlopt failure")

    # ---- callbacks -----------------------------------------------------------
    def _wrap_func(self, f):
        n_fixed = self._n

        def thunk(n, x, grad, _data):
            try:
                xa = np.ctypeslib.as_array(x, shape=(n,))
                ga = np.ctypeslib.as_array(grad, shape=(n,)) if grad else np.empty(0)
                return float(f(xa, ga))
            except BaseException as e:      # nlopt.hpp:149-166: exception => forced stop
                self._exc = e
                self._lib.nlopt_force_stop(self._h)
                return float("nan")

        cb = NLOPT_FUNC(thunk)
        self._keep.append(cb)
        del n_fixed
        return cb

    def _wrap_mfunc(self, f):
        def thunk(m, result, n, x, grad, _data):
            try:
                ra = np.ctypeslib.as_array(result, shape=(m,))
                xa = np.ctypeslib.as_array(x, shape=(n,))
                ga = np.ctypeslib.as_array(grad, shape=(m, n)) if grad else np.empty(0)
                f(ra, xa, ga)
            except BaseException as e:
                self._exc = e
                self._lib.nlopt_force_stop(self._h)

        cb = NLOPT_MFUNC(thunk)
        self._keep.append(cb)
        return cb

    def set_min_objective(self, f):
        self._check(self._lib.nlopt_set_min_objective(self._h, self._wrap_func(f), None))

    def set_max_objective(self, f):
        self._check(self._lib.nlopt_set_max_objective(self._h, self._wrap_func(f), None))

    # preconditioned forms (nlopt.h:70, options.c:322-337): pre(x, v, vpre) writes vpre = H(x) v
    def _wrap_precond(self, pre):
        def thunk(n, x, v, vpre, _data):
            try:
                pre(np.ctypeslib.as_array(x, shape=(n,)), np.ctypeslib.as_array(v, shape=(n,)), np.ctypeslib.as_array(vpre, shape=(n,)))
            except BaseException as e:
                self._exc = e
                self._lib.nlopt_force_stop(self._h)

        cb = NLOPT_PRECOND(thunk)
        self._keep.append(cb)
        return C.cast(cb, C.c_void_p)

    def set_precond_min_objective(self, f, pre):
        self._check(self._lib.nlopt_set_precond_min_objective(self._h, self._wrap_func(f), self._wrap_precond(pre), None))

    def add_precond_inequality_constraint(self, fc, pre, tol=0.0):
        self._check(self._lib.nlopt_add_precond_inequality_constraint(self._h, self._wrap_func(fc), self._wrap_precond(pre), None, float(tol)))

    def add_inequality_constraint(self, fc, tol=0.0):
        self._check(self._lib.nlopt_add_inequality_constraint(self._h, self._wrap_func(fc), None, float(tol)))

    def add_equality_constraint(self, h, tol=0.0):
        self._check(self._lib.nlopt_add_equality_constraint(self._h, self._wrap_func(h), None, float(tol)))

    def add_inequality_mconstraint(self, fc, tol):
        tol = _as_f64(tol)
        self._check(self._lib.nlopt_add_inequality_mconstraint(
            self._h, tol.size, self._wrap_mfunc(fc), None, _ptr(tol)))

    def add_equality_mconstraint(self, h, tol):
        tol = _as_f64(tol)
        self._check(self._lib.nlopt_add_equality_mconstraint(
            self._h, tol.size, self._wrap_mfunc(h), None, _ptr(tol)))

    def remove_inequality_constraints(self):
        self._check(self._lib.nlopt_remove_inequality_constraints(self._h))

    def remove_equality_constraints(self):
        self._check(self._lib.nlopt_remove_equality_constraints(self._h))

    # ---- extension: device-resident callbacks (raw C function pointers) -------
    def set_min_objective_device(self, fn_ptr, data_ptr=None):
        self._check(self._lib.nlopt_b200_set_min_objective_device(self._h, fn_ptr, data_ptr))

    def add_inequality_constraint_device(self, fn_ptr, data_ptr=None, tol=0.0):
        self._check(self._lib.nlopt_b200_add_inequality_constraint_device(
            self._h, fn_ptr, data_ptr, float(tol)))

    def optimize_device(self, x_dev_ptr):
        f = C.c_double(0.0)
        self._exc = None
        ret = self._lib.nlopt_b200_optimize_device(self._h, x_dev_ptr, C.byref(f))
        self._last_result, self._last_optf = ret, f.value
        self._check(ret)
        return ret

    def get_stats(self):
        s = Stats()
        self._check(self._lib.nlopt_b200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    # ---- run (nlopt.hpp:299-321) ---------------------------------------------
    def optimize(self, x):
        xa = np.array(x, dtype=np.float64, copy=True).reshape(-1)
        if xa.size != self._n:
            raise ValueError("dimension mismatch")
        f = C.c_double(0.0)
        self._exc = None
        ret = self._lib.nlopt_optimize(self._h, _ptr(xa), C.byref(f))
        self._last_result, self._last_optf = ret, f.value
        if ret == FORCED_STOP and self._exc is not None:
            e, self._exc = self._exc, None
            raise e
        self._check(ret)
        return xa

    def optimize_inplace(self, xa):
        """The C call itself (nlopt.h: nlopt_optimize(opt, x, &minf)): `xa` is the caller's own contiguous float64 buffer,
        start point on entry, solution on return -- no Python-side copy of the n doubles.  Returns the nlopt_result."""
        if not (isinstance(xa, np.ndarray) and xa.dtype == np.float64 and xa.flags["C_CONTIGUOUS"] and xa.size == self._n):
            raise ValueError("optimize_inplace needs a contiguous float64 array of the problem's dimension")
        f = C.c_double(0.0)
        self._exc = None
        ret = self._lib.nlopt_optimize(self._h, _ptr(xa), C.byref(f))
        self._last_result, self._last_optf = ret, f.value
        if ret == FORCED_STOP and self._exc is not None:
            e, self._exc = self._exc, None
            raise e
        self._check(ret)
        return ret

    def last_optimize_result(self):
        return self._last_result

    def last_optimum_value(self):
        return self._last_optf

    # ---- accessors -----------------------------------------------------------
    def get_algorithm(self):
        return self._lib.nlopt_get_algorithm(self._h)

    def get_algorithm_name(self):
        return self._lib.nlopt_algorithm_name(self.get_algorithm()).decode()

    def get_dimension(self):
        return self._lib.nlopt_get_dimension(self._h)

    def get_errmsg(self):
        m = self._lib.nlopt_get_errmsg(self._h)
        return m.decode() if m else None

    def get_numevals(self):
        return self._lib.nlopt_get_numevals(self._h)

    def set_param(self, name, val):
        self._check(self._lib.nlopt_set_param(self._h, name.encode(), float(val)))

    def get_param(self, name, default):
        return self._lib.nlopt_get_param(self._h, name.encode(), float(default))

    def has_param(self, name):
        return bool(self._lib.nlopt_has_param(self._h, name.encode()))

    def num_params(self):
        return self._lib.nlopt_num_params(self._h)

    def nth_param(self, i):
        s = self._lib.nlopt_nth_param(self._h, int(i))
        return s.decode() if s else None

    def _set_vec_or_scalar(self, vec_fn, scalar_fn, v):
        if np.isscalar(v):
            self._check(scalar_fn(self._h, float(v)))
        else:
            a = _as_f64(v, self._n)
            self._check(vec_fn(self._h, _ptr(a)))

    def _get_vec(self, fn):
        a = np.empty(self._n)
        self._check(fn(self._h, _ptr(a)))
        return a

    def set_lower_bounds(self, v):
        self._set_vec_or_scalar(self._lib.nlopt_set_lower_bounds, self._lib.nlopt_set_lower_bounds1, v)

    def set_upper_bounds(self, v):
        self._set_vec_or_scalar(self._lib.nlopt_set_upper_bounds, self._lib.nlopt_set_upper_bounds1, v)

    def set_lower_bound(self, i, v):
        self._check(self._lib.nlopt_set_lower_bound(self._h, int(i), float(v)))

    def set_upper_bound(self, i, v):
        self._check(self._lib.nlopt_set_upper_bound(self._h, int(i), float(v)))

    def get_lower_bounds(self):
        return self._get_vec(self._lib.nlopt_get_lower_bounds)

    def get_upper_bounds(self):
        return self._get_vec(self._lib.nlopt_get_upper_bounds)

    def set_xtol_abs(self, v):
        self._set_vec_or_scalar(self._lib.nlopt_set_xtol_abs, self._lib.nlopt_set_xtol_abs1, v)

    def get_xtol_abs(self):
        return self._get_vec(self._lib.nlopt_get_xtol_abs)

    def set_x_weights(self, v):
        self._set_vec_or_scalar(self._lib.nlopt_set_x_weights, self._lib.nlopt_set_x_weights1, v)

    def get_x_weights(self):
        return self._get_vec(self._lib.nlopt_get_x_weights)

    def set_initial_step(self, v):
        self._set_vec_or_scalar(self._lib.nlopt_set_initial_step, self._lib.nlopt_set_initial_step1, v)

    def get_initial_step(self, x):
        xa = _as_f64(x, self._n)
        a = np.empty(self._n)
        self._check(self._lib.nlopt_get_initial_step(self._h, _ptr(xa), _ptr(a)))
        return a

    def set_default_initial_step(self, x):
        xa = _as_f64(x, self._n)
        self._check(self._lib.nlopt_set_default_initial_step(self._h, _ptr(xa)))

    def set_local_optimizer(self, lo: "opt"):
        self._check(self._lib.nlopt_set_local_optimizer(self._h, lo._h))

    def force_stop(self):
        self._check(self._lib.nlopt_force_stop(self._h))

    def set_force_stop(self, v):
        self._check(self._lib.nlopt_set_force_stop(self._h, int(v)))

    def get_force_stop(self):
        return self._lib.nlopt_get_force_stop(self._h)


def _scalar_accessors():
    # nlopt.hpp NLOPT_GETSET (:546-568)
    for name, conv in (("stopval", float), ("ftol_rel", float), ("ftol_abs", float), ("xtol_rel", float),
                       ("maxeval", int), ("maxtime", float), ("population", int), ("vector_storage", int)):
        def setter(self, v, _n=name, _c=conv):
            self._check(getattr(self._lib, "nlopt_set_" + _n)(self._h, _c(v)))

        def getter(self, _n=name):
            return getattr(self._lib, "nlopt_get_" + _n)(self._h)

        setattr(opt, "set_" + name, setter)
        setattr(opt, "get_" + name, getter)


_scalar_accessors()


def algorithm_name(a):
    return default_library().nlopt_algorithm_name(int(a)).decode()


def version_major():
    v = [C.c_int() for _ in range(3)]
    default_library().nlopt_version(*[C.byref(i) for i in v])
    return v[0].value


def device_count():
    return default_library().nlopt_b200_device_count()


__all__ = ["opt", "Library", "RoundoffLimited", "ForcedStop", "algorithm_name", "device_count",
           "NLOPT_B200_DFUNC"] + _ALG_NAMES
