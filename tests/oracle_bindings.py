"""ctypes bindings of the ORACLE (tests only): oracle/liboracle_port.so (our plain-C restatement)
and oracle/_ref/libref_dual.so (include-trick access to the reference's static dual_func)."""
import ctypes as C
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_SO = os.path.join(REPO, "oracle", "liboracle_port.so")
REF_SO = os.path.join(REPO, "oracle", "_ref", "libnlopt_ref.so")
REF_DUAL_SO = os.path.join(REPO, "oracle", "_ref", "libref_dual.so")

dp = C.POINTER(C.c_double)
PORT_FUNC = C.CFUNCTYPE(C.c_double, C.c_uint, dp, dp, C.c_void_p)
MMA, CCSAQ = 0, 1


def _p(a):
    return a.ctypes.data_as(dp) if a is not None else None


class DualIn(C.Structure):
    _fields_ = [("n", C.c_uint), ("m", C.c_uint), ("x", dp), ("lb", dp), ("ub", dp), ("sigma", dp),
                ("grad_f", dp), ("grad_c", dp), ("f0", C.c_double), ("rho", C.c_double),
                ("c0", dp), ("rhoc", dp)]


class DualOut(C.Structure):
    _fields_ = [("xcur", dp), ("gc", dp), ("g0", C.c_double), ("w", C.c_double)]


class Options(C.Structure):
    _fields_ = [("stopval", C.c_double), ("ftol_rel", C.c_double), ("ftol_abs", C.c_double),
                ("xtol_rel", C.c_double), ("xtol_abs", dp), ("x_weights", dp), ("maxeval", C.c_int),
                ("maxtime", C.c_double), ("inner_maxeval", C.c_int), ("rho_init", C.c_double),
                ("inner_gradients", C.c_int), ("always_improve", C.c_int), ("sigma_min", C.c_double),
                ("sigma_init", dp), ("dual_ftol_rel", C.c_double), ("dual_ftol_abs", C.c_double),
                ("dual_xtol_rel", C.c_double), ("dual_xtol_abs", C.c_double), ("dual_maxeval", C.c_int),
                ("force_stop", C.POINTER(C.c_int))]


class PortStats(C.Structure):
    _fields_ = [("numevals", C.c_int), ("dual_evals", C.c_long), ("inner_iters", C.c_int),
                ("outer_iters", C.c_int), ("dual_count_log", C.c_long * 64)]


_port = None


def port():
    global _port
    if _port is None:
        L = C.CDLL(PORT_SO)
        for name in ("port_dual_mma", "port_dual_ccsaq"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.POINTER(DualIn), dp, dp, C.POINTER(DualOut)]
        L.port_sigma_init.argtypes = [C.c_uint, dp, dp, dp, C.c_double, dp]
        L.port_sigma_init.restype = None
        L.port_sigma_update.argtypes = [C.c_int, C.c_uint, dp, dp, dp, dp, dp, C.c_double, dp]
        L.port_sigma_update.restype = None
        L.port_relstop.argtypes = [C.c_double] * 4
        L.port_stop_x.argtypes = [C.c_uint, dp, dp, dp, C.c_double, dp]
        L.port_default_options.argtypes = [C.POINTER(Options)]
        L.port_default_options.restype = None
        L.port_ccsa_minimize.argtypes = [C.c_int, C.c_uint, PORT_FUNC, C.c_void_p, C.c_uint,
                                         C.POINTER(PORT_FUNC), C.POINTER(C.c_void_p), dp, dp, dp, dp, dp,
                                         C.POINTER(Options), C.POINTER(PortStats)]
        _port = L
    return _port


def port_dual(variant, inst, y=None, want_grad=True):
    """Evaluate the port's dual function on an instance dict (tests/synth.py layout).
    Returns dict(ret, g0, w, gc, grad, xcur)."""
    n, m = inst["n"], inst["m"]
    arrs = {k: np.ascontiguousarray(inst[k], dtype=np.float64) for k in
            ("x", "lb", "ub", "sigma", "grad_f", "grad_c", "c0", "rhoc")}
    y = np.ascontiguousarray(inst["y"] if y is None else y, dtype=np.float64)
    din = DualIn(n, m, _p(arrs["x"]), _p(arrs["lb"]), _p(arrs["ub"]), _p(arrs["sigma"]),
                 _p(arrs["grad_f"]), _p(arrs["grad_c"]), inst["f0"], inst["rho"], _p(arrs["c0"]), _p(arrs["rhoc"]))
    xcur = np.empty(n)
    gc = np.empty(max(m, 1))
    grad = np.empty(max(m, 1))
    dout = DualOut(_p(xcur), _p(gc), 0.0, 0.0)
    fn = port().port_dual_mma if variant == MMA else port().port_dual_ccsaq
    ret = fn(C.byref(din), _p(y), _p(grad) if want_grad else None, C.byref(dout))
    return dict(ret=ret, g0=dout.g0, w=dout.w, gc=gc[:m].copy(), grad=grad[:m].copy(), xcur=xcur)


_refdual = None


def ref_dual_available():
    return os.path.exists(REF_DUAL_SO) and os.path.exists(REF_SO)


def ref_dual(variant, inst, y=None):
    """Same as port_dual but through the reference's own static dual_func (oracle/_ref)."""
    global _refdual
    if _refdual is None:
        L = C.CDLL(REF_DUAL_SO)
        for name in ("ref_mma_dual_eval", "ref_ccsaq_dual_eval"):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_uint, C.c_uint, dp, dp, dp, dp, dp, dp, dp, dp, C.c_double, C.c_double,
                          dp, dp, dp, dp, dp]
        _refdual = L
    n, m = inst["n"], inst["m"]
    arrs = {k: np.ascontiguousarray(inst[k], dtype=np.float64) for k in
            ("x", "lb", "ub", "sigma", "grad_f", "grad_c", "c0", "rhoc")}
    y = np.ascontiguousarray(inst["y"] if y is None else y, dtype=np.float64)
    xcur = np.empty(n)
    gc = np.empty(max(m, 1))
    grad = np.empty(max(m, 1))
    gw = np.empty(2)
    fn = _refdual.ref_mma_dual_eval if variant == MMA else _refdual.ref_ccsaq_dual_eval
    ret = fn(n, m, _p(y), _p(grad), _p(arrs["x"]), _p(arrs["lb"]), _p(arrs["ub"]), _p(arrs["sigma"]),
             _p(arrs["grad_f"]), _p(arrs["grad_c"]), inst["f0"], inst["rho"], _p(arrs["c0"]),
             _p(arrs["rhoc"]), _p(xcur), _p(gc), _p(gw))
    return dict(ret=ret, g0=gw[0], w=gw[1], gc=gc[:m].copy(), grad=grad[:m].copy(), xcur=xcur)


def port_minimize(variant, f, constraints, tols, lb, ub, x0, **opts):
    """Run the port's full solver.  f(x, grad)->float with grad (size n or 0) written in place;
    constraints: list of such callables.  Returns dict(ret, x, minf, stats)."""
    L = port()
    n = len(x0)
    m = len(constraints)

    def wrap(fn):
        def thunk(nn, x, g, _d):
            xa = np.ctypeslib.as_array(x, shape=(nn,))
            ga = np.ctypeslib.as_array(g, shape=(nn,)) if g else np.empty(0)
            return float(fn(xa, ga))
        return PORT_FUNC(thunk)

    fcb = wrap(f)
    ccbs = [wrap(c) for c in constraints]
    carr = (PORT_FUNC * max(m, 1))(*ccbs)
    cdata = (C.c_void_p * max(m, 1))()
    tol = np.ascontiguousarray(tols if m else [0.0], dtype=np.float64)
    o = Options()
    L.port_default_options(C.byref(o))
    keep = []
    for k, v in opts.items():
        if k in ("xtol_abs", "x_weights", "sigma_init"):
            if v is not None:
                a = np.ascontiguousarray(v, dtype=np.float64)
                keep.append(a)
                setattr(o, k, _p(a))
        else:
            setattr(o, k, v)
    lb = np.ascontiguousarray(lb, dtype=np.float64)
    ub = np.ascontiguousarray(ub, dtype=np.float64)
    x = np.array(x0, dtype=np.float64)
    minf = C.c_double(0.0)
    st = PortStats()
    ret = L.port_ccsa_minimize(variant, n, fcb, None, m, carr, cdata, _p(tol), _p(lb), _p(ub), _p(x),
                               C.byref(minf), C.byref(o), C.byref(st))
    return dict(ret=ret, x=x, minf=minf.value, numevals=st.numevals, dual_evals=st.dual_evals,
                inner_iters=st.inner_iters, outer_iters=st.outer_iters,
                dual_count_log=list(st.dual_count_log)[:min(st.inner_iters, 64)])
