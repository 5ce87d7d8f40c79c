"""Thin test helper over the kernel-level C ABI (nlopt_b200_dual_*, include/nlopt_b200.h)."""
import ctypes as C

import numpy as np

from nlopt_b200._capi import c_double_p, default_library


def _p(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


class DualHandle:
    def __init__(self, variant, inst=None, n=None, m=None, synthetic_seed=None):
        self.lib = default_library()
        self.n = inst["n"] if inst is not None else n
        self.m = inst["m"] if inst is not None else m
        self.h = self.lib.nlopt_b200_dual_create(int(variant), self.n, self.m)
        if not self.h:
            raise RuntimeError("nlopt_b200_dual_create failed (no CUDA device?)")
        if inst is not None:
            self.upload(inst)
        elif synthetic_seed is not None:
            self.check(self.lib.nlopt_b200_dual_fill_synthetic(self.h, synthetic_seed))

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.lib.nlopt_b200_dual_destroy(h)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.nlopt_b200_dual_errmsg(self.h).decode())

    def upload(self, inst):
        a = {k: np.ascontiguousarray(inst[k], dtype=np.float64) for k in ("x", "lb", "ub", "sigma", "grad_f")}
        gc = np.ascontiguousarray(inst["grad_c"], dtype=np.float64) if self.m else None
        self.check(self.lib.nlopt_b200_dual_upload(self.h, _p(a["x"]), _p(a["lb"]), _p(a["ub"]), _p(a["sigma"]),
                                                   _p(a["grad_f"]), _p(gc)))
        self.set_scalars(inst["f0"], inst["rho"], inst["c0"], inst["rhoc"])

    def set_scalars(self, f0, rho, c0, rhoc):
        c0 = np.ascontiguousarray(c0 if self.m else [0.0], dtype=np.float64)
        rhoc = np.ascontiguousarray(rhoc if self.m else [0.0], dtype=np.float64)
        self.check(self.lib.nlopt_b200_dual_set_scalars(self.h, float(f0), float(rho), _p(c0), _p(rhoc)))

    def eval(self, y, want_xcur=False):
        y = np.ascontiguousarray(y if self.m else [0.0], dtype=np.float64)
        out = np.zeros(3 + max(self.m, 1))
        grad = np.zeros(max(self.m, 1))
        self.check(self.lib.nlopt_b200_dual_eval(self.h, _p(y), int(want_xcur), _p(out), _p(grad)))
        r = dict(ret=out[0], g0=out[1], w=out[2], gc=out[3:3 + self.m].copy(), grad=grad[:self.m].copy())
        if want_xcur:
            r["xcur"] = self.download("xcur")
        return r

    def download(self, which):
        a = np.zeros(self.n if which != "grad_c" else self.n * self.m)
        self.check(self.lib.nlopt_b200_dual_download(self.h, which.encode(), _p(a)))
        return a.reshape(self.m, self.n) if which == "grad_c" else a

    def time(self, y, want_xcur=False, iters=20):
        y = np.ascontiguousarray(y if self.m else [0.0], dtype=np.float64)
        ms = C.c_double(0.0)
        self.check(self.lib.nlopt_b200_dual_time(self.h, _p(y), int(want_xcur), int(iters), C.byref(ms)))
        return ms.value

    def configure(self, key, value):
        self.check(self.lib.nlopt_b200_dual_configure(self.h, key.encode(), int(value)))

    def query(self, key):
        return self.lib.nlopt_b200_dual_query(self.h, key.encode())

    def sigma_init(self, sigma_init, sigma_min):
        a = np.ascontiguousarray(sigma_init, dtype=np.float64) if sigma_init is not None else None
        self.check(self.lib.nlopt_b200_dual_sigma_init(self.h, _p(a), float(sigma_min)))

    def set_prev(self, xcur=None, xprev=None, xprevprev=None):
        arrs = [np.ascontiguousarray(v, dtype=np.float64) if v is not None else None for v in (xcur, xprev, xprevprev)]
        self.check(self.lib.nlopt_b200_dual_set_prev(self.h, *[_p(a) for a in arrs]))

    def end_outer(self, k, sigma_min=0.0, x_weights=None, xtol_abs=None):
        w = np.ascontiguousarray(x_weights, dtype=np.float64) if x_weights is not None else None
        t = np.ascontiguousarray(xtol_abs, dtype=np.float64) if xtol_abs is not None else None
        norms = np.zeros(2)
        below = C.c_int(0)
        self.check(self.lib.nlopt_b200_dual_end_outer(self.h, int(k), float(sigma_min), _p(w), _p(t), _p(norms),
                                                      C.byref(below)))
        return norms[0], norms[1], bool(below.value)

    def solve(self, y0, lo=None, hi=None, ftol_rel=0.0, maxeval=6):
        """one whole dual solve inside the persistent kernel (nlopt_b200_dual_solve); x*(y) of the final pass is left in xcur"""
        m = max(self.m, 1)
        y = np.ascontiguousarray(y0, dtype=np.float64).copy()
        lo = np.zeros(m) if lo is None else np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.full(m, 1e40) if hi is None else np.ascontiguousarray(hi, dtype=np.float64)
        out = np.zeros(3 + m)
        res, nev, kms = C.c_int(0), C.c_long(0), C.c_double(0.0)
        self.check(self.lib.nlopt_b200_dual_solve(self.h, _p(y), _p(lo), _p(hi), float(ftol_rel), int(maxeval), _p(out),
                                                  C.byref(res), C.byref(nev), C.byref(kms)))
        return dict(y=y, out=out, result=res.value, nevals=nev.value, xcur=self.download("xcur"))
