"""Run a few dual evaluations of one configuration (for ncu): profile_one.py <variant 0|1> <n> <m> [cfg] [pmax] [store]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_dual import DualHandle  # noqa: E402
import synth  # noqa: E402

variant, n, m = int(sys.argv[1]), int(float(sys.argv[2])), int(sys.argv[3])
cfg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pmax = int(sys.argv[5]) if len(sys.argv) > 5 else 0
store = int(sys.argv[6]) if len(sys.argv) > 6 else 0
h = DualHandle(variant, n=n, m=m, synthetic_seed=synth.SEED0)
i = np.arange(m, dtype=float)
h.set_scalars(1.0, 1.0, -0.1 * (i + 1), 1.0 + 0.1 * i)
h.configure("kernel_cfg", cfg)
if pmax:
    h.configure("pmax", pmax)
y = 0.5 * (i + 1)
for k in range(8):
    r = h.eval(y + 0.01 * k, want_xcur=bool(store))
print("ok", r["ret"])
