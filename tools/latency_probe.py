"""Where does the time of one dual evaluation go when the host waits for every result?
back-to-back launches vs synchronous evaluations (with / without per-launch events)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_dual import DualHandle  # noqa: E402
import synth  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**7
m = 4
for variant, name in ((1, "CCSAQ"), (0, "MMA")):
    h = DualHandle(variant, n=n, m=m, synthetic_seed=synth.SEED0)
    i = np.arange(m, dtype=float)
    h.set_scalars(1.0, 1.0, -0.1 * (i + 1), 1.0 + 0.1 * i)
    y = 0.5 * (i + 1)
    h.time(y, 0, 10)
    b2b = h.time(y, 0, 50)
    for _ in range(5):
        h.eval(y)
    t0 = time.perf_counter()
    N = 100
    for k in range(N):
        h.eval(y + 1e-3 * k)
    sync_us = (time.perf_counter() - t0) / N * 1e6
    h.configure("time_kernels", 1)
    k0 = h.query("kernel_ns")
    t0 = time.perf_counter()
    for k in range(N):
        h.eval(y + 1e-3 * k)
    sync_ev_us = (time.perf_counter() - t0) / N * 1e6
    kern_us = (h.query("kernel_ns") - k0) / N / 1e3
    print(f"{name} n={n}: back-to-back {b2b*1e3:.1f} us/launch | synchronous {sync_us:.1f} us/eval | "
          f"synchronous+events {sync_ev_us:.1f} us/eval, event-timed kernel {kern_us:.1f} us", flush=True)
