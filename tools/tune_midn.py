"""mid-size n: group size (fill_div -> chunks per group) x launch geometry, kernel only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gpu_dual import DualHandle
import synth
for n in (10**5, 10**6, 3 * 10**6):
    for variant, name in ((1, "CCSAQ"), (0, "MMA")):
        m = 4
        h = DualHandle(variant, n=n, m=m, synthetic_seed=synth.SEED0)
        i = np.arange(m, dtype=float)
        h.set_scalars(1.0, 1.0, -0.1 * (i + 1), 1.0 + 0.1 * i)
        y = 0.5 * (i + 1)
        best = None
        for fd in (7104, 3552, 1776, 888, 444, 222, 111):
            h.configure("fill_div", fd)
            for cfg in (0, 1, 2, 3):
                h.configure("kernel_cfg", cfg)
                for cps in (0, 4, 8):
                    h.configure("ctas_per_sm", cps)
                    h.time(y, 0, 5)
                    t = min(h.time(y, 0, 40) for _ in range(3))
                    row = dict(n=n, variant=name, fill_div=fd, groups=h.query("segments"), cfg=cfg, cps=cps, us=round(t * 1e3, 2))
                    if best is None or t * 1e3 < best["us"]:
                        best = row
                    if cfg == (1 if variant else 0) and cps == 0:
                        print(json.dumps(row), flush=True)
        print("BEST", json.dumps(best), flush=True)
