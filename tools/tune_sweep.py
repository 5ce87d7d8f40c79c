"""Launch-geometry tuning of the dual kernel on one GPU: kernel_cfg x pmax at (n, m).
Writes gpurun_out/tune_sweep.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_dual import DualHandle  # noqa: E402
import synth  # noqa: E402

CFGS = {0: "256x1", 1: "256x2", 2: "512x1", 3: "256x1,minb3", 4: "256x1,minb4", 5: "128x2", 6: "512x2", 7: "128x4",
        8: "256x2,minb2", 9: "1024x1"}


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**7
    ms_list = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4]
    peak = 6567.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for m in ms_list:
        for variant, name in ((0, "MMA"), (1, "CCSAQ")):
            h = DualHandle(variant, n=n, m=m, synthetic_seed=synth.SEED0)
            i = np.arange(m, dtype=float)
            h.set_scalars(1.0, 1.0, -0.1 * (i + 1), 1.0 + 0.1 * i)
            y = 0.5 * (i + 1)
            cfgs = list(CFGS) if m == 4 else ([0, 1, 2] if m == 1 else [0, 2, 5, 9])
            for cfg in cfgs:
                h.configure("kernel_cfg", cfg)
                for pmax in (18, 37, 74, 148, 296):
                    h.configure("pmax", pmax)
                    try:
                        h.time(y, 0, 5)
                        t = min(h.time(y, 0, 30) for _ in range(3))
                    except RuntimeError as e:
                        print("fail", name, cfg, pmax, e, flush=True)
                        continue
                    byts = 8.0 * n * (5 + m)
                    rows.append(dict(n=n, m=m, variant=name, cfg=cfg, cfg_name=CFGS[cfg], pmax=pmax, ms=t,
                                     gbs=byts / t / 1e6, frac=byts / t / 1e6 / peak))
                    print(json.dumps(rows[-1]), flush=True)
            del h
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"tune_sweep_n{n}.json"), "w"), indent=1)
    best = {}
    for r in rows:
        k = (r["variant"], r["m"])
        if k not in best or r["ms"] < best[k]["ms"]:
            best[k] = r
    for k, r in best.items():
        print("BEST", k, r["cfg_name"], "pmax", r["pmax"], f"{r['ms']*1e3:.1f} us", f"{r['frac']*100:.1f}%")


if __name__ == "__main__":
    main()
