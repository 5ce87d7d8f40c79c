"""Tuning of the warp-granular persistent dual kernel: cfg (block x unroll, min CTAs/SM) x CTAs/SM x
segment size.  Usage: tune_warp.py <n> <m,m,...>.  Prints one JSON row per point + the best per (variant, m)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_dual import DualHandle  # noqa: E402
import synth  # noqa: E402

WCFG = {100: (256, 2, 2), 101: (256, 1, 3), 102: (256, 2, 3), 103: (512, 2, 1), 104: (256, 1, 4), 105: (128, 2, 4),
        106: (256, 4, 1), 107: (512, 1, 1), 108: (256, 1, 2), 109: (128, 2, 6)}


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
            h.configure("pmax", 1 << 20)
            for tp in (512, 1024, 4096):
                h.configure("target_pairs", tp)
                for cfg, (blk, unr, minb) in WCFG.items():
                    h.configure("kernel_cfg", cfg)
                    for cps in sorted(set([minb, max(1, minb - 1), minb * 2])):
                        h.configure("ctas_per_sm", cps)
                        try:
                            h.time(y, 0, 4)
                            t = min(h.time(y, 0, 25) for _ in range(3))
                        except RuntimeError as e:
                            print("fail", name, cfg, cps, e, flush=True)
                            continue
                        byts = 8.0 * n * (5 + m)
                        rows.append(dict(n=n, m=m, variant=name, cfg=cfg, block=blk, unroll=unr, minb=minb, ctas_per_sm=cps,
                                         target_pairs=tp, segments=h.query("segments"), ms=t, gbs=byts / t / 1e6,
                                         frac=byts / t / 1e6 / peak))
                        print(json.dumps(rows[-1]), flush=True)
            del h
    best = {}
    for r in rows:
        k = (r["variant"], r["m"])
        if k not in best or r["ms"] < best[k]["ms"]:
            best[k] = r
    for k, r in best.items():
        print("BEST", k, json.dumps(r))


if __name__ == "__main__":
    main()
