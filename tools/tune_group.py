"""Tuning of the persistent dual kernel: cfg (block x unroll, min CTAs/SM) x CTAs/SM x group size.
Usage: tune_group.py <n> <m,m,...> [quick].  One JSON row per point + the best per (variant, m)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_dual import DualHandle  # noqa: E402
import synth  # noqa: E402

CFG = {0: (256, 1, 3), 1: (256, 1, 4), 2: (256, 2, 3), 3: (256, 1, 2)}


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**7
    ms_list = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4]
    quick = len(sys.argv) > 3
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
            for tc in ((8,) if quick else (2, 8, 32, 128)):
                h.configure("target_chunks", tc)
                for cfg, (blk, unr, minb) in CFG.items():
                    h.configure("kernel_cfg", cfg)
                    for cps in sorted(set([minb, minb * 2, minb * 4])):
                        h.configure("ctas_per_sm", cps)
                        try:
                            h.time(y, 0, 4)
                            t = min(h.time(y, 0, 25) for _ in range(3))
                        except RuntimeError as e:
                            print("fail", name, cfg, cps, e, flush=True)
                            continue
                        byts = 8.0 * n * (5 + m)
                        rows.append(dict(n=n, m=m, variant=name, cfg=cfg, block=blk, unroll=unr, minb=minb, ctas_per_sm=cps,
                                         target_chunks=tc, groups=h.query("segments"), ms=t, gbs=byts / t / 1e6,
                                         frac=byts / t / 1e6 / peak))
                        print(json.dumps(rows[-1]), flush=True)
            # TMA-staged variants (kernel_cfg 10 / 11 / 12 = 3 / 2 / 4 stages)
            h.configure("target_chunks", 8)
            for cfg in (10, 11, 12):
                h.configure("kernel_cfg", cfg)
                for cps in (0, 1, 2, 3, 4, 6):
                    h.configure("ctas_per_sm", cps)
                    try:
                        h.time(y, 0, 4)
                        t = min(h.time(y, 0, 25) for _ in range(3))
                    except RuntimeError as e:
                        print("fail", name, cfg, cps, e, flush=True)
                        continue
                    byts = 8.0 * n * (5 + m)
                    rows.append(dict(n=n, m=m, variant=name, cfg=cfg, block=288, unroll=1, minb=0, ctas_per_sm=cps,
                                     target_chunks=8, groups=h.query("segments"), ms=t, gbs=byts / t / 1e6,
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
