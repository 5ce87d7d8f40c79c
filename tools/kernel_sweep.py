"""Kernel-only sweep on one GPU: dual-evaluation time, dual-evals/s and achieved HBM GB/s
(algorithmic bytes 8 n (5+m), +8n with x* stored) for n x m x variant x segment geometry.
Writes gpurun_out/kernel_sweep.json.  Timing: CUDA events around `iters` back-to-back launches
(after warm-up) on the launching stream; inputs larger than L2 from n >= 2e6 (flagged otherwise)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from gpu_dual import DualHandle  # noqa: E402
import synth  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    rows = []
    ns = [10**7] if quick else [10**3, 10**4, 10**5, 10**6, 10**7, 10**8]
    ms = [4] if quick else [1, 4, 16]
    pmaxes = [None]
    for n in ns:
        for m in ms:
            if 8 * n * (9 + 2 * m) > 150e9:
                continue
            for variant, name in ((0, "MMA"), (1, "CCSAQ")):
                for pmax in pmaxes:
                    h = DualHandle(variant, n=n, m=m, synthetic_seed=synth.SEED0)
                    if pmax:
                        h.configure("pmax", pmax)
                    i = np.arange(m, dtype=float)
                    h.set_scalars(1.0, 1.0, -0.1 * (i + 1), 1.0 + 0.1 * i)
                    y = 0.5 * (i + 1)
                    iters = max(5, min(200, int(2e9 / (8 * n * (5 + m)))))
                    for store in (0, 1):
                        h.time(y, store, max(3, iters // 4))
                        ms_ = min(h.time(y, store, iters) for _ in range(3))
                        byts = 8.0 * n * (5 + m + store)
                        rows.append(dict(n=n, m=m, variant=name, store_xcur=store, pmax=pmax, segments=h.query("segments"),
                                         ms=ms_, evals_per_s=1e3 / ms_, gbs=byts / ms_ / 1e6, frac=byts / ms_ / 1e6 / peak,
                                         l2_resident=bool(8 * n * (5 + m) < 120e6)))
                        print(json.dumps(rows[-1]), flush=True)
                    del h
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(peak_gbs=peak, rows=rows), open(os.path.join(ROOT, "gpurun_out", "kernel_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
