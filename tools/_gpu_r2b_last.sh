#!/bin/bash
# last sanity run of the committed tree: GPU test-suite, default bench line, reference arm
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== bench c3 (default flags)"; timeout 400 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json; cut -c1-330 gpurun_out/bench_default.json
echo "== bench c3 (20 steps)"; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n1_k20.json; cut -c1-330 gpurun_out/bench_n1_k20.json
python - <<'PY'
import json
for f in ("bench_default", "bench_n1_k20"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "us/eval", round(d["roofline"]["avg_launch_us"], 2), "frac", round(d["roofline"]["frac"], 4),
          "clocks", d["clocks"], "parity", d.get("parity", {}).get("f_after_steps", {}).get("ok"), d.get("parity", {}).get("dual_eval_at_size", {}).get("ok"))
PY
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
