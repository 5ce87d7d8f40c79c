#!/bin/bash
# One-box validation: GPU tests, then a handful of bench variants summarised one line each.
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["config"]["algorithm"], round(d["value"]), round(d["ms_per_step"],2), d["dual_evals"], round(d["roofline"]["avg_launch_us"],1), round(d["roofline"]["frac"],4), d.get("e2e") and (round(d["e2e"]["value"]), {k:round(v,4) for k,v in d["e2e"].get("wall_breakdown_s",{}).items()}))'
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
b() { echo "== bench $*"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu "$@" 2>/dev/null | python -c "$P"; }
b --no-e2e
b --alg mma --no-e2e
for tc in 2 3 4 6; do b --no-e2e --param b200_target_chunks=$tc; done
b --alg mma --no-e2e --param b200_target_chunks=3
b --n 1000000 --no-e2e
b --n 1250000 --no-e2e
b --n 100000 --no-e2e
b --n 10000 --no-e2e
for t in "$@"; do timeout 200 python tools/trace_solve.py run $t ccsaq; done
