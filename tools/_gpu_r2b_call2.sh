#!/bin/bash
# round 2, session 2, call 2: corrected pair_math range test, parity, runtime A/B of the solve kernel's forms
mkdir -p gpurun_out
echo "== fastmath probe"; timeout 120 build/fastmath_probe run 2>&1 | tail -8
echo "== parity tests (product build)"
timeout 1000 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q --maxfail=6 2>&1 | tail -25
echo "== A/B sweeps"
fmt='
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print("%-9s n=%-9d m=%-2d %8.2f us  %5.1f%%  value %.17g" % (r["alg"], r["n"], r["m"], r["us_per_eval"], 100 * r["frac_of_peak"], r["value"]))
'
sw() { echo "-- $1 [$2]"; SWEEP_CFG="$2" SWEEP_N=${3:-1e4,1e5,1e6,1250000,2500000,1e7} SWEEP_M=${4:-1,4,16} SWEEP_CPU=0 SWEEP_TAG=_$1 timeout 300 python tools/sweep_c5.py 2>&1 | grep -v "^wrote" | python -c "$fmt"; }
NLOPT_B200_LIBDIR=$PWD/build/ab/base sw base "" 1e4,1250000,1e7 1,4
sw default ""
sw minb3 "solve_minb=3"
sw minb2 "solve_minb=2"
sw async3 "solve_async=3"
sw async2 "solve_async=2"
sw async3_gb288 "solve_async=3,group_base=288" 1e6,1250000,2500000,1e7 4
sw shard8_default "pmax=55" 1250000 1,4
sw shard8_async3 "pmax=55,solve_async=3" 1250000 1,4
sw shard8_async3_36 "pmax=36,solve_async=3" 1250000 1,4
echo "== trace (product sources, instrumented build)"
timeout 200 python tools/trace_solve.py run 1250000 ccsaq 2>&1 | tail -11
timeout 200 python tools/trace_solve.py run 1250000 ccsaq --param b200_solve_async=3 2>&1 | tail -11
timeout 200 python tools/trace_solve.py run 100000 ccsaq 2>&1 | tail -11
du -sh gpurun_out
