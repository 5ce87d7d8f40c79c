#!/bin/bash
# round 2, session 2, call 5: start-of-generation skew between the warps of an SM sub-partition (b200_stagger_ns)
mkdir -p gpurun_out
fmt='
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print("%-9s n=%-9d m=%-2d %8.2f us  %5.1f%%  groups %-6d value %.17g" % (r["alg"], r["n"], r["m"], r["us_per_eval"], 100 * r["frac_of_peak"], r["groups"], r["value"]))
'
sw() { echo "-- $1 [$2]"; SWEEP_CFG="$2" SWEEP_N=$3 SWEEP_M=${4:-1,4} SWEEP_CPU=0 SWEEP_TAG=_$1 timeout 400 python tools/sweep_c5.py 2>&1 | grep -v "^wrote" | python -c "$fmt"; }
N=1250000,1e6,1250000,2500000,1e7
for s in 0 50 100 200 400 800; do sw stagger$s "stagger_ns=$s" $N; done
du -sh gpurun_out
