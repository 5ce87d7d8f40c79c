#!/bin/bash
# round 2, session 2, call 3: geometry (smallest group) and the cp.async ring at large n
mkdir -p gpurun_out
echo "== solve-kernel forms (bit identity)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_equal or pair_forms or tma_staged" 2>&1 | tail -4
fmt='
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print("%-9s n=%-9d m=%-2d %8.2f us  %5.1f%%  groups %-6d value %.17g" % (r["alg"], r["n"], r["m"], r["us_per_eval"], 100 * r["frac_of_peak"], r["groups"], r["value"]))
'
sw() { echo "-- $1 [$2]"; SWEEP_CFG="$2" SWEEP_N=$3 SWEEP_M=${4:-1,4} SWEEP_CPU=0 SWEEP_TAG=_$1 timeout 400 python tools/sweep_c5.py 2>&1 | grep -v "^wrote" | python -c "$fmt"; }
MID=3e5,5e5,1e6,1250000,2500000,5e6
for g in 2 3 4 6; do sw gmin$g "group_min_chunks=$g" $MID; done
sw pmax220 "pmax=220" 1e7
sw pmax305 "pmax=305" 1e7
sw pmax880 "pmax=880" 1e7
BIG=2500000,5e6,1e7,1e8
sw big_default "" 1e7,1e8
sw big_async3 "solve_async=3" $BIG
sw big_async2 "solve_async=2" $BIG
sw big_async3_gb288 "solve_async=3,group_base=296" 1e7,1e8
du -sh gpurun_out
