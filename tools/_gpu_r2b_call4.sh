#!/bin/bash
# round 2, session 2, call 4: the cp.async ring on a 296-group geometry (2 CTAs/SM = 295 sweepers + folder)
mkdir -p gpurun_out
fmt='
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print("%-9s n=%-9d m=%-2d %8.2f us  %5.1f%%  groups %-6d value %.17g" % (r["alg"], r["n"], r["m"], r["us_per_eval"], 100 * r["frac_of_peak"], r["groups"], r["value"]))
'
sw() { echo "-- $1 [$2]"; SWEEP_CFG="$2" SWEEP_N=$3 SWEEP_M=${4:-1,4} SWEEP_CPU=0 SWEEP_TAG=_$1 timeout 400 python tools/sweep_c5.py 2>&1 | grep -v "^wrote" | python -c "$fmt"; }
sw A_gmin4 "group_min_chunks=4" 1250000,1250000,1e7
sw B_async2_296 "solve_async=2,group_base=296,pmax=37" 1250000,1250000
sw C_async3_296 "solve_async=3,group_base=296,pmax=37" 1250000,1250000
sw D_roomy_296 "solve_minb=2,group_base=296,pmax=37" 1250000,1250000
sw E_async2_2368 "solve_async=2,group_base=296,pmax=296" 1e7,1e7
sw F_async3_2368 "solve_async=3,group_base=296,pmax=296" 1e7,1e7
sw G_roomy_2368 "solve_minb=2,group_base=296,pmax=296" 1e7,1e7
sw H_async2_gmin4 "solve_async=2,group_base=296,group_min_chunks=4" 6e5,1e6,2500000,5e6
sw I_async3_gmin4 "solve_async=3,group_base=296,group_min_chunks=4" 6e5,1e6,2500000,5e6
sw J_default_gmin4 "group_min_chunks=4" 6e5,1e6,2500000,5e6
du -sh gpurun_out
