#!/bin/bash
# round 2, session 2, call 1: pair_math parity + A/B of the compile-time variants (tools/ab_build.py)
mkdir -p gpurun_out
echo "== fastmath probe"; timeout 120 build/fastmath_probe run 2>&1 | tail -8
echo "== parity tests (product build)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -8
echo "== A/B sweeps"
for v in base pair_only nounroll minb2 product; do
  if [ $v = product ]; then unset NLOPT_B200_LIBDIR; else export NLOPT_B200_LIBDIR=$PWD/build/ab/$v; fi
  echo "-- $v"
  SWEEP_N=1e4,1e5,1e6,1250000,2500000,1e7 SWEEP_M=1,4,16 SWEEP_CPU=0 SWEEP_TAG=_$v timeout 300 python tools/sweep_c5.py 2>&1 | grep -v "^wrote" | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print('%-9s n=%-9d m=%-2d %8.2f us  %5.1f%%  value %.17g' % (r['alg'], r['n'], r['m'], r['us_per_eval'], 100 * r['frac_of_peak'], r['value']))
"
done
unset NLOPT_B200_LIBDIR
echo "== trace (product sources, instrumented build)"
for a in ccsaq mma; do timeout 200 python tools/trace_solve.py run 1250000 $a 2>&1 | tail -12; done
timeout 200 python tools/trace_solve.py run 100000 ccsaq 2>&1 | tail -12
du -sh gpurun_out
