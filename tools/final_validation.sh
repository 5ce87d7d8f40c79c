#!/bin/bash
# Final one-GPU validation of a round: full GPU test-suite, the default bench line (with e2e, parity, cpu baseline),
# the other workloads, the config-5 sweep with CPU references, the launch list and one ncu --set full capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-400 gpurun_out/bench_ref.json
echo "== bench c3 (20 steps)"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_n1.json
echo "== bench c3 (8 steps)"; timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n1_k8.json
echo "== bench c3 mma"; timeout 600 python bench.py --alg mma --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n1_mma.json
echo "== bench c2"; timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c2_n1.json; cut -c1-300 gpurun_out/bench_c2_n1.json
echo "== bench c4 (1 GPU)"; timeout 900 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c4_n1.json; cut -c1-300 gpurun_out/bench_c4_n1.json
echo "== sweep"; timeout 1200 python tools/sweep_c5.py 2>&1 | tail -3
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/launches_bench.log 2>&1
echo "== ncu --set full (dual_solve_kernel, 21 generations)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_solve -c 1 -o gpurun_out/prof_solve python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --param dual_maxeval=20 > gpurun_out/prof_solve.log 2>&1
ncu -i gpurun_out/prof_solve.ncu-rep --page details --csv > gpurun_out/prof_solve_details.csv 2>/dev/null
ncu -i gpurun_out/prof_solve.ncu-rep --page raw --csv > gpurun_out/prof_solve_raw.csv 2>/dev/null
echo "== ncu --set full (MMA)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dual_solve -c 1 -o gpurun_out/prof_solve_mma python bench.py --alg mma --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --param dual_maxeval=20 > gpurun_out/prof_solve_mma.log 2>&1
ncu -i gpurun_out/prof_solve_mma.ncu-rep --page raw --csv > gpurun_out/prof_solve_mma_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_solve_mma.ncu-rep --page details --csv > gpurun_out/prof_solve_mma_details.csv 2>/dev/null
# the reports themselves exceed what gpurun copies back (64 MiB for the whole directory): keep the CSV exports
rm -f gpurun_out/*.ncu-rep
echo "== L1 prefetch A/B"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["config"]["algorithm"], d["config"]["n"], "evals/s", round(d["value"]), "us/eval", round(d["roofline"]["avg_launch_us"],2), "frac", round(d["roofline"]["frac"],4), "f", d["f_after_steps"])'
b() { echo "== bench $*"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-parity "$@" 2>&1 | tail -1 | python -c "$P"; }
for n in 1250000 2500000 10000000; do for l in 0 1; do b --n $n --param b200_l1_prefetch=$l; done; done
b --alg mma --param b200_l1_prefetch=1
b --alg mma --n 1250000 --param b200_l1_prefetch=0
b --alg mma --n 1250000 --param b200_l1_prefetch=1
du -sh gpurun_out
