#!/bin/bash
# round 2, session 2: final one-GPU validation -- full GPU test-suite, bench lines, ncu --set full captures, launch list, sweep
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench c3 (20 steps)"; timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_n1.json; cut -c1-330 gpurun_out/bench_n1.json
echo "== ncu --set full (dual_solve_kernel, CCSAQ, 21 generations)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_solve -c 1 -o gpurun_out/prof_solve python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --param dual_maxeval=20 > gpurun_out/prof_solve.log 2>&1
ncu -i gpurun_out/prof_solve.ncu-rep --page details --csv > gpurun_out/prof_solve_details.csv 2>/dev/null
ncu -i gpurun_out/prof_solve.ncu-rep --page raw --csv > gpurun_out/prof_solve_raw.csv 2>/dev/null
echo "== launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-parity > gpurun_out/launches_bench.log 2>&1
echo "== ncu --set full (MMA)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dual_solve -c 1 -o gpurun_out/prof_solve_mma python bench.py --alg mma --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity --param dual_maxeval=20 > gpurun_out/prof_solve_mma.log 2>&1
ncu -i gpurun_out/prof_solve_mma.ncu-rep --page raw --csv > gpurun_out/prof_solve_mma_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
echo "== reference arm"; timeout 400 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
echo "== bench c3 mma"; timeout 300 python bench.py --alg mma --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_n1_mma.json; cut -c1-200 gpurun_out/bench_n1_mma.json
echo "== bench c2"; timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_c2_n1.json; cut -c1-200 gpurun_out/bench_c2_n1.json
echo "== bench c4 (1 GPU)"; timeout 600 python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_c4_n1.json; cut -c1-200 gpurun_out/bench_c4_n1.json
echo "== sweep"; SWEEP_N=1e3,1e4,1e5,1e6,1e7,1e8 timeout 900 python tools/sweep_c5.py 2>&1 | tail -40 | cut -c1-260
du -sh gpurun_out
