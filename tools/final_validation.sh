#!/bin/bash
# Round-end style validation on one B200: GPU tests, smoke, both bench arms, ncu launch list + DRAM bytes of one
# persistent solve launch.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json; echo
timeout 300 python bench.py --alg mma --no-cpu > gpurun_out/bench_n1_mma.json 2>/dev/null
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; tail -c 400 gpurun_out/bench_ref.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dual_solve --launch-skip 2 --launch-count 1 --csv --log-file gpurun_out/solve_dram.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --param dual_maxeval=20 > /dev/null 2>&1
tail -4 gpurun_out/solve_dram.csv | cut -c1-300
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["config"]["algorithm"], "n", d["config"]["n"], "m", d["config"]["m"], round(d["value"]), "evals/s;", round(d["roofline"]["avg_launch_us"],1), "us per evaluation in the kernel; frac", round(d["roofline"]["frac"],4))'
for args in "--n 100000000 --steps 3" "--m 1" "--m 16" "--alg mma --m 1" "--alg mma --m 8" "--n 1000000" "--n 100000" "--n 10000" "--n 1000"; do timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e $args 2>/dev/null | python -c "$P"; done | tee gpurun_out/solve_sweep.txt
