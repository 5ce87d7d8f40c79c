P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["config"]["algorithm"], d["config"]["n"], "evals/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "evals", d["dual_evals"], "us/eval", round(d["roofline"]["avg_launch_us"],2), "frac", round(d["roofline"]["frac"],4), "f", d["f_after_steps"], "eval_wall", round(d["wall_breakdown_s"]["seconds_eval_wall"]*1e3/d["steps"],3))'
b() { echo "== bench $*"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-parity "$@" 2>&1 | tail -1 | python -c "$P"; }
for n in 1250000 2500000 10000000; do for u in 1 2; do b --n $n --param b200_solve_unroll=$u; done; done
b --alg mma --param b200_solve_unroll=2
b --alg mma --n 1250000 --param b200_solve_unroll=1
b --alg mma --n 1250000 --param b200_solve_unroll=2
b --n 1250000 --param b200_l2_keep_mb=100
b --n 100000
b --n 10000
python tools/trace_solve.py run 1250000 ccsaq
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
