P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["config"]["algorithm"], d["config"]["n"], "evals/s", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "evals", d["dual_evals"], "us/eval", round(d["roofline"]["avg_launch_us"],2), "frac", round(d["roofline"]["frac"],4), "f", d["f_after_steps"], "eval_wall", round(d["wall_breakdown_s"]["seconds_eval_wall"]*1e3/d["steps"],3))'
build/prefetch_probe
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30
b() { echo "== bench $*"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e "$@" 2>&1 | tail -1 | python -c "$P"; }
for n in 1250000 2500000 10000000; do
  for pc in 0 3 6; do b --n $n --param b200_prefetch_chunks=$pc; done
done
b --n 1250000 --param b200_l2_keep_mb=40
b --alg mma
b --n 1000000
b --n 100000
b --n 10000
