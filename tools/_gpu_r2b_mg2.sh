#!/bin/bash
# round 2, session 2: 2-GPU validation of this session's kernels: bit identity against 1 GPU, bench c3, exchange time
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
python tools/multigpu_check.py single 2>&1 | tail -1 | cut -c1-200
$T --master-port 29511 tools/multigpu_check.py sharded 2>&1 | grep -E "bit-identical|MULTIGPU_CHECK|Error|error" | tail -12
echo "== bench c3 N=$N"
$T --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/bench_n$N.json; cut -c1-400 gpurun_out/bench_n$N.json
echo "== trace N=$N (n = 2.5e6: 1.25e6 per rank)"
rm -f gpurun_out/trace_mg.txt*
NLOPT_B200_LIBDIR=$PWD/build/trace NLOPT_B200_TRACE_FILE=$PWD/gpurun_out/trace_mg.txt $T --master-port 29517 bench.py --gpus $N --n 2500000 --steps 4 --warmup 2 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
python - <<'PY'
import re, statistics, glob
for f in sorted(glob.glob("gpurun_out/trace_mg.txt*")):
    pat = re.compile(r"seen\[(-?\d+)\.\.(-?\d+)\] recs\[(-?\d+)\.\.(-?\d+)\] rank_done (-?\d+) totals (-?\d+) machine (-?\d+) next_pub (-?\d+)")
    rows = [list(map(int, m.groups())) for m in map(pat.search, open(f)) if m]
    rows = [r for r in rows if r[7] > 0]
    med = lambda i: statistics.median(r[i] for r in rows)
    print(f, len(rows), "gens | seen_hi %d rec_hi %d rank_done %d totals %d (exchange %d) machine %d next_pub %d" %
          (med(1), med(3), med(4), med(5), med(5) - med(4), med(6), med(7)))
PY
