#!/bin/bash
# multi-GPU validation + measurements on N GPUs of one box: bit-identity check, bench c3, bench c4, config-5 sweep
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='grep -v "^\*\|OMP_NUM\|^$"'
mkdir -p gpurun_out
if [ "${CHECK:-1}" = "1" ]; then
  python tools/multigpu_check.py single 2>&1 | tail -1 | cut -c1-200
  $T --master-port 29511 tools/multigpu_check.py sharded 2>&1 | grep -E "bit-identical|MULTIGPU_CHECK|Error|error" | tail -12
fi
echo "== bench c3 N=$N"
$T --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/bench_n$N.json; cut -c1-250 gpurun_out/bench_n$N.json
echo "== bench c3 mma N=$N"
$T --master-port 29514 bench.py --gpus $N --alg mma --steps 8 --warmup 3 --no-e2e 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/bench_mma_n$N.json; cut -c1-250 gpurun_out/bench_mma_n$N.json
echo "== bench c4 N=$N"
$T --master-port 29515 bench.py --gpus $N --workload c4 --steps 6 --warmup 3 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/bench_c4_n$N.json; cut -c1-250 gpurun_out/bench_c4_n$N.json
echo "== sweep N=$N"
SWEEP_N=${SWEEP_N:-1e5,1e6,1e7,1e8} $T --master-port 29516 tools/sweep_c5.py 2>&1 | grep -E "wrote|rror" | tail -3
