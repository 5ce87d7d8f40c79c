N=${1:-2}
python tools/multigpu_check.py single 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py sharded 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -14
echo "== nccl exchange"
NLOPT_B200_EXCHANGE=nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multigpu_check.py sharded 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
echo "== bench N=$N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -1 | cut -c1-3000
echo "== bench N=1"
python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-3000
