N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for shm in 1 0; do
  echo "== e2e N=$N shared_host_x=$shm"
  NLOPT_B200_SHARED_HOST_X=$shm $T --master-port 2961$shm bench.py --gpus $N --steps 20 --warmup 5 --no-parity 2>&1 | grep '^{"metric' | tail -1 | python -c '
import json,sys; d=json.loads(sys.stdin.read()); e=d["e2e"]; print("value",round(d["value"]),"e2e",round(e["value"]),"d2h MB",e["d2h_bytes_per_step"]/1e6,"cb",round(e["seconds_in_user_callbacks"],3),{k:round(v,4) for k,v in e["wall_breakdown_s"].items()})'
done
