for v in "VFS_OVERLAP=0" "VFS_OVERLAP=1"; do
echo "== $v"
env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['rhs_only']['ms'],3), round(d['les_only']['ms'],3))"
done
