N=${1:-2}
B="timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --resident-only --hot-only --no-cpu-baseline --no-profile"
$B 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
CTI_NO_COLL=1 $B 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
