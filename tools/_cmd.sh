N=${1:-2}
B="timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --resident-only --hot-only --no-cpu-baseline --no-profile"
$B > gpurun_out/d_peer_n$N.json 2> gpurun_out/d_peer_n$N.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/d_peer_n$N.json | head -1
tail -c 600 gpurun_out/d_peer_n$N.err
