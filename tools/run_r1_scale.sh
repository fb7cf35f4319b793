set -x
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_swarm_gpu.py -x -q 2>&1 | tail -3; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r1_n$N.json 2> gpurun_out/bench_r1_n$N.err
tail -3 gpurun_out/bench_r1_n$N.err; cut -c1-200 gpurun_out/bench_r1_n$N.json
