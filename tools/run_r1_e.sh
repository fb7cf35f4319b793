set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_swarm_gpu.py -x -q 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err
tail -5 gpurun_out/bench_r1_n2.err; cat gpurun_out/bench_r1_n2.json
