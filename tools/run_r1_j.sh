set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1.log
timeout 900 python tools/bench_kernels.py 2>&1 | grep -v Warning | tee gpurun_out/bench_kernels_r1.jsonl | grep -E "A1|A3|A4|A5|512 queries" | cut -c1-230
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -2 gpurun_out/bench_r1f.err; cut -c1-1200 gpurun_out/bench_r1f.json
