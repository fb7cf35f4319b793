set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python tools/bench_kernels.py 2>&1 | grep -v Warning > gpurun_out/bench_kernels_r1.jsonl; grep -E "A6" gpurun_out/bench_kernels_r1.jsonl | cut -c1-330
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err; cut -c1-300 gpurun_out/bench_r1_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_n1_ref.json 2>> gpurun_out/bench_r1_n1.err; cut -c1-200 gpurun_out/bench_r1_n1_ref.json
