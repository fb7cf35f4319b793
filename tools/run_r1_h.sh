set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1.log
timeout 900 python tools/bench_kernels.py 2>&1 | grep -v Warning | tee gpurun_out/bench_kernels_r1.jsonl | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; tail -2 gpurun_out/bench_r1e.err; cut -c1-1500 gpurun_out/bench_r1e.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nns_coarse_tc -c 4 -o gpurun_out/prof_nns_coarse_q64_r1 -f python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 2 --check 0 > gpurun_out/ncu_q64.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nns_coarse_tc -c 4 -o gpurun_out/prof_nns_coarse_q512_r1 -f python tools/probe_nns.py --n 1000000 --d 512 --q 512 --reps 2 --check 0 > gpurun_out/ncu_q512.log 2>&1
tail -2 gpurun_out/ncu_q512.log
