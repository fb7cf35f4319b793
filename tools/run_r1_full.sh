set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r1.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_fp32.json 2> gpurun_out/bench_r1_fp32.err
tail -3 gpurun_out/bench_r1_fp32.err; cat gpurun_out/bench_r1_fp32.json
timeout 600 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline > gpurun_out/bench_r1_tf32.json 2>> gpurun_out/bench_r1_fp32.err
cat gpurun_out/bench_r1_tf32.json
timeout 600 python bench.py --steps 10 --warmup 3 --sparsify-every 0 --no-cpu-baseline > gpurun_out/bench_r1_nosparsify.json 2>> gpurun_out/bench_r1_fp32.err
cat gpurun_out/bench_r1_nosparsify.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nns_coarse_tc -c 6 -o gpurun_out/prof_nns_coarse_r1b -f python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 3 --check 0 > gpurun_out/ncu_full_coarse.log 2>&1
tail -3 gpurun_out/ncu_full_coarse.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
