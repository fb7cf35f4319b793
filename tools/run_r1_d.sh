set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -3 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1c_ref.json 2>> gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
