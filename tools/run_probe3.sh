set -x
mkdir -p gpurun_out
python -m pytest tests/test_nns_gpu.py -x -q 2>&1 | tail -5
python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 4 2>&1 | tee gpurun_out/probe_c3.log
python tools/probe_nns.py --n 250000 --d 4096 --q 64 --reps 3 --check 2 2>&1 | tee gpurun_out/probe_d4096.log
python tools/probe_nns.py --n 1000000 --d 512 --q 64 --k 1 --reps 3 --check 2 2>&1 | tee gpurun_out/probe_k1.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nns_r1b.csv python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 3 --check 0 > gpurun_out/ncu_run2.log 2>&1
python bench.py --steps 20 --warmup 3 2>&1 | tee gpurun_out/bench_first.json
