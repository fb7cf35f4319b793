set -x
mkdir -p gpurun_out
python tools/probe_nns.py --n 1000000 --d 512 --q 64 2>&1 | tee gpurun_out/probe_c3.log
python tools/probe_nns.py --n 1000000 --d 512 --q 128 --reps 4 --check 0 2>&1 | tee gpurun_out/probe_c3_q128.log
python tools/probe_nns.py --n 250000 --d 4096 --q 64 --reps 4 --check 2 2>&1 | tee gpurun_out/probe_d4096.log
ncu --set full --clock-control none --import-source on -k regex:k_nns_coarse_tc -s 3 -c 2 -o gpurun_out/prof_nns_coarse_r1 python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 3 --check 0 > gpurun_out/ncu_run.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nns_r1.csv python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 3 --check 0 > gpurun_out/ncu_run2.log 2>&1
python -m pytest tests/test_nns_gpu.py -x -q 2>&1 | tail -5
