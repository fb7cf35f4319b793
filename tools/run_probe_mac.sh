set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mac_gpu.py tests/test_acm_gpu.py -x -q 2>&1 | tail -5
python tools/probe_mac.py --reps 3 2>&1 | tee gpurun_out/probe_mac_c5.log
python tools/probe_mac.py --reps 2 --bs 2 2>&1 | tee gpurun_out/probe_mac_c5_bs2.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_mac_r1.csv python tools/probe_mac.py --reps 1 --iters 2 > gpurun_out/ncu_mac.log 2>&1
