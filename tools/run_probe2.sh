set -x
mkdir -p gpurun_out
python -m pytest tests/test_nns_gpu.py -x -q 2>&1 | tail -15
python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 5 2>&1 | tee gpurun_out/probe_c3.log
python tools/probe_nns.py --n 1000000 --d 512 --q 128 --reps 3 --check 0 2>&1 | tee gpurun_out/probe_c3_q128.log
python tools/probe_nns.py --n 250000 --d 4096 --q 64 --reps 3 --check 2 2>&1 | tee gpurun_out/probe_d4096.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nns_r1b.csv python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 3 --check 0 > gpurun_out/ncu_run2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_nns_select|k_nns_tau_select" -s 4 -c 2 -o gpurun_out/prof_nns_select python tools/probe_nns.py --n 1000000 --d 512 --q 64 --reps 3 --check 0 > gpurun_out/ncu_run3.log 2>&1
