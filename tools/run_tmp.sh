#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1b.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r1b.log; tail -6 gpurun_out/pytest_gpu_r1b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/smoke_r1b.log
CSLAM_VLAD_TC=1 timeout 100 python tools/probe_vlad.py | tee gpurun_out/probe_vlad_tc1.json
CSLAM_VLAD_TC=0 timeout 100 python tools/probe_vlad.py | tee gpurun_out/probe_vlad_tc0.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vlad' --csv --log-file gpurun_out/vlad_launches.csv python tools/probe_vlad.py > /dev/null 2>&1; tail -4 gpurun_out/vlad_launches.csv | cut -c1-250
timeout 200 python tools/probe_sc.py --n 200000 --q 64 2>/dev/null | tee gpurun_out/probe_sc.json
timeout 200 python tools/probe_sc.py --n 200000 --q 1 2>/dev/null | tee gpurun_out/probe_sc_q1.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1b_n1.json 2> gpurun_out/bench_r1b_n1.err; tail -2 gpurun_out/bench_r1b_n1.err; cut -c1-400 gpurun_out/bench_r1b_n1.json
