#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1_final.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r1_final.log; tail -5 gpurun_out/pytest_gpu_r1_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_r1_final.log
timeout 100 python tools/probe_mac.py --reps 4 --bs 2 2>&1 | grep "rep " | cut -c1-100
