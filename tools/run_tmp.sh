#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/probe_rr.py > gpurun_out/probe_rr.txt 2>&1; cat gpurun_out/probe_rr.txt
timeout 300 python -m pytest tests/test_mac_gpu.py tests/test_acm_gpu.py -x -q -m gpu > gpurun_out/t_mac.txt 2>&1; tail -3 gpurun_out/t_mac.txt
run() { tag=$1; shift; env "$@" timeout 200 python tools/probe_mac.py --reps 4 --bs 2 > gpurun_out/probe_mac_$tag.log 2>&1; echo "== $tag"; grep "rep 2\|rep 3\|prof" gpurun_out/probe_mac_$tag.log | cut -c1-420; }
run rr0 CSLAM_RR_IMPL=0
run rr1 CSLAM_RR_IMPL=1
run rr1_s2 CSLAM_RR_IMPL=1 CSLAM_RR_SWEEPS=2
run rr1_t24 CSLAM_RR_IMPL=1 CSLAM_RR_TOL2=1e-24
run rr1_prof CSLAM_RR_IMPL=1 CSLAM_LOBPCG_PROF=1
CSLAM_RR_SWEEPS=2 timeout 300 python tools/probe_mac.py --reps 1 --bs 2 --R 4 --P 5000 --m 20000 --k 100 --oracle 1 > gpurun_out/probe_mac_s2_oracle.log 2>&1; grep -c "= 0 " gpurun_out/probe_mac_s2_oracle.log; tail -1 gpurun_out/probe_mac_s2_oracle.log
CSLAM_RR_TOL2=1e-24 timeout 300 python tools/probe_mac.py --reps 1 --bs 2 --R 4 --P 5000 --m 20000 --k 100 --oracle 1 > gpurun_out/probe_mac_t24_oracle.log 2>&1; grep -c "= 0 " gpurun_out/probe_mac_t24_oracle.log; tail -1 gpurun_out/probe_mac_t24_oracle.log
