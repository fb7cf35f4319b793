set -x
CSLAM_LOBPCG_PROF=1 timeout 300 python tools/probe_mac.py --reps 2 --bs 2 2>&1 | tee gpurun_out/probe_mac_prof_final.log
