timeout 600 python -m pytest tests/test_mac_gpu.py tests/test_acm_gpu.py -x -q 2>&1 | tail -2
CSLAM_LOBPCG_PROF=1 timeout 300 python tools/probe_mac.py --reps 3 --bs 2 2>&1 | grep -E "prof|fw_subset" | tail -3
timeout 300 python tools/probe_mac.py --R 4 --P 5000 --m 20000 --k 200 --reps 1 --bs 2 --oracle 1 2>&1 | tail -1
