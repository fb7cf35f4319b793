set -x
timeout 900 python -m pytest tests/test_mac_gpu.py tests/test_acm_gpu.py -x -q 2>&1 | tail -3
for cm in 0 1; do
echo "COARSE=$cm"
CSLAM_MAC_COARSE=$cm timeout 300 python tools/probe_mac.py --reps 3 --bs 2 2>&1 | grep -E "fw_subset|single" | tail -2
CSLAM_MAC_COARSE=$cm timeout 300 python tools/probe_mac.py --reps 2 --bs 1 2>&1 | grep -E "fw_subset" | tail -1
done
CSLAM_MAC_PROF=1 CSLAM_LOBPCG_PROF=1 timeout 300 python tools/probe_mac.py --reps 2 --bs 2 2>&1 | grep -E "prof" | tail -2
timeout 300 python tools/probe_mac.py --R 4 --P 5000 --m 20000 --k 200 --reps 1 --bs 2 --oracle 1 2>&1 | tail -2
