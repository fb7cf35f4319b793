set -x
timeout 900 python -m pytest tests/test_mac_gpu.py tests/test_acm_gpu.py -x -q 2>&1 | tail -3
CSLAM_MAC_PROF=1 timeout 300 python tools/probe_mac.py --reps 2 --bs 2 2>&1 | grep -E "prof|fw_subset" | tail -2
for t in 1e-32 1e-24 1e-18 1e-14; do
echo "TOL2=$t"
CSLAM_RR_TOL2=$t timeout 300 python tools/probe_mac.py --reps 3 --bs 2 2>&1 | grep -E "fw_subset" | tail -1
done
CSLAM_RR_TOL2=1e-18 timeout 300 python tools/probe_mac.py --R 4 --P 5000 --m 20000 --k 200 --reps 1 --bs 2 --oracle 1 2>&1 | tail -2
