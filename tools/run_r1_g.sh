set -x
timeout 900 python -m pytest tests/test_nns_gpu.py -x -q 2>&1 | tail -3
CSLAM_NNS_DEBUG=1 timeout 300 python tools/probe_nns.py --n 1000000 --d 512 --q 512 --reps 3 --check 0 2>&1 | grep -E "cslam nns|rep 2" | tail -2
for q in 64 128 256 512 1024; do
timeout 300 python tools/probe_nns.py --n 1000000 --d 512 --q $q --reps 4 --check 2 2>&1 | tail -2
done
timeout 300 python tools/probe_nns.py --n 1000000 --d 128 --q 128 --reps 3 --check 0 2>&1 | tail -1
timeout 300 python tools/probe_nns.py --n 250000 --d 4096 --q 64 --reps 3 --check 0 2>&1 | tail -1
timeout 300 python tools/probe_nns.py --n 250000 --d 4096 --q 512 --reps 3 --check 0 2>&1 | tail -1
