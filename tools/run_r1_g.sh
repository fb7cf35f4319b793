set -x
timeout 900 python -m pytest tests/test_nns_gpu.py tests/test_full_size_gpu.py tests/test_frontend_gpu.py -x -q 2>&1 | tail -3
for q in 64 256 512; do
timeout 300 python tools/probe_nns.py --n 1000000 --d 512 --q $q --reps 3 --check 0 2>&1 | tail -1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nns_coarse_pair -c 2 -o gpurun_out/prof_nns_coarse_pair_r1 -f python tools/probe_nns.py --n 1000000 --d 512 --q 256 --reps 1 --check 0 > gpurun_out/ncu_pair.log 2>&1
tail -1 gpurun_out/ncu_pair.log
