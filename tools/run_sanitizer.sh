set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_nns_gpu.py -x -q -k "clusters or config1 or ragged" > gpurun_out/sanitizer_nns.log 2>&1; echo "memcheck nns rc=$?"; tail -5 gpurun_out/sanitizer_nns.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_mac_gpu.py tests/test_heads_gpu.py -x -q -k "fw_subset or fiedler or pca or gem" > gpurun_out/sanitizer_mac.log 2>&1; echo "memcheck mac/heads rc=$?"; tail -5 gpurun_out/sanitizer_mac.log
