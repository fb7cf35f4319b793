# compute-sanitizer over the kernels added in round 2 (fused Frank-Wolfe set-up / tail, two-stage
# Rayleigh-Ritz, swarm round filters, candidate key map, cross-stream NNS ordering)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_mac_gpu.py -x -q -k "fused_tail or two_stage or fw_subset or fiedler" > gpurun_out/r2_sanitizer_memcheck_mac.log 2>&1; echo "memcheck mac rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_mac.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_candidate_index_gpu.py tests/test_swarm_gpu.py -x -q -k "keymap or device_mode or round_filters or bulk_add" > gpurun_out/r2_sanitizer_memcheck_index_swarm.log 2>&1; echo "memcheck index/swarm rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck_index_swarm.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_mac_gpu.py -x -q -k "fused_tail" > gpurun_out/r2_sanitizer_racecheck_mac.log 2>&1; echo "racecheck mac rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck_mac.log
