set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mac_gpu.py tests/test_acm_gpu.py tests/test_frontend_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_mac_persist.log
for v in 0 1 2; do
  CSLAM_LOBPCG_VARIANT=$v timeout 300 python tools/probe_mac.py --reps 2 --bs 2 2>&1 | tee gpurun_out/probe_mac_persist_v${v}_bs2.log
done
CSLAM_LOBPCG_VARIANT=1 timeout 300 python tools/probe_mac.py --reps 2 --bs 1 2>&1 | tee gpurun_out/probe_mac_persist_v1_bs1.log
timeout 300 python tools/probe_mac.py --R 4 --P 5000 --m 20000 --k 200 --reps 1 --bs 2 --oracle 1 2>&1 | tail -8 | tee gpurun_out/probe_mac_mid_persist.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err
tail -3 gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
