set -x
CSLAM_LOBPCG_PROF=1 timeout 300 python tools/probe_mac.py --reps 3 --bs 2 > gpurun_out/probe_mac_prof_final.log 2>&1; tail -3 gpurun_out/probe_mac_prof_final.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err; cut -c1-250 gpurun_out/bench_r1_n1.json
