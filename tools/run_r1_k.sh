set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err; cut -c1-400 gpurun_out/bench_r1_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lobpcg_persist -s 3 -c 2 -o gpurun_out/prof_mac_persist_r1 -f python tools/probe_mac.py --reps 1 --iters 2 > gpurun_out/ncu_mac_full.log 2>&1
tail -2 gpurun_out/ncu_mac_full.log
