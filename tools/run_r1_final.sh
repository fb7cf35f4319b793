# Round-1 evidence run on one B200 (what profiles/ was produced with; ~6 minutes of box time).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n1_ref.json 2>> gpurun_out/bench_n1.err; cut -c1-200 gpurun_out/bench_n1_ref.json
timeout 900 python tools/bench_kernels.py 2>/dev/null > gpurun_out/bench_kernels.jsonl
timeout 100 python tools/probe_vlad.py | tee gpurun_out/probe_vlad.json
timeout 200 python tools/probe_sc.py --n 200000 --q 64 2>/dev/null | tee gpurun_out/probe_sc.json
timeout 100 python tools/probe_rr.py | tee gpurun_out/probe_rr.txt
CSLAM_LOBPCG_PROF=1 timeout 200 python tools/probe_mac.py --reps 3 --bs 2 2>&1 | tail -2
CSLAM_MAC_PROF=1 timeout 200 python tools/probe_mac.py --reps 3 --bs 2 2>&1 | grep "mac prof"
timeout 200 python tools/bench_select_host.py --solver gpu --legacy 0 2>/dev/null | tee gpurun_out/select_c5.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/fp64_lat tools/ubench/fp64_lat.cu && tools/ubench/fp64_lat
# ncu: launch list of the bench step, full captures of the dominant kernels (summaries: tools/ncu_summary.py)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 1 --warmup 1 --profile-mode
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_vlad' -c 3 -o gpurun_out/vlad_tc_ncu_full python tools/probe_vlad.py
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_sc_knn|k_sc_distance|k_sc_pick' -c 4 -o gpurun_out/sc_search_ncu_full python tools/probe_sc.py --n 200000 --q 64 --reps 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_nns_coarse' -c 2 -o gpurun_out/nns_coarse_ncu_full python tools/probe_nns.py --n 1000000 --d 512 --q 64
