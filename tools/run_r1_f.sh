set -x
mkdir -p gpurun_out
timeout 900 python tools/bench_kernels.py 2>&1 | grep -v Warning | tee gpurun_out/bench_kernels_r1.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -2 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_vlad|k_pca|k_gem|k_preproc" -c 12 -o gpurun_out/prof_heads_r1 -f python tools/bench_kernels.py > gpurun_out/ncu_heads.log 2>&1
tail -2 gpurun_out/ncu_heads.log
