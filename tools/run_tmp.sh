#!/bin/bash
# scratch GPU call: new tests + probes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_scancontext_gpu.py tests/test_acm_gpu.py tests/test_frontend_gpu.py -x -q -m gpu > gpurun_out/t_sc.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/t_sc.txt
tail -15 gpurun_out/t_sc.txt
timeout 200 python tools/probe_sc.py --n 200000 --q 64 > gpurun_out/probe_sc.json 2> gpurun_out/probe_sc.err
cat gpurun_out/probe_sc.json; tail -3 gpurun_out/probe_sc.err
timeout 200 python tools/bench_select_host.py --solver gpu --legacy 0 > gpurun_out/select_c5.json 2> gpurun_out/select_c5.err
cat gpurun_out/select_c5.json; tail -3 gpurun_out/select_c5.err
