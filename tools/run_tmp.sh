#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_sc_knn|k_sc_distance|k_sc_pick' -c 4 -o gpurun_out/r1_sc_search_ncu_full python tools/probe_sc.py --n 200000 --q 64 --reps 1 > gpurun_out/ncu_sc.log 2>&1; tail -2 gpurun_out/ncu_sc.log
