#!/bin/bash
# round 2, first GPU call: HEAD baseline in every regime, HEAD ncu captures, the parked large-N tests
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1,k1s0 26 > gpurun_out/r2_base_regimes.log 2>&1
python tools/gpu_case.py c3 30 > gpurun_out/r2_base_c3.log 2>&1
timeout 600 python -m pytest tools/next_round/test_gpu_large.py -x -q > gpurun_out/r2_large.log 2>&1
python tools/gpu_case.py s1 26 '{}' 2 > gpurun_out/r2_plain_s1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bsi_resolve -c 8 -o gpurun_out/r2_prof_s1_resolve \
    python tools/gpu_case.py s1 26 '{}' 2 > gpurun_out/r2_ncu_s1.log 2>&1
python tools/gpu_case.py s0 26 '{}' 2 > gpurun_out/r2_plain_s0.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bsi_attract_dense -s 1 -c 1 -o gpurun_out/r2_prof_s0_dense \
    python tools/gpu_case.py s0 26 '{}' 2 > gpurun_out/r2_ncu_s0.log 2>&1
tail -3 gpurun_out/r2_base_regimes.log gpurun_out/r2_base_c3.log gpurun_out/r2_large.log
