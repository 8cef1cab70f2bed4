#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1,k1s0 26 > gpurun_out/r2_c4_regimes.log 2>&1
python tools/gpu_case.py s1,s2 26 "{'keep_dict':1}" > gpurun_out/r2_c4_regimes_keepdict.log 2>&1
python tools/gpu_case.py c3 30 > gpurun_out/r2_c4_c3.log 2>&1
cat gpurun_out/r2_c4_regimes.log gpurun_out/r2_c4_regimes_keepdict.log gpurun_out/r2_c4_c3.log | cut -c1-330
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c4_tests.log
tail -n 30 gpurun_out/r2_c4_tests.log
python tools/gpu_case.py s1 26 '{}' 2 > gpurun_out/r2_c4_plain_s1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bsi_resolve_early -c 1 -o gpurun_out/r2_c4_prof_s1_early \
    python tools/gpu_case.py s1 26 '{}' 2 > gpurun_out/r2_c4_ncu_s1.log 2>&1
