#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0 26 > gpurun_out/r2_c8_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> gpurun_out/r2_c8_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c8_regimes.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c8_regimes.log
python tools/gpu_case.py c3 26 '{}' 2 > gpurun_out/r2_c8_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bsi_resolve_early' -c 1 -o gpurun_out/r2_c8_prof_c3 \
    python tools/gpu_case.py c3 26 '{}' 2 > gpurun_out/r2_c8_ncu_c3.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c8_tests.log
tail -n 4 gpurun_out/r2_c8_tests.log
