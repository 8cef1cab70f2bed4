#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > gpurun_out/r2_c7_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> gpurun_out/r2_c7_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c7_regimes.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c7_regimes.log
python tools/e2e_breakdown.py > gpurun_out/r2_c7_e2e.log 2>&1; tail -8 gpurun_out/r2_c7_e2e.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c7_tests.log
tail -n 5 gpurun_out/r2_c7_tests.log
python tools/gpu_case.py c3 26 '{}' 2 > gpurun_out/r2_c7_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bsi_resolve_early|bsi_attract_dense' -c 2 -o gpurun_out/r2_c7_prof_c3 \
    python tools/gpu_case.py c3 26 '{}' 2 > gpurun_out/r2_c7_ncu_c3.log 2>&1
python tools/gpu_case.py s2 26 '{}' 2 > gpurun_out/r2_c7_plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bsi_resolve_mu' -c 1 -o gpurun_out/r2_c7_prof_s2_mu \
    python tools/gpu_case.py s2 26 '{}' 2 > gpurun_out/r2_c7_ncu_s2.log 2>&1
python tools/gpu_case.py s1 26 '{}' 2 > gpurun_out/r2_c7_plain_s1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'bsi_resolve_early|bsi_attract_dense' -c 2 -o gpurun_out/r2_c7_prof_s1 \
    python tools/gpu_case.py s1 26 '{}' 2 > gpurun_out/r2_c7_ncu_s1.log 2>&1
