#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > gpurun_out/r2_c6_regimes.log 2>&1
python tools/gpu_case.py s1,s2,k2s0 26 "{'tune':1}" > gpurun_out/r2_c6_regimes_tune1.log 2>&1
python tools/gpu_case.py c3 30 > gpurun_out/r2_c6_c3.log 2>&1
python tools/gpu_case.py c3 30 "{'tune':1}" >> gpurun_out/r2_c6_c3.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c6_regimes.log gpurun_out/r2_c6_regimes_tune1.log gpurun_out/r2_c6_c3.log
python tools/e2e_breakdown.py > gpurun_out/r2_c6_e2e.log 2>&1; tail -8 gpurun_out/r2_c6_e2e.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c6_tests.log
tail -n 5 gpurun_out/r2_c6_tests.log
