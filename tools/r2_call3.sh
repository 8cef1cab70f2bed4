#!/bin/bash
# GPU tests + regime timings (dictionary on / off), groups
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1,k1s0 26 > gpurun_out/r2_c3_regimes.log 2>&1
python tools/gpu_case.py s1,s2 26 "{'keep_dict':1}" > gpurun_out/r2_c3_regimes_keepdict.log 2>&1
python tools/gpu_case.py c3 30 > gpurun_out/r2_c3_c3.log 2>&1
cat gpurun_out/r2_c3_regimes.log gpurun_out/r2_c3_regimes_keepdict.log gpurun_out/r2_c3_c3.log | cut -c1-330
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c3_tests.log
tail -n 30 gpurun_out/r2_c3_tests.log
