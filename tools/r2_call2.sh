#!/bin/bash
# GPU tests + regime timings (dictionary on / off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c2_tests.log
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1,k1s0 26 > gpurun_out/r2_c2_regimes.log 2>&1
python tools/gpu_case.py s1,s2,k2s0 26 "{'no_dict':1}" > gpurun_out/r2_c2_regimes_nodict.log 2>&1
python tools/gpu_case.py s1,s2 26 "{'keep_dict':1}" > gpurun_out/r2_c2_regimes_keepdict.log 2>&1
python tools/gpu_case.py c3 30 > gpurun_out/r2_c2_c3.log 2>&1
tail -n 3 gpurun_out/r2_c2_tests.log; cat gpurun_out/r2_c2_regimes.log gpurun_out/r2_c2_regimes_nodict.log gpurun_out/r2_c2_regimes_keepdict.log gpurun_out/r2_c2_c3.log
