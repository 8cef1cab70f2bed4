#!/bin/bash
# advance of P(c0) by latch events (bsi_mu_word): full parity suite, regimes; small-job latency with stage times
mkdir -p gpurun_out
O=gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,c3 26 > $O/r2_c20_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> $O/r2_c20_regimes.log 2>&1
python tools/gpu_case.py s0,s2 26 "{'no_dict':1}" >> $O/r2_c20_regimes.log 2>&1
cut -c1-20,150-330 $O/r2_c20_regimes.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_c20_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c20_tests.log
tail -n 3 $O/r2_c20_tests.log
timeout 600 python tools/cold_start.py > $O/r2_c20_cold_start.json 2> $O/r2_c20_cold_start.err; cat $O/r2_c20_cold_start.json
nvidia-smi -q | grep -i "persistence" | head -2
