#!/bin/bash
# streaming loads of the read-once queues in the throughput-mode resolution kernels (A/B: tune 128 = plain loads)
mkdir -p gpurun_out
O=gpurun_out
python tools/gpu_case.py s1,s2,k2s0,k2s1,c3 26 > $O/r2_c23_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> $O/r2_c23_regimes.log 2>&1
python tools/gpu_case.py s1,s2,k2s0,k2s1,c3 26 "{'tune':128}" > $O/r2_c23_regimes_plain_loads.log 2>&1
python tools/gpu_case.py c3 30 "{'tune':128}" >> $O/r2_c23_regimes_plain_loads.log 2>&1
cut -c1-20,150-330 $O/r2_c23_regimes.log $O/r2_c23_regimes_plain_loads.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "fuzz or masters or regime or ordered or published or thousands or dictionary or seeded or headline" > $O/r2_c23_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c23_tests.log
tail -n 3 $O/r2_c23_tests.log
python tools/gpu_case.py s1 26 '{}' 2 > $O/r2_c23_plain_s1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsi_resolve_early -c 1 -o $O/r2_c23_prof_s1_early \
    python tools/gpu_case.py s1 26 '{}' 2 > $O/r2_c23_ncu_s1_early.log 2>&1
