#!/bin/bash
# final-checkpoint states of the found lanes by transposes: full parity suite, regimes (A/B with tune 64), bench
mkdir -p gpurun_out
O=gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,c3 26 > $O/r2_c22_regimes.log 2>&1
python tools/gpu_case.py s0,s2 26 "{'tune':64}" >> $O/r2_c22_regimes.log 2>&1
python tools/gpu_case.py s0,s2 26 "{'tune':32}" >> $O/r2_c22_regimes.log 2>&1
cut -c1-20,150-330 $O/r2_c22_regimes.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_c22_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c22_tests.log
tail -n 3 $O/r2_c22_tests.log
timeout 600 python bench.py > $O/r2_c22_bench_n1.json 2> $O/r2_c22_bench_n1.err; echo "bench rc=$?"
cut -c1-300 $O/r2_c22_bench_n1.json
