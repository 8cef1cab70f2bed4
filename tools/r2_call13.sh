#!/bin/bash
# per-lane accounting, lazy units, threshold switch: full parity suite, regimes, bench, profiles
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > gpurun_out/r2_c13_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> gpurun_out/r2_c13_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c13_regimes.log 2>&1
python tools/gpu_case.py s1,c3 26 "{'tune':8}" > gpurun_out/r2_c13_regimes_tune8.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c13_regimes.log gpurun_out/r2_c13_regimes_tune8.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c13_tests.log
tail -n 3 gpurun_out/r2_c13_tests.log
timeout 600 python bench.py > gpurun_out/r2_c13_bench_n1.json 2> gpurun_out/r2_c13_bench_n1.err; echo "bench rc=$?"
cut -c1-200 gpurun_out/r2_c13_bench_n1.json
for c in c3:bsi_resolve_early:early s1:bsi_resolve_early:early; do
  case=${c%%:*}; r=${c#*:}; k=${r%%:*}; tag=${r##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c13_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o gpurun_out/r2_c13_prof_${case}_${tag} \
      python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c13_ncu_${case}_${tag}.log 2>&1
done
