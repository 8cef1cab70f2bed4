#!/bin/bash
# event/group early-word resolution: regimes, tests, profile
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > gpurun_out/r2_c10_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> gpurun_out/r2_c10_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c10_regimes.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c10_regimes.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c10_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c10_tests.log
tail -n 4 gpurun_out/r2_c10_tests.log
for c in c3:bsi_resolve_early:early c3:bsi_attract_dense:dense s1:bsi_resolve_early:early; do
  case=${c%%:*}; r=${c#*:}; k=${r%%:*}; tag=${r##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c10_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o gpurun_out/r2_c10_prof_${case}_${tag} \
      python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c10_ncu_${case}_${tag}.log 2>&1
done
ls -la gpurun_out | grep r2_c10
