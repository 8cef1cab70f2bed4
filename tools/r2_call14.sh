#!/bin/bash
# hybrid accounting, count-ordered cache, refined dictionary threshold; profile of the latency tail
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > gpurun_out/r2_c14_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> gpurun_out/r2_c14_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c14_regimes.log 2>&1
python tools/gpu_case.py s1,c3,k2s0 26 "{'tune':2}" > gpurun_out/r2_c14_regimes_tune2.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c14_regimes.log gpurun_out/r2_c14_regimes_tune2.log
timeout 900 python -m pytest tests -m gpu -x -q -k "fuzz or masters or regime or ordered" > gpurun_out/r2_c14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c14_tests.log
tail -n 3 gpurun_out/r2_c14_tests.log
for c in 's0:bsi_resolve$:tail' s0:bsi_resolve_finish:finish c3:bsi_resolve_early:early s1:bsi_resolve_early:early; do
  case=${c%%:*}; r=${c#*:}; k=${r%%:*}; tag=${r##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c14_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o gpurun_out/r2_c14_prof_${case}_${tag} \
      python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c14_ncu_${case}_${tag}.log 2>&1
done
