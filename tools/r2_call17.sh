#!/bin/bash
# shared-memory tier of groups / attractor cache (A/B against tune 16 = tier off); one-pass target run
mkdir -p gpurun_out
O=gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > $O/r2_c17_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> $O/r2_c17_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> $O/r2_c17_regimes.log 2>&1
python tools/gpu_case.py s1,k2s0,c3 26 "{'tune':16}" > $O/r2_c17_regimes_tune16.log 2>&1
python tools/gpu_case.py c3 30 "{'tune':16}" >> $O/r2_c17_regimes_tune16.log 2>&1
cut -c1-60,170-330 $O/r2_c17_regimes.log $O/r2_c17_regimes_tune16.log
timeout 1200 python -m pytest tests -m gpu -x -q -k "target or fuzz or masters or regime or ordered or published or thousands or group" > $O/r2_c17_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c17_tests.log
tail -n 3 $O/r2_c17_tests.log
timeout 600 python bench.py --config c5 --no-cpu > $O/r2_c17_bench_c5_n1.json 2> $O/r2_c17_bench_c5_n1.err; echo "bench c5 rc=$?"
cut -c1-300 $O/r2_c17_bench_c5_n1.json
for c in c3:bsi_resolve_early:early s1:bsi_resolve_early:early; do
  case=${c%%:*}; r=${c#*:}; k=${r%%:*}; tag=${r##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > $O/r2_c17_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o $O/r2_c17_prof_${case}_${tag} \
      python tools/gpu_case.py $case 26 '{}' 2 > $O/r2_c17_ncu_${case}_${tag}.log 2>&1
done
ls -la $O | grep c17
