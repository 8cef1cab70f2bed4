#!/bin/bash
# validation of the tree after the out-of-line dictionary refactor + captures for profiles/
mkdir -p gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0 26 > gpurun_out/r2_c9_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> gpurun_out/r2_c9_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c9_regimes.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c9_regimes.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c9_tests.log
tail -n 4 gpurun_out/r2_c9_tests.log
timeout 600 python bench.py > gpurun_out/r2_c9_bench_n1.json 2> gpurun_out/r2_c9_bench_n1.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r2_c9_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_c9_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_c9_ncu_launch.log 2>&1
for c in s0:bsi_attract_dense s1:bsi_resolve_mu c3:bsi_resolve_early; do
  case=${c%%:*}; k=${c##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c9_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o gpurun_out/r2_c9_prof_${case} \
      python tools/gpu_case.py $case 26 '{}' 2 > gpurun_out/r2_c9_ncu_$case.log 2>&1
done
ls -la gpurun_out | tail -12
