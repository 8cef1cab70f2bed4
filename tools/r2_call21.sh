#!/bin/bash
# lazy attractor states, one atomic per word for target hits, small jobs on one device: full suite, both benches, captures
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_c21_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c21_tests.log
tail -n 3 $O/r2_c21_tests.log
timeout 600 python bench.py > $O/r2_c21_bench_n1.json 2> $O/r2_c21_bench_n1.err; echo "bench rc=$?"
cut -c1-300 $O/r2_c21_bench_n1.json
timeout 600 python bench.py --config c5 --no-cpu > $O/r2_c21_bench_c5_n1.json 2> $O/r2_c21_bench_c5_n1.err; echo "bench c5 rc=$?"
cut -c1-300 $O/r2_c21_bench_c5_n1.json
python tools/e2e_breakdown.py > $O/r2_c21_e2e_breakdown.log 2>&1; tail -n 5 $O/r2_c21_e2e_breakdown.log
for c in s2:bsi_attract_dense:main 's0:bsi_resolve$:tail'; do
  case=${c%%:*}; r=${c#*:}; k=${r%%:*}; tag=${r##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > $O/r2_c21_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o $O/r2_c21_prof_${case}_${tag} \
      python tools/gpu_case.py $case 26 '{}' 2 > $O/r2_c21_ncu_${case}_${tag}.log 2>&1
done
python tools/gpu_case.py n128 26 > $O/r2_c21_n128.log 2>&1; cut -c1-20,150-330 $O/r2_c21_n128.log
