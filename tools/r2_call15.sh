#!/bin/bash
# HEAD (hybrid accounting): full parity suite, regimes, bench c4 + c5, launch list, --set full of main + resolve kernels
mkdir -p gpurun_out
O=gpurun_out
python tools/gpu_case.py s0,s1,s2,k2s0,k2s1 26 > $O/r2_c15_regimes.log 2>&1
python tools/gpu_case.py c3 26 >> $O/r2_c15_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> $O/r2_c15_regimes.log 2>&1
cut -c1-60,170-330 $O/r2_c15_regimes.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_c15_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c15_tests.log
tail -n 3 $O/r2_c15_tests.log
timeout 600 python bench.py > $O/r2_c15_bench_n1.json 2> $O/r2_c15_bench_n1.err; echo "bench rc=$?"
cut -c1-300 $O/r2_c15_bench_n1.json
timeout 600 python bench.py --config c5 > $O/r2_c15_bench_c5_n1.json 2> $O/r2_c15_bench_c5_n1.err; echo "bench c5 rc=$?"
cut -c1-300 $O/r2_c15_bench_c5_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_c15_bench_ref.json 2> $O/r2_c15_bench_ref.err; echo "bench ref rc=$?"
cut -c1-300 $O/r2_c15_bench_ref.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_c15_bench_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_c15_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_c15_ncu_launches.log 2>&1
for c in s0:bsi_attract_dense:main 's0:bsi_resolve$:tail' s0:bsi_resolve_finish:finish s1:bsi_resolve_early:early s1:bsi_resolve_mu:mu s2:bsi_resolve_mu:mu c3:bsi_resolve_early:early; do
  case=${c%%:*}; r=${c#*:}; k=${r%%:*}; tag=${r##*:}
  python tools/gpu_case.py $case 26 '{}' 2 > $O/r2_c15_plain_$case.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$k" -c 1 -o $O/r2_c15_prof_${case}_${tag} \
      python tools/gpu_case.py $case 26 '{}' 2 > $O/r2_c15_ncu_${case}_${tag}.log 2>&1
done
ls -la $O | tail -30
