#!/bin/bash
# final-state check: shared-memory tier reverted, target hit ordering on the host made cheap; full parity suite, both benches
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_c18_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c18_tests.log
tail -n 3 $O/r2_c18_tests.log
timeout 600 python bench.py --config c5 > $O/r2_c18_bench_c5_n1.json 2> $O/r2_c18_bench_c5_n1.err; echo "bench c5 rc=$?"
cut -c1-300 $O/r2_c18_bench_c5_n1.json
timeout 600 python bench.py > $O/r2_c18_bench_n1.json 2> $O/r2_c18_bench_n1.err; echo "bench rc=$?"
cut -c1-300 $O/r2_c18_bench_n1.json
python tools/gpu_case.py s0,s1,s2,k2s0,c3 26 > $O/r2_c18_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> $O/r2_c18_regimes.log 2>&1
cut -c1-20,150-330 $O/r2_c18_regimes.log
timeout 300 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > $O/r2_c18_c5_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_c18_launches_c5.csv \
    python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > $O/r2_c18_ncu_launches_c5.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_c18_smoke.log 2>&1; tail -n 2 $O/r2_c18_smoke.log
