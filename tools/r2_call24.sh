#!/bin/bash
# HEAD of round 2: full parity suite, smoke, both benches, reference arm, launch list and --set full of the main kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_c24_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c24_tests.log
tail -n 3 $O/r2_c24_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_c24_smoke.log 2>&1; tail -n 1 $O/r2_c24_smoke.log
timeout 600 python bench.py > $O/r2_c24_bench_n1.json 2> $O/r2_c24_bench_n1.err; echo "bench rc=$?"
cut -c1-300 $O/r2_c24_bench_n1.json
timeout 600 python bench.py --config c5 > $O/r2_c24_bench_c5_n1.json 2> $O/r2_c24_bench_c5_n1.err; echo "bench c5 rc=$?"
cut -c1-300 $O/r2_c24_bench_c5_n1.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_c24_bench_plain.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_c24_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_c24_ncu_launches.log 2>&1
python tools/gpu_case.py s0 26 '{}' 2 > $O/r2_c24_plain_s0.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsi_attract_dense -c 1 -o $O/r2_c24_prof_s0_main \
    python tools/gpu_case.py s0 26 '{}' 2 > $O/r2_c24_ncu_s0_main.log 2>&1
