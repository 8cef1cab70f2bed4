#!/bin/bash
# target kernel with the tight loop past the last perturbation: parity (target tests, fuzz, masters, published), c5 bench, profile
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "target or fuzz or masters or published or group" > $O/r2_c16_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c16_tests.log
tail -n 3 $O/r2_c16_tests.log
timeout 600 python bench.py --config c5 > $O/r2_c16_bench_c5_n1.json 2> $O/r2_c16_bench_c5_n1.err; echo "bench c5 rc=$?"
cut -c1-300 $O/r2_c16_bench_c5_n1.json
timeout 300 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > $O/r2_c16_c5_plain.json 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsi_target_main -c 1 -o $O/r2_c16_prof_c5_target \
    python bench.py --config c5 --steps 3 --warmup 3 --no-cpu > $O/r2_c16_ncu_c5.log 2>&1
ls -la $O | grep c16
