#!/bin/bash
# last look at HEAD: smoke, the reference arm and the own arm of bench.py as the driver calls them (short)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_c28_ref.json 2> gpurun_out/r2_c28_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2_c28_ref.json
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c28_bench.json 2> gpurun_out/r2_c28_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2_c28_bench.json
timeout 120 python -m pytest tests -m gpu -x -q -k "golden and attract" 2>&1 | tail -n 2
