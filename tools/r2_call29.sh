#!/bin/bash
# target: only the prefix of the hit list that holds the first n_to_find hits is fetched; parity of the target paths, c5 bench
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q -k "target" 2>&1 | tail -n 2
timeout 100 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_c29_bench_c5.json 2> gpurun_out/r2_c29_bench_c5.err; echo "bench c5 rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2_c29_bench_c5.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['result'])
P
