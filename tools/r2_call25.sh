#!/bin/bash
# the driver's own command once more at HEAD (bench.py with its small-job extras)
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py > $O/r2_c25_bench_n1.json 2> $O/r2_c25_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2_c25_bench_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'])
print(json.dumps(d.get('small_jobs'))[:1500])
P
