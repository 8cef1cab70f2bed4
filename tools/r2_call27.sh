#!/bin/bash
# four GPUs: the scaling bench at N=4 (completes the 1/2/4/8 series of round 2)
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_c27_bench_n4.json 2> gpurun_out/r2_c27_bench_n4.err; echo "bench n4 rc=$?"
grep '^{' gpurun_out/r2_c27_bench_n4.json | cut -c1-200
