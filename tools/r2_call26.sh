#!/bin/bash
# eight GPUs: the scaling bench at N=8 (both configurations), nothing else
mkdir -p gpurun_out
O=gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_c26_bench_n8.json 2> $O/r2_c26_bench_n8.err; echo "bench n8 rc=$?"
grep '^{' $O/r2_c26_bench_n8.json | cut -c1-200
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config c5 --steps 5 --warmup 3 > $O/r2_c26_bench_c5_n8.json 2> $O/r2_c26_bench_c5_n8.err; echo "bench c5 n8 rc=$?"
grep '^{' $O/r2_c26_bench_c5_n8.json | cut -c1-200
