#!/bin/bash
# two GPUs: one-process groups, torchrun forms (table growth across ranks, cached engine), bench at N=2, small-job latency
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L > $O/r2_c19_gpus.log
timeout 900 python -m pytest tests -m gpu -x -q -k "group or engine or run_forms or masters" > $O/r2_c19_tests.log 2>&1; echo "tests rc=$?" >> $O/r2_c19_tests.log
tail -n 3 $O/r2_c19_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/torchrun_checks.py > $O/r2_c19_torchrun_checks.log 2>&1; echo "torchrun checks rc=$?"
tail -n 3 $O/r2_c19_torchrun_checks.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2_c19_bench_n2.json 2> $O/r2_c19_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-250 $O/r2_c19_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config c5 --steps 5 --warmup 3 > $O/r2_c19_bench_c5_n2.json 2> $O/r2_c19_bench_c5_n2.err; echo "bench c5 n2 rc=$?"
cut -c1-250 $O/r2_c19_bench_c5_n2.json
timeout 300 python tools/group_scaling.py > $O/r2_c19_group_scaling.log 2>&1; tail -n 6 $O/r2_c19_group_scaling.log
timeout 600 python tools/cold_start.py > $O/r2_c19_cold_start.json 2> $O/r2_c19_cold_start.err; cat $O/r2_c19_cold_start.json
