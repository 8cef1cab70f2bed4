#!/bin/bash
# two GPUs: groups, masters over all devices, torchrun bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_c5_gpus.log 2>&1
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_masters.py -x -q > gpurun_out/r2_c5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_c5_tests.log
tail -n 15 gpurun_out/r2_c5_tests.log
python tools/gpu_case.py s1,s2 26 > gpurun_out/r2_c5_regimes.log 2>&1
python tools/gpu_case.py c3 30 >> gpurun_out/r2_c5_regimes.log 2>&1
cut -c1-60,170-330 gpurun_out/r2_c5_regimes.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_c5_bench_n2.json 2> gpurun_out/r2_c5_bench_n2.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2_c5_bench_n2.json; tail -5 gpurun_out/r2_c5_bench_n2.err
timeout 300 python tools/group_scaling.py > gpurun_out/r2_c5_group.log 2>&1; cat gpurun_out/r2_c5_group.log
