#!/bin/bash
# 8-GPU box: the headline at N = 8 only
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_8gpu_b.json 2> gpurun_out/r2_bench_8gpu_b.err
echo "rc=$? $(head -c 250 gpurun_out/r2_bench_8gpu_b.json)"
