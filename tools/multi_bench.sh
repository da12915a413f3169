#!/bin/bash
# 8-GPU box: NCCL tests on 2 ranks, then the bench at N = 8 (full), 4 and 2 (headline only)
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_multi_tests.log 2>&1; tail -2 gpurun_out/r2_multi_tests.log
for n in 8 4 2; do
  extra=""; [ $n != 8 ] && extra="--no-extras"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 $extra > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err
  echo "N=$n rc=$? $(head -c 300 gpurun_out/r2_bench_${n}gpu.json)"
done
