#!/bin/bash
# 8-GPU box: layouts of the retrieval headline side by side (R catalog shards x N / R query groups)
run() { n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-extras "$@" > gpurun_out/r2_layout_${tag}.json 2> gpurun_out/r2_layout_${tag}.err
  echo "$tag rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_layout_${tag}.json').read().strip().splitlines()[-1]); r=d['roofline']
    print(round(d['value']), 'q/s  e2e', round(d['e2e']['value']), ' sweep', round(r['launch_ms'],1), 'ms share', round(r['sweep_share_of_step'],3), 'TF/s', round(r['achieved']), 'clk', d['clocks']['sm_mhz'], 'recall', d['recall_at_k_sampled']['value'], d['layout'])
except Exception as e: print('ERR', e)
PY
)"; tail -2 gpurun_out/r2_layout_${tag}.err | cut -c1-300; }
run 8 n8_r4
run 8 n8_r8 --retrieval-shards 8
run 8 n8_r2 --retrieval-shards 2
run 4 n4_r2
run 4 n4_r4 --retrieval-shards 4
