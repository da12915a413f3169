#!/bin/bash
# Debugging aid: which switch makes the "vector::reserve" failure of the full GPU suite go away?
T="tests/test_gpu_configs.py tests/test_gpu_eval.py tests/test_gpu_hash.py"
run() { name=$1; shift; env "$@" timeout 300 python -m pytest $T -x -q -m gpu > gpurun_out/bis_$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/bis_$name.log)"; }
run base A=1
run nofork XB_FORK=0
run nowg XB_WG=0
run nort XB_RT=0
run mcheck MALLOC_CHECK_=3
run pymalloc PYTHONMALLOC=debug
timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_eval.py tests/test_gpu_hash.py -x -q -m gpu -k "not recall" > gpurun_out/bis_norecall.log 2>&1; echo "norecall rc=$? $(tail -1 gpurun_out/bis_norecall.log)"
timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_eval.py tests/test_gpu_hash.py -x -q -m gpu -k "not bundle" > gpurun_out/bis_nobundle.log 2>&1; echo "nobundle rc=$? $(tail -1 gpurun_out/bis_nobundle.log)"
