#!/bin/bash
# usage: tools/gpu_ab_env.sh VAR v1 v2 ...  -- the driver's bench command under each value of an environment switch
VAR=$1; shift
mkdir -p gpurun_out/r3
python -m pytest tests -m gpu -x -q -k "patch or sampl or pipeline or describe" 2>&1 | tail -3
for v in "$@"; do
  env $VAR=$v CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3/bench_${VAR}_$v.json 2> gpurun_out/r3/bench_${VAR}_$v.err
  echo "== $VAR=$v"; python tools/show_bench.py gpurun_out/r3/bench_${VAR}_$v.json | grep -E "^value|k_sample|k_large|stages"
done
