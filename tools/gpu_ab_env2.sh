#!/bin/bash
# usage: tools/gpu_ab_env2.sh "K1=v1 K2=v2" "K1=v3" ...  -- tests once, then the driver's bench command under each environment
mkdir -p gpurun_out/r3
python -m pytest tests -m gpu -x -q -k "patch or sampl or pipeline or describe" 2>&1 | tail -3
i=0
for e in "$@"; do
  i=$((i+1))
  env $e CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3/bench_ab$i.json 2> gpurun_out/r3/bench_ab$i.err
  echo "== $e"; python tools/show_bench.py gpurun_out/r3/bench_ab$i.json | grep -E "^value|k_sample|k_large_res|stages"
done
