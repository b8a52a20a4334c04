#!/bin/bash
# A/B: CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues) under 16 contexts per GPU
mkdir -p gpurun_out/r3
for c in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3/bench_conn$c.json 2> gpurun_out/r3/bench_conn$c.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r3/bench_conn$c.json"))
print($c, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["host_cpu_ms_per_pair"], 2))
PY
done
