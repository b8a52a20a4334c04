#!/bin/bash
mkdir -p gpurun_out/r3
python -m pytest tests -m gpu -x -q -k "detect or blur or pipeline or mods" 2>&1 | tail -3
for wk in 16 24 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=32 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workers $wk > gpurun_out/r3/bench_wk$wk.json 2> gpurun_out/r3/bench_wk$wk.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r3/bench_wk$wk.json"))
print($wk, round(d["value"], 1), round(d["e2e"]["value"], 1), round(d["host_cpu_ms_per_pair"], 2), d["stages"]["detect"]["ms_per_pair"])
PY
done
