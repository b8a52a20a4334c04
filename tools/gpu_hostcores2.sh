#!/bin/bash
# Host-core sensitivity on one GPU: the driver's bench command pinned to 4 cores (the 8-GPU box has 4 per rank)
mkdir -p gpurun_out/r3
for spec in "16:16" "4:16" "4:12" "4:8" "3:16" "2:16"; do
  IFS=: read -r cores wk <<< "$spec"
  taskset -c 0-$((cores-1)) python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workers $wk > gpurun_out/r3/hc_c${cores}_w$wk.json 2> gpurun_out/r3/hc_c${cores}_w$wk.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r3/hc_c${cores}_w$wk.json").read().strip().splitlines()[-1])
print("cores $cores workers $wk: value %.1f e2e %.1f host_cpu_ms_per_pair %.2f" % (d["value"], d["e2e"]["value"], d["host_cpu_ms_per_pair"]))
PY
done
MODSGPU_HOST_PROFILE=1 taskset -c 0-3 python tools/ncu_target.py 6 2>&1 | tail -3
