#!/bin/bash
# N = 8 with and without NUMA pinning of the ranks (per-rank device times in the JSON line)
mkdir -p gpurun_out/r3
lscpu | grep -E "Socket|NUMA|^CPU\(s\)"; nvidia-smi topo -m 2>/dev/null | head -12
for mode in pin nopin; do
  if [ $mode = nopin ]; then export MODSGPU_NO_NUMA_PIN=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3/scale8_$mode.json 2> gpurun_out/r3/scale8_$mode.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r3/scale8_$mode.json").read().strip().splitlines()[-1])
print("$mode value %.1f e2e %.1f host_cpu %.2f numa %s rank_ms %s" % (d["value"], d["e2e"]["value"], d["host_cpu_ms_per_pair"], d.get("numa_pin"), [round(x) for x in d.get("rank_ms", [])]))
PY
done
