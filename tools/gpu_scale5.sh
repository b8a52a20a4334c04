#!/bin/bash
# N = 8 with the driver's command (16 contexts per rank) and with 8 contexts per rank
mkdir -p gpurun_out/r3
for wk in 16 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$((wk % 10)) bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --workers $wk > gpurun_out/r3/scale8_w$wk.json 2> gpurun_out/r3/scale8_w$wk.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r3/scale8_w$wk.json").read().strip().splitlines()[-1])
print("workers $wk: value %.1f e2e %.1f host_cpu %.2f rank_ms %s single %s" % (d["value"], d["e2e"]["value"], d["host_cpu_ms_per_pair"], [round(x) for x in d.get("rank_ms", [])], d.get("single_pair")))
PY
done
