#!/bin/bash
# N-GPU bench under several host settings: "name:ENV=..;ENV2=..:extra args" entries in VARIANTS
N=${1:-8}
mkdir -p gpurun_out
IFS='|' read -ra VS <<< "${VARIANTS:-default::}"
port=29520
for v in "${VS[@]}"; do
  name=${v%%:*}; rest=${v#*:}; envs=${rest%%:*}; extra=${rest#*:}
  port=$((port+1))
  env $(echo $envs | tr ';' ' ') timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps ${STEPS:-96} --warmup 3 $extra > gpurun_out/s2_${name}_gpus$N.json 2> gpurun_out/s2_${name}_gpus$N.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s2_${name}_gpus$N.json").read().strip().splitlines()[-1])
    print("%-16s N=%d value %.1f e2e %.1f workers %s host_cpu_ms/step(rank0) %.2f cores %s" % ("$name", d["n_gpus"], d["value"], d["e2e"]["value"], d["config"]["workers_per_gpu"], d.get("host_cpu_ms_per_step",-1), d.get("host_cores")))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/s2_${name}_gpus$N.err").read()[-600:])
PY
done
