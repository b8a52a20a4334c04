#!/bin/bash
# The driver's scaling commands on one 8-GPU box: N = 1 (plain python) and N = 8 (torchrun), --steps 20 --warmup 5
mkdir -p gpurun_out/r3
nproc; free -g | head -2
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3/scale_n1.json 2> gpurun_out/r3/scale_n1.err
for N in ${NS:-8}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3/scale_n$N.json 2> gpurun_out/r3/scale_n$N.err
  echo "N=$N rc=$?"
done
python - <<'PY'
import json, glob
base = None
for f in sorted(glob.glob("gpurun_out/r3/scale_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if d["n_gpus"] == 1: base = d["value"]
    print("N=%d value %.1f e2e %.1f host_cpu_ms_per_pair %.2f cores %s eff %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["host_cpu_ms_per_pair"], d.get("host_cores"),
          round(d["value"] / d["n_gpus"] / base, 3) if base else None))
PY
