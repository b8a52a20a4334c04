#!/bin/bash
mkdir -p gpurun_out/r3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r3/final_bench_n4.json 2> gpurun_out/r3/final_bench_n4.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3/final_bench_n4.json").read().strip().splitlines()[-1])
print("N=4 value %.1f e2e %.1f host_cpu %.2f rank_ms %s" % (d["value"], d["e2e"]["value"], d["host_cpu_ms_per_pair"], [round(x) for x in d["rank_ms"]]))
PY
