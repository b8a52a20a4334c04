#!/bin/bash
# the driver's N = 2 commands for both arms on the final build
mkdir -p gpurun_out/r3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3/final_bench_n2.json 2> gpurun_out/r3/final_bench_n2.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3/final_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value %.1f e2e %.1f host_cpu %.2f rank_ms %s keys %s" % (d["value"], d["e2e"]["value"], d["host_cpu_ms_per_pair"], [round(x) for x in d["rank_ms"]], sorted(d.keys())))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3/final_reference_n2.json 2> gpurun_out/r3/final_reference_n2.err
echo "reference rc=$?"; cut -c1-300 gpurun_out/r3/final_reference_n2.json
