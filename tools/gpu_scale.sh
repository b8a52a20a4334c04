#!/bin/bash
# Multi-GPU rehearsal of both bench arms exactly as the driver launches them: N = $1
N=${1:-2}
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-96} --warmup 3 > gpurun_out/bench_gpus$N.json 2> gpurun_out/bench_gpus$N.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_gpus$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_gpus$N.json").read().strip().splitlines()[-1])
print("N=%d value %.1f e2e %.1f ms/step %.3f workers %s launches %d clocks %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["workers_per_gpu"], d["gpu_launches"], d["clocks"]))
PY
if [ -n "$REFARM" ]; then
  /usr/bin/time -v timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps ${RSTEPS:-8} --warmup 1 > gpurun_out/ref_gpus$N.json 2> gpurun_out/ref_gpus$N.err
  echo "ref rc=$?"; tail -c 300 gpurun_out/ref_gpus$N.json; grep -E "Elapsed" gpurun_out/ref_gpus$N.err
fi
