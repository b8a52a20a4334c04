#!/bin/bash
# bench at several CNN chunk sizes / worker counts (no code change: MODSGPU_CNN_CHUNK)
mkdir -p gpurun_out
for c in ${CHUNKS:-512 1024 2048 4096}; do
  for w in ${WORKERS:-1 4}; do
    MODSGPU_CNN_CHUNK=$c timeout 600 python bench.py --steps ${STEPS:-16} --warmup 3 --workers $w --no-cpu-baseline > gpurun_out/bench_c${c}_w$w.json 2> gpurun_out/bench_c${c}_w$w.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c${c}_w$w.json").read().strip().splitlines()[-1])
    ks=d["kernels"]; conv=sum(v["ms_per_step"] for k,v in ks.items() if "conv" in k or "head" in k)
    print("chunk $c workers $w value %.1f e2e %.1f pairs/s  ms/step %.2f launches %d  cnn kernel ms/step %.2f sum %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"],conv,sum(v["ms_per_step"] for v in ks.values())))
except Exception as e:
    print("bench c$c w$w failed", e); print(open("gpurun_out/bench_c${c}_w$w.err").read()[-1500:])
PY
  done
done
