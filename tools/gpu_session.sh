#!/bin/bash
# One GPU session: parity tests, bench at several worker counts, ncu launch list + full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/tests.log
for w in ${WORKERS:-1 2 4}; do
  timeout 600 python bench.py --steps ${STEPS:-16} --warmup 3 --workers $w --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_w$w.json").read().strip().splitlines()[-1])
    print("workers $w value %.1f e2e %.1f pairs/s  ms/step %.2f launches %d"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["gpu_launches"]), d["roofline"]["kernel"], "%.4f"%d["roofline"]["frac"])
except Exception as e:
    print("bench w$w failed", e); print(open("gpurun_out/bench_w$w.err").read()[-1500:])
PY
done
if [ -n "$NCU" ]; then
  # launch list of the second (warm) pair: ~470 launches per pair
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-470} -c 600 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 2 > gpurun_out/ncu_list.log 2>&1
  tail -2 gpurun_out/ncu_list.log
  for k in $NCU; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_KSKIP:-6} -c ${NCU_KCOUNT:-2} -f -o gpurun_out/prof_$k python tools/ncu_target.py 2 > gpurun_out/ncu_$k.log 2>&1
    tail -1 gpurun_out/ncu_$k.log
  done
fi
