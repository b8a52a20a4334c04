#!/bin/bash
# Host-core sensitivity on one GPU: the bench pinned to 4 cores (the 8-GPU box has 4 per rank), spin vs sleeping waits
mkdir -p gpurun_out
run() {  # name, env..., then the command
  name=$1; shift
  env "$@" > gpurun_out/hc_$name.json 2> gpurun_out/hc_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/hc_$name.json").read().strip().splitlines()[-1])
    print("%-28s value %6.1f e2e %6.1f  workers %2s host_cpu_ms/step %.2f" % ("$name", d["value"], d["e2e"]["value"], d["config"]["workers_per_gpu"], d.get("host_cpu_ms_per_step", -1)))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/hc_$name.err").read()[-800:])
PY
}
B="python bench.py --no-cpu-baseline --steps 96 --warmup 3"
run all16_w16_sleep   X=1 $B
run all16_w16_spin    MODSGPU_SPIN_SYNC=1 $B
run c4_w4_spin        MODSGPU_SPIN_SYNC=1 taskset -c 0-3 $B --workers 4
run c4_w16_spin       MODSGPU_SPIN_SYNC=1 taskset -c 0-3 $B --workers 16
run c4_w4_sleep       X=1 taskset -c 0-3 $B --workers 4
run c4_w8_sleep       X=1 taskset -c 0-3 $B --workers 8
run c4_w16_sleep      X=1 taskset -c 0-3 $B --workers 16
run c4_w24_sleep      X=1 taskset -c 0-3 $B --workers 24
run c4_w16_sleep_st4  MODSGPU_CONV_STAGES=4 taskset -c 0-3 $B --workers 16
run all16_w16_sleep_st3  MODSGPU_CONV_STAGES=3 $B
run all16_w16_sleep_st4  MODSGPU_CONV_STAGES=4 $B
