#!/bin/bash
# throughput with few contexts (what a single `mods` process with 1-4 pairs in flight sees)
mkdir -p gpurun_out/r3
for wk in 1 2 4 8; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workers $wk > gpurun_out/r3/bench_lowwk$wk.json 2> gpurun_out/r3/bench_lowwk$wk.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r3/bench_lowwk$wk.json"))
print("workers $wk: value %.1f e2e %.1f pairs/s (%.2f ms per pair per context) host_cpu %.2f" % (d["value"], d["e2e"]["value"], 1e3 * $wk / d["value"], d["host_cpu_ms_per_pair"]))
PY
done
