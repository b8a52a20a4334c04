#!/usr/bin/env python
"""Pretty-print the per-kernel table of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f e2e %.1f pairs/s  ms/step %.2f  launches/step %.0f  workers %s" % (
    d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"] / d["steps"], d["config"].get("workers_per_gpu")))
tot = 0
for k, v in d["kernels"].items():
    print("%-34s %7.3f ms %6.1f launches %5.1f%%  %s %s%s" % (k, v.get("ms_per_pair", v.get("ms_per_step")), v.get("launches_per_pair", v.get("launches_per_step")), 100 * v["share"],
          v["achieved"] and round(v["achieved"], 2), v["unit"],
          ("  | %.0f GB/s algorithmic" % v["hbm_gbs"]) if v.get("hbm_gbs") else ""))
    tot += v.get("ms_per_pair", v.get("ms_per_step"))
print("sum kernel ms/step %.2f" % tot)
print("stages", {k: (round(v["ms_per_pair"], 3) if isinstance(v, dict) else round(v, 3)) for k, v in d.get("stages", {}).items()})
print("host_cpu_ms_per_pair", d.get("host_cpu_ms_per_pair"), "single_pair", d.get("single_pair"))
print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "avg_us")})
