#!/bin/bash
# Round-end session on one GPU: parity tests, both bench arms, ncu launch list of one warm pair, full captures.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit=$?"; tail -2 gpurun_out/tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench_final.json 2>&1 | head -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][-30:], "clocks", d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_final_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 388 -c 200 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 3 > gpurun_out/ncu_list.log 2>&1; tail -1 gpurun_out/ncu_list.log
for k in k_conv_umma k_conv1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_KSKIP:-20} -c ${NCU_KCOUNT:-15} -f -o gpurun_out/prof_$k python tools/ncu_target.py 3 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
