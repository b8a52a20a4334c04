#!/bin/bash
# End-of-round validation on one GPU: parity tests, smoke(), the driver's two bench commands, the ncu launch list of a warm pair
mkdir -p gpurun_out/r3
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r3/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r3/final_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3/final_bench_n1.json 2> gpurun_out/r3/final_bench_n1.err
echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r3/final_bench_n1.json | grep -E "^value|stages|single|roofline"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3/final_bench_n1.json"))
print("cpu_baseline", d.get("cpu_baseline"), "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
PY
python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/r3/final_reference_n1.json 2> gpurun_out/r3/final_reference_n1.err
echo "reference rc=$?"; cut -c1-400 gpurun_out/r3/final_reference_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 400 --csv --log-file gpurun_out/r3/final_launches.csv python tools/ncu_target.py 2 > gpurun_out/r3/final_ncu_list.log 2>&1
tail -1 gpurun_out/r3/final_ncu_list.log
