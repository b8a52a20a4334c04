#!/bin/bash
# Session E: net parity tests, bench, full captures of the kernels named in NCU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider ${TESTSEL:+-k "$TESTSEL"} > gpurun_out/tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; python tools/show_bench.py gpurun_out/bench_e.json 2>&1 | head -${SHOW:-32}
for k in $NCU; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s ${NCU_KSKIP:-6} -c ${NCU_KCOUNT:-2} -f -o gpurun_out/prof_$k python tools/ncu_target.py 2 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
