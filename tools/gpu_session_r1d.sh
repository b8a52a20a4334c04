#!/bin/bash
# Session D: parity tests, A/B of the conv tile deal (contiguous vs strided), launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_contig.json 2> gpurun_out/bench_contig.err; python tools/show_bench.py gpurun_out/bench_contig.json | head -30
MODSGPU_CONV_STRIDED_TILES=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_strided.json 2> gpurun_out/bench_strided.err; python tools/show_bench.py gpurun_out/bench_strided.json | head -24
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${NCU_SKIP:-470} -c 600 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 2 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
