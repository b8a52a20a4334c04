#!/bin/bash
# usage: tools/gpu_ncu_one.sh KERNEL_REGEX SKIP COUNT TAG -- one `ncu --set full` capture with source, report under gpurun_out/r3
mkdir -p gpurun_out/r3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-0} -c ${3:-1} -f -o gpurun_out/r3/prof_$4 python tools/ncu_target.py 2 > gpurun_out/r3/ncu_$4.log 2>&1
tail -2 gpurun_out/r3/ncu_$4.log
